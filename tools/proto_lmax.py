#!/usr/bin/env python
"""CPU prototype (numpy / scipy on the oracle's elasticity matrices; diagnostic only): how accurate are k steps of the
power method -- the "last growth factor" the device's AMG set-up uses for lambda_max(D^-1 A) -- and k steps of Lanczos
on D^-1/2 A D^-1/2, from the same start vector, for the homogeneous cube and for mu from its Poisson solve?
Background and use: DESIGN.md section 8 (the smoother's Jacobi weight is hostage to this estimate).

usage: python tools/proto_lmax.py [n]"""
import os
import sys

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import meshes, pipeline  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    marks = (4, 8, 12, 16, 24, 32, 40)
    for name, kw in (("cube", {}), ("cube-mu", dict(shape_bdry_def=(2,), shape_bdry_fix=(1, 3, 4, 5, 6), mu_def=5e2, mu_fix=1.0))):
        om = meshes.regular_mesh(n, 1.0, 1.0, 1.0)
        cfg = pipeline.ShapeConfig(**({"shape_bdry_def": (1, 2, 3, 4, 5, 6)} | kw))
        prob = pipeline.PoissonShapeProblem(om, bc_markers=(1, 2, 3, 4, 5, 6), config=cfg)
        A = prob.A_elas.tocsr()
        d = A.diagonal()
        S = sps.diags(1 / np.sqrt(d)) @ A @ sps.diags(1 / np.sqrt(d))
        lam_true = spla.eigsh(S, k=1, which="LA", return_eigenvectors=False, tol=1e-8)[0]
        x0 = np.random.RandomState(1234).rand(A.shape[0]) + 0.5
        B = sps.diags(1 / d) @ A
        x, growth = x0.copy(), {}
        for k in range(1, max(marks) + 1):
            y = B @ x
            g = np.linalg.norm(y) / np.linalg.norm(x)
            x = y / np.linalg.norm(y)
            if k in marks:
                growth[k] = g
        v = np.sqrt(d) * x0
        V, alphas, betas, ritz, beta = [v / np.linalg.norm(v)], [], [], {}, 0.0
        for k in range(1, max(marks) + 1):
            w = S @ V[-1] - (beta * V[-2] if len(V) > 1 else 0.0)
            a = w @ V[-1]
            w = w - a * V[-1]
            alphas.append(a)
            if k in marks:
                ritz[k] = np.linalg.eigvalsh(np.diag(alphas) + np.diag(betas, 1) + np.diag(betas, -1))[-1]
            beta = np.linalg.norm(w)
            betas.append(beta)
            V.append(w / beta)
        print(f"{name}: n = {n}, {A.shape[0]} dofs, lambda_max(D^-1 A) = {lam_true:.6f}")
        for k in marks:
            print(f"  k = {k:2d}: power-method growth factor {100 * (1 - growth[k] / lam_true):5.2f} % low,  "
                  f"Lanczos Ritz value {100 * (1 - ritz[k] / lam_true):5.2f} % low")


if __name__ == "__main__":
    main()
