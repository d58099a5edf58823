"""Taylor tests (``cashocs/verification.py:182-302``) on the device pipeline."""

from __future__ import annotations

import numpy as np
import torch


def compute_convergence_rates(epsilons, residuals):
    return [float(np.log(residuals[i] / residuals[i - 1]) / np.log(epsilons[i] / epsilons[i - 1]))
            for i in range(1, len(epsilons))]


def shape_gradient_test(sop, h=None, rng=None) -> float:
    rng = rng or np.random
    mesh = sop.mesh
    if h is None:
        h = torch.from_numpy(rng.rand(mesh.num_vertices * mesh.dim)).to(mesh.device)
    h = h.clone()
    sop.form_handler.apply_shape_bcs(h)
    sop._erase_pde_memory()
    j0 = sop.reduced_cost_functional.evaluate()
    g = sop.compute_shape_gradient()[0].vector
    dj = sop.form_handler.scalar_product(g, h)
    length = float((mesh.coords.max() - mesh.coords.min()).item())
    epsilons = [length * 1e-4 / 2**i for i in range(4)]
    residuals = []
    for eps in epsilons:
        if sop.mesh_handler.move_mesh(h * eps):
            sop._erase_pde_memory()
            j1 = sop.reduced_cost_functional.evaluate()
            residuals.append(abs(j1 - j0 - eps * dj))
            sop.mesh_handler.revert_transformation()
        else:
            residuals.append(float("inf"))
    sop._erase_pde_memory()
    return compute_convergence_rates(epsilons, residuals)[-1]


def control_gradient_test(ocp, u=None, h=None, rng=None) -> float:
    """``cashocs.verification.control_gradient_test`` (``cashocs/verification.py:85-180``): Taylor test of the reduced
    gradient of an optimal control problem at ``u`` (default: the current control) in the direction ``h`` (default: a
    random one); returns the last convergence rate (2 for a correct gradient).  The control is restored afterwards."""
    rng = rng or np.random.RandomState()
    control = ocp.controls[0]
    initial = control.vector.clone()
    u0 = control.vector.clone() if u is None else (u.vector if hasattr(u, "vector") else u).clone()
    if h is None:
        h = torch.from_numpy(rng.rand(control.vector.numel())).to(control.vector.device)
    elif hasattr(h, "vector"):
        h = h.vector
    control.vector.copy_(u0)
    scaling = float(np.sqrt(ocp.form_handler.scalar_product(control.vector, control.vector)))
    if scaling < 1e-3:
        scaling = 1.0
    ocp._erase_pde_memory()
    j0 = ocp.reduced_cost_functional.evaluate()
    g = ocp.compute_gradient()[0].vector
    dj = ocp.form_handler.scalar_product(g, h)
    epsilons = [scaling * 1e-2 / 2**i for i in range(4)]
    residuals = []
    for eps in epsilons:
        control.vector.copy_(u0)
        control.vector.add_(h, alpha=eps)
        ocp._erase_pde_memory()
        j1 = ocp.reduced_cost_functional.evaluate()
        residuals.append(abs(j1 - j0 - eps * dj))
    control.vector.copy_(initial)
    ocp._erase_pde_memory()
    return compute_convergence_rates(epsilons, residuals)[-1]
