"""Oracle P1 finite elements on simplices: element tensors, CSR pattern, symmetric BCs.

Test infrastructure only (see ``oracle/__init__.py``).  Vectorised numpy; every routine
names the reference call site it restates.

Conventions
-----------
* scalar P1 dof ``i`` == vertex ``i``; vector P1 dof ``d*v + c`` (vertex-blocked, the layout
  of dolfin's blocked dofmap for ``VectorFunctionSpace(mesh, "CG", 1)``).
* pattern == all (test dof, trial dof) pairs of every cell, because
  ``bilinear_boundary_form_modification`` (``cashocs/_utils/forms.py:274-288``) adds
  ``Constant(0)*dot(u,v)*dx``; the diagonal is always present (``keep_diagonal=True``,
  ``cashocs/_utils/linalg.py:153``); columns sorted ascending inside a row (PETSc AIJ).
* Dirichlet conditions are eliminated symmetrically *per element tensor*
  (dolfin ``SystemAssembler`` [third party, recalled -- SURVEY.md 9.3]): BC row/col zeroed,
  ``A_e[i,i] = 1``, ``b_e[i] = g`` and ``b_e -= A_e[:, i] g``; the global BC diagonal is the
  number of cells sharing the dof.
"""

from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp
from scipy.special import roots_jacobi

DOLFIN_EPS = 3.0e-16


# ----------------------------------------------------------------------------- geometry
def cell_geometry(coords: np.ndarray, cells: np.ndarray):
    """Return (detJ [Nc], vol [Nc], G [Nc, d+1, d]) with G[c, a] = grad(phi_a) on cell c."""
    d = coords.shape[1]
    x = coords[cells]  # [Nc, d+1, d]
    J = np.transpose(x[:, 1:, :] - x[:, :1, :], (0, 2, 1))  # columns = edge vectors
    detJ = np.linalg.det(J)
    Jinv = np.linalg.inv(J)  # rows a = grad(lambda_{a+1})
    G = np.empty((cells.shape[0], d + 1, d))
    G[:, 1:, :] = Jinv
    G[:, 0, :] = -Jinv.sum(axis=1)
    vol = np.abs(detJ) / math.factorial(d)
    return detJ, vol, G


def quadrature(d: int, degree: int):
    """Collapsed Gauss-Jacobi rule on the reference simplex, exact for ``degree``.

    Returns barycentric points ``lam [nq, d+1]`` and weights summing to 1 (i.e. relative to
    the cell volume).  Any exact rule gives the same integrals up to roundoff for the
    polynomial integrands of the supported form family, so FIAT's particular
    scheme [third party] does not need to be reproduced.
    """
    m = degree // 2 + 1
    if d == 2:
        x0, w0 = roots_jacobi(m, 0, 0)
        x1, w1 = roots_jacobi(m, 1, 0)
        u = 0.5 * (x1 + 1.0)  # weight (1-u)
        v = 0.5 * (x0 + 1.0)
        U, V = np.meshgrid(u, v, indexing="ij")
        W = np.outer(w1 / 4.0, w0 / 2.0)
        px = U.reshape(-1)
        py = (V * (1 - U)).reshape(-1)
        w = W.reshape(-1)
        lam = np.stack([1 - px - py, px, py], axis=1)
    else:
        x0, w0 = roots_jacobi(m, 0, 0)
        x1, w1 = roots_jacobi(m, 1, 0)
        x2, w2 = roots_jacobi(m, 2, 0)
        u = 0.5 * (x2 + 1.0)
        v = 0.5 * (x1 + 1.0)
        t = 0.5 * (x0 + 1.0)
        U, V, T = np.meshgrid(u, v, t, indexing="ij")
        W = w2[:, None, None] / 8.0 * w1[None, :, None] / 4.0 * w0[None, None, :] / 2.0
        px = U.reshape(-1)
        py = (V * (1 - U)).reshape(-1)
        pz = (T * (1 - U) * (1 - V)).reshape(-1)
        w = W.reshape(-1)
        lam = np.stack([1 - px - py - pz, px, py, pz], axis=1)
    w = w / w.sum()
    return lam, w


# ----------------------------------------------------------------------------- element tensors
def stiffness_elements(vol, G):
    """K_ab = |K| G_a . G_b  (``inner(grad(u), grad(v))*dx``)."""
    return vol[:, None, None] * np.einsum("cai,cbi->cab", G, G)


def mass_elements(vol, d):
    """M_ab = |K| (1 + delta_ab) / ((d+1)(d+2))  (``u*v*dx``)."""
    n = d + 1
    ref = (np.ones((n, n)) + np.eye(n)) / ((d + 1) * (d + 2))
    return vol[:, None, None] * ref[None]


def elasticity_elements(vol, G, mu_vertex, cells, lam, delta):
    """Element tensors of the Riesz scalar product, vertex-blocked ordering.

    ``2 mu inner(eps(U), eps(W)) + lambda div U div W + delta U.W``
    (``cashocs/_forms/shape_form_handler.py:665-685`` with ``volumes == 1``); mu is a CG1
    function (``:293-295``) so its cell integral against constants is |K| * mean(mu).
    Returns [Nc, (d+1)*d, (d+1)*d] with local index ``a*d + i``.
    """
    d = G.shape[2]
    n = d + 1
    mubar = mu_vertex[cells].mean(axis=1)
    gg = np.einsum("cak,cbk->cab", G, G)
    eye = np.eye(d)
    A = np.einsum("c,cab,ij->caibj", vol * mubar, gg, eye, optimize=True)
    A += np.einsum("c,caj,cbi->caibj", vol * mubar, G, G, optimize=True)
    A += lam * np.einsum("c,cai,cbj->caibj", vol, G, G, optimize=True)
    A += delta * np.einsum("cab,ij->caibj", mass_elements(vol, d), eye)
    return A.reshape(-1, n * d, n * d)


def load_elements_callable(coords, cells, vol, f, degree):
    """b_a = int_K f phi_a with ``f(x)`` a callable evaluated at quadrature points."""
    d = coords.shape[1]
    lam, w = quadrature(d, degree)
    xq = np.einsum("qa,cad->cqd", lam, coords[cells])
    fq = f(xq)  # [Nc, nq]
    return vol[:, None] * np.einsum("cq,q,qa->ca", fq, w, lam, optimize=True)


def load_elements_p1(vol, cells, f_vertex, d):
    """b_a = int_K f_h phi_a for a P1 function f_h  (= M_e f_e)."""
    return np.einsum("cab,cb->ca", mass_elements(vol, d), f_vertex[cells])


# ----------------------------------------------------------------------------- pattern + assembly
def csr_pattern(cell_dofs: np.ndarray, n: int):
    """(rowptr, col) of the dofmap graph; bit-exact contract for the CUDA pattern builder."""
    nd = cell_dofs.shape[1]
    rows = np.repeat(cell_dofs, nd, axis=1).reshape(-1).astype(np.int64)
    cols = np.tile(cell_dofs, (1, nd)).reshape(-1).astype(np.int64)
    key = np.unique(rows * n + cols)
    r = key // n
    c = (key % n).astype(np.int32)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, r + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr, c


def vector_cell_dofs(cells: np.ndarray, d: int) -> np.ndarray:
    """Vertex-blocked vector dofs, local index a*d + i -> d*cells[:, a] + i."""
    return (d * cells[:, :, None].astype(np.int64) + np.arange(d)[None, None, :]).reshape(
        cells.shape[0], -1
    )


def apply_bcs_elementwise(Ae, be, cell_dofs, bc_dofs, bc_vals, n):
    """dolfin SystemAssembler symmetric elimination on element tensors."""
    is_bc = np.zeros(n, dtype=bool)
    g = np.zeros(n)
    is_bc[bc_dofs] = True
    g[bc_dofs] = bc_vals
    m = is_bc[cell_dofs]  # [Nc, nd]
    ge = np.where(m, g[cell_dofs], 0.0)
    if be is None:
        be = np.zeros(cell_dofs.shape, dtype=float)
    if Ae is not None:
        be = be - np.einsum("cab,cb->ca", Ae, ge)
        Ae = Ae * (~m)[:, :, None] * (~m)[:, None, :]
        nd = cell_dofs.shape[1]
        idx = np.arange(nd)
        Ae[:, idx, idx] += m.astype(float)
    be = np.where(m, ge, be)
    return Ae, be


def assemble_matrix(Ae, cell_dofs, n):
    nd = cell_dofs.shape[1]
    rows = np.repeat(cell_dofs, nd, axis=1).reshape(-1)
    cols = np.tile(cell_dofs, (1, nd)).reshape(-1)
    A = sp.coo_matrix((Ae.reshape(-1), (rows, cols)), shape=(n, n)).tocsr()
    A.sort_indices()  # explicit zeros are kept: pattern == dofmap graph
    return A


def assemble_vector(be, cell_dofs, n):
    b = np.zeros(n)
    np.add.at(b, cell_dofs.reshape(-1), be.reshape(-1))
    return b


def ident_zeros(A: sp.csr_matrix) -> sp.csr_matrix:
    """``PETScMatrix.ident_zeros`` (``cashocs/_utils/linalg.py:167``): unit diagonal on zero rows."""
    rowsum = np.asarray(abs(A).sum(axis=1)).reshape(-1)
    zero = np.nonzero(rowsum < DOLFIN_EPS)[0]
    if zero.size:
        A = A.tolil()
        for i in zero:
            A[i, i] = 1.0
        A = A.tocsr()
        A.sort_indices()
    return A


def assemble_system(Ae, be, cell_dofs, n, bc_dofs=None, bc_vals=None):
    """``assemble_petsc_system`` (``cashocs/_utils/linalg.py:107-188``) for element tensors."""
    if bc_dofs is not None and len(bc_dofs):
        Ae, be = apply_bcs_elementwise(Ae, be, cell_dofs, np.asarray(bc_dofs), bc_vals, n)
    A = ident_zeros(assemble_matrix(Ae, cell_dofs, n)) if Ae is not None else None
    b = assemble_vector(be, cell_dofs, n) if be is not None else None
    return A, b
