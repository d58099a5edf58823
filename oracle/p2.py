"""Oracle P2 (quadratic Lagrange) finite elements on simplices -- test infrastructure only.

BASELINE ``configs[4]`` asks for a P2 discretisation of the state / adjoint systems (the
reference forces the *deformation* space to vector CG1, ``cashocs/_optimization/
shape_optimization/shape_optimization_problem.py:251-256``, so "P2" is the space of ``states`` /
``adjoints`` handed to ``ShapeOptimizationProblem``).  The element tensors are what FFC generates
for ``FunctionSpace(mesh, "CG", 2)`` [third party, restated from the textbook definition]:

* dofs: one per vertex, then one per edge (edge ``e`` -> dof ``Nv + e``); edges are the sorted
  unique vertex pairs, local edge order (0,1),(0,2),(0,3),(1,2),(1,3),(2,3) (2-D: (0,1),(0,2),(1,2)).
  dolfin's own dof numbering is a graph reordering that cannot be reproduced;
  the contract takes the cell->dof table as input, like for P1.
* basis in barycentric coordinates: vertex ``phi_i = lam_i (2 lam_i - 1)``, edge
  ``phi_ij = 4 lam_i lam_j``.

**Parity unpinned**: the reference holds no golden values for P2 systems; the oracle is validated
by patch tests (quadratics are reproduced exactly), by the P1/P2 agreement of the cost functional
under refinement and by the Taylor test of the shape gradient (rate 2), see ``tests/test_oracle.py``.
"""

from __future__ import annotations

import numpy as np

from oracle import fem

LOCAL_EDGES = {2: ((0, 1), (0, 2), (1, 2)), 3: ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))}


def edges_from_cells(cells: np.ndarray):
    """(edges [Ne, 2] sorted unique, cell_edges [Nc, ne]) -- cells ascending per cell."""
    d = cells.shape[1] - 1
    loc = LOCAL_EDGES[d]
    pairs = np.stack([np.stack([cells[:, i], cells[:, j]], axis=1) for i, j in loc], axis=1).astype(np.int64)
    lo = pairs.min(axis=2)
    hi = pairs.max(axis=2)
    nv = int(cells.max()) + 1
    key = lo * nv + hi
    uniq, inv = np.unique(key.reshape(-1), return_inverse=True)
    edges = np.stack([uniq // nv, uniq % nv], axis=1)
    return edges, inv.reshape(key.shape)


def cell_dofs(cells: np.ndarray, nv: int):
    """P2 cell->dof table [Nc, nd]: vertices, then ``nv + edge``; also returns the edge list."""
    edges, cell_edges = edges_from_cells(cells)
    return np.concatenate([cells.astype(np.int64), nv + cell_edges], axis=1), edges


def basis(lam: np.ndarray):
    """phi [nq, nd] and dphi/dlam [nq, nd, d+1] at barycentric points lam [nq, d+1]."""
    nq, k = lam.shape
    d = k - 1
    loc = LOCAL_EDGES[d]
    nd = k + len(loc)
    phi = np.zeros((nq, nd))
    dphi = np.zeros((nq, nd, k))
    for i in range(k):
        phi[:, i] = lam[:, i] * (2.0 * lam[:, i] - 1.0)
        dphi[:, i, i] = 4.0 * lam[:, i] - 1.0
    for e, (i, j) in enumerate(loc):
        phi[:, k + e] = 4.0 * lam[:, i] * lam[:, j]
        dphi[:, k + e, i] = 4.0 * lam[:, j]
        dphi[:, k + e, j] = 4.0 * lam[:, i]
    return phi, dphi


def reference_tensors(d: int):
    """Cell-independent tensors of the affine P2 element (relative to the cell volume):

    ``R[a, b, k, l] = int dphi_a/dlam_k dphi_b/dlam_l``  (stiffness: ``K_ab = |K| sum_kl R G_k.G_l``),
    ``M[a, b] = int phi_a phi_b``,  ``I[a] = int phi_a``.
    """
    lam, w = fem.quadrature(d, 4)
    phi, dphi = basis(lam)
    R = np.einsum("q,qak,qbl->abkl", w, dphi, dphi)
    M = np.einsum("q,qa,qb->ab", w, phi, phi)
    I = np.einsum("q,qa->a", w, phi)
    return R, M, I


def stiffness_elements(vol, G):
    d = G.shape[2]
    R, _, _ = reference_tensors(d)
    gg = np.einsum("cki,cli->ckl", G, G)
    return vol[:, None, None] * np.einsum("abkl,ckl->cab", R, gg)


def mass_elements(vol, d):
    _, M, _ = reference_tensors(d)
    return vol[:, None, None] * M[None]


def load_elements_callable(coords, cells, vol, f, degree):
    """b_a = int_K f phi_a, f a callable of x, exact for polynomial f of the given degree."""
    d = coords.shape[1]
    lam, w = fem.quadrature(d, degree + 2)
    phi, _ = basis(lam)
    xq = np.einsum("qa,cad->cqd", lam, coords[cells])
    return vol[:, None] * np.einsum("cq,q,qa->ca", f(xq), w, phi)


def integral(u, dofs, vol, d):
    """int u dx for a P2 function."""
    _, _, I = reference_tensors(d)
    return float(np.sum(vol * (u[dofs] @ I)))


def interpolate(f, coords, edges):
    """Nodal P2 interpolant: vertex values, then edge-midpoint values."""
    mid = 0.5 * (coords[edges[:, 0]] + coords[edges[:, 1]])
    return np.concatenate([f(coords), f(mid)])


def boundary_dofs(mesh, markers, edges):
    """Dirichlet dofs of a P2 space on the facets carrying ``markers`` (dolfin topological search):
    the facets' vertices and the edges of those facets."""
    markers = list(markers)
    sel = np.isin(mesh.facet_tags, markers)
    facets = np.sort(mesh.facets[sel], axis=1)
    verts = np.unique(facets.reshape(-1))
    nv = mesh.num_vertices
    key = edges[:, 0] * nv + edges[:, 1]
    d = mesh.dim
    if d == 2:
        fe = facets[:, 0] * nv + facets[:, 1]
    else:
        fe = np.concatenate([facets[:, i] * nv + facets[:, j] for i, j in ((0, 1), (0, 2), (1, 2))])
    eid = np.searchsorted(key, np.unique(fe))
    return np.concatenate([verts, nv + eid]).astype(np.int64)


def shape_derivative_poisson(coords, cells, dofs, u, p, f, grad_f, f_degree, c_int_u=1.0):
    """Cell vectors [Nc, (d+1)*d] of dJ[V] for P1 vector test functions ``V = phi_a e_i`` and P2
    state / adjoint (analytic form ``tests/test_shape_optimization.py:446-451``):

    ``G_a[i] (c int u + int gu.gp - int f p) - int (G_a.gu) d_i p - int (G_a.gp) d_i u - int d_i f phi_a p``.
    """
    d = coords.shape[1]
    _, vol, G = fem.cell_geometry(coords, cells)
    lam, w = fem.quadrature(d, f_degree + 2)
    phi, dphi = basis(lam)
    ue, pe = u[dofs], p[dofs]  # [Nc, nd]
    uq = np.einsum("qa,ca->cq", phi, ue)
    pq = np.einsum("qa,ca->cq", phi, pe)
    # physical gradients at the quadrature points: sum_a u_a sum_k dphi_a/dlam_k G_k
    gu = np.einsum("ca,qak,cki->cqi", ue, dphi, G)
    gp = np.einsum("ca,qak,cki->cqi", pe, dphi, G)
    xq = np.einsum("qa,cad->cqd", lam, coords[cells])
    fq = f(xq)
    gfq = grad_f(xq)[..., :d]
    s0 = vol * np.einsum("q,cq->c", w, c_int_u * uq + np.einsum("cqi,cqi->cq", gu, gp) - fq * pq)
    Ggu = np.einsum("cai,cqi->cqa", G, gu)
    Ggp = np.einsum("cai,cqi->cqa", G, gp)
    t1 = np.einsum("q,cqa,cqi->cai", w, Ggu, gp) + np.einsum("q,cqa,cqi->cai", w, Ggp, gu)
    t2 = np.einsum("q,qa,cq,cqi->cai", w, lam, pq, gfq)
    be = G * s0[:, None, None] - vol[:, None, None] * (t1 + t2)
    return be.reshape(cells.shape[0], -1)
