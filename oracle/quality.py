"""Oracle mesh quality, a-priori determinant test and collision counting.

Test infrastructure only (see ``oracle/__init__.py``).

Restates the reference's JIT-compiled C++ (``cashocs/geometry/quality.py:101-271``:
``angles_triangle`` :114-136, ``dihedral_angles`` :140-177, ``skewness`` :181-221,
``maximum_angle`` :225-262), the DG0 condition-number projection (``:406-461``), dolfin's
``MeshQuality.radius_ratios`` [third party: d * r_in / R_circ] (``:371-400``), the reductions
``MeshQuality.min/avg/quantile`` (``:464-546``), ``APrioriMeshTester.test``
(``cashocs/geometry/mesh_testing.py:84-132``) and ``compute_collisions`` (``:201-241``).
"""

from __future__ import annotations

import math

import numpy as np


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def cell_angles(coords: np.ndarray, cells: np.ndarray) -> np.ndarray:
    """Interior angles (2D, [Nc,3]) or dihedral angles (3D, [Nc,6]) exactly as the C++ does."""
    x = coords[cells]
    if coords.shape[1] == 2:
        p0, p1, p2 = x[:, 0], x[:, 1], x[:, 2]
        e0, e1, e2 = _unit(p1 - p0), _unit(p2 - p0), _unit(p2 - p1)
        dots = np.stack(
            [np.sum(e0 * e1, -1), -np.sum(e0 * e2, -1), np.sum(e1 * e2, -1)], axis=1
        )
    else:
        p0, p1, p2, p3 = x[:, 0], x[:, 1], x[:, 2], x[:, 3]
        e0, e1, e2, e3, e4 = p1 - p0, p2 - p0, p3 - p0, p2 - p1, p3 - p1
        n0 = _unit(np.cross(e0, e1))
        n1 = _unit(np.cross(e0, e2))
        n2 = _unit(np.cross(e1, e2))
        n3 = _unit(np.cross(e3, e4))
        dots = np.stack(
            [
                np.sum(n0 * n1, -1),
                -np.sum(n0 * n2, -1),
                np.sum(n1 * n2, -1),
                np.sum(n0 * n3, -1),
                np.sum(n1 * (-n3), -1),
                np.sum(n2 * n3, -1),
            ],
            axis=1,
        )
    with np.errstate(invalid="ignore"):
        return np.arccos(dots)  # acos outside [-1,1] -> NaN, as in C++


def _opt_angle(d):
    return math.pi / 3.0 if d == 2 else math.acos(1.0 / 3.0)


def skewness(coords, cells):
    ang = cell_angles(coords, cells)
    opt = _opt_angle(coords.shape[1])
    q = 1.0 - np.maximum((ang - opt) / (math.pi - opt), (opt - ang) / opt)
    return q.min(axis=1)


def maximum_angle(coords, cells):
    ang = cell_angles(coords, cells)
    opt = _opt_angle(coords.shape[1])
    q = 1.0 - np.maximum((ang - opt) / (math.pi - opt), 0.0)
    return q.min(axis=1)


def condition_number(coords, cells):
    """sqrt(d) / (||J||_F ||J^-1||_F)  (``quality.py:439-451``; DG0 jacobi solve is the identity)."""
    d = coords.shape[1]
    x = coords[cells]
    J = np.transpose(x[:, 1:, :] - x[:, :1, :], (0, 2, 1))
    Ji = np.linalg.inv(J)
    nf = np.sqrt(np.sum(J * J, axis=(1, 2))) * np.sqrt(np.sum(Ji * Ji, axis=(1, 2)))
    return math.sqrt(d) / nf


def radius_ratios(coords, cells):
    """dolfin ``Cell::radius_ratio`` = d * inradius / circumradius [third party, recalled]."""
    d = coords.shape[1]
    x = coords[cells]
    if d == 2:
        a = np.linalg.norm(x[:, 1] - x[:, 2], axis=1)
        b = np.linalg.norm(x[:, 0] - x[:, 2], axis=1)
        c = np.linalg.norm(x[:, 0] - x[:, 1], axis=1)
        e1, e2 = x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]
        area = 0.5 * np.abs(e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0])
        r_in = 2.0 * area / (a + b + c)
        r_circ = a * b * c / (4.0 * area)
        return 2.0 * r_in / r_circ
    p0, p1, p2, p3 = x[:, 0], x[:, 1], x[:, 2], x[:, 3]
    vol = np.abs(np.einsum("ci,ci->c", p1 - p0, np.cross(p2 - p0, p3 - p0))) / 6.0

    def tri_area(a, b, c):
        return 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)

    area = tri_area(p1, p2, p3) + tri_area(p0, p2, p3) + tri_area(p0, p1, p3) + tri_area(p0, p1, p2)
    r_in = 3.0 * vol / area
    # circumradius: sqrt((aA+bB+cC)(aA+bB-cC)(aA-bB+cC)(-aA+bB+cC)) / (24 V), opposite edge products
    la = np.linalg.norm(p1 - p0, axis=1) * np.linalg.norm(p3 - p2, axis=1)
    lb = np.linalg.norm(p2 - p0, axis=1) * np.linalg.norm(p3 - p1, axis=1)
    lc = np.linalg.norm(p3 - p0, axis=1) * np.linalg.norm(p2 - p1, axis=1)
    s = (la + lb + lc) * (la + lb - lc) * (la - lb + lc) * (-la + lb + lc)
    r_circ = np.sqrt(np.maximum(s, 0.0)) / (24.0 * vol)
    return 3.0 * r_in / r_circ


MEASURES = {
    "skewness": skewness,
    "maximum_angle": maximum_angle,
    "radius_ratios": radius_ratios,
    "condition_number": condition_number,
}


def compute_mesh_quality(coords, cells, quality_type="min", quality_measure="skewness", quantile=0.0):
    """``cashocs.compute_mesh_quality`` (``cashocs/geometry/quality.py:41-95``)."""
    q = MEASURES[quality_measure](coords, cells)
    if quality_type in ("min", "minimum"):
        return float(np.min(q))
    if quality_type in ("avg", "average"):
        return float(np.average(q))
    if quality_type == "quantile":
        return float(np.quantile(q, quantile))
    raise ValueError(quality_type)


def deformation_determinants(coords, cells, V):
    """det(I + grad V) per cell for a vertex field V [Nv, d] (``mesh_testing.py:61-73``)."""
    d = coords.shape[1]
    x = coords[cells]
    J = np.transpose(x[:, 1:, :] - x[:, :1, :], (0, 2, 1))
    v = V[cells]
    dV = np.transpose(v[:, 1:, :] - v[:, :1, :], (0, 2, 1))
    gradV = dV @ np.linalg.inv(J)
    return np.linalg.det(np.eye(d)[None] + gradV)


def a_priori_test(coords, cells, V, volume_change=float("inf")):
    """``APrioriMeshTester.test``: min det >= 1/volume_change and max det <= volume_change."""
    det = deformation_determinants(coords, cells, V)
    mn, mx = float(det.min()), float(det.max())
    return bool((mn >= 1.0 / volume_change) and (mx <= volume_change)), mn, mx


def gradient_frobenius_max(coords, cells, D):
    """max_K sqrt(grad d : grad d) (``cashocs/geometry/mesh_handler.py:317-327,366``)."""
    x = coords[cells]
    J = np.transpose(x[:, 1:, :] - x[:, :1, :], (0, 2, 1))
    v = D[cells]
    dV = np.transpose(v[:, 1:, :] - v[:, :1, :], (0, 2, 1))
    g = dV @ np.linalg.inv(J)
    return float(np.sqrt(np.sum(g * g, axis=(1, 2))).max())


def vertex_valence(cells, nv):
    """``IntersectionTester.__init__`` occurrences (``mesh_testing.py:148-158``)."""
    return np.bincount(cells.reshape(-1), minlength=nv).astype(np.int32)


def compute_collisions(coords, cells, tol=0.0):
    """Per vertex: number of closed cells containing the vertex point (brute force, small meshes).

    Restates ``compute_collisions`` (``mesh_testing.py:201-241``) whose arithmetic is dolfin's
    ``BoundingBoxTree::compute_entity_collisions`` -> exact ``CollisionPredicates`` [third party].
    Point-in-simplex is decided with barycentric coordinates and a tolerance ``tol``
    (0 = closed simplex in floating point); vertices of the cell itself always count.
    """
    nv = coords.shape[0]
    d = coords.shape[1]
    x = coords[cells]
    J = np.transpose(x[:, 1:, :] - x[:, :1, :], (0, 2, 1))
    Ji = np.linalg.inv(J)
    counts = np.zeros(nv, dtype=np.int32)
    lo = x.min(axis=1)
    hi = x.max(axis=1)
    for v in range(nv):
        p = coords[v]
        cand = np.nonzero(np.all((p >= lo - 1e-12) & (p <= hi + 1e-12), axis=1))[0]
        if cand.size == 0:
            continue
        lam = np.einsum("cij,cj->ci", Ji[cand], p[None, :] - x[cand, 0, :])
        lam0 = 1.0 - lam.sum(axis=1)
        inside = np.all(lam >= -tol, axis=1) & (lam0 >= -tol)
        own = np.any(cells[cand] == v, axis=1)
        counts[v] = int(np.count_nonzero(inside | own))
    return counts


def compute_collisions_tree(coords, cells, tol=0.0, chunk=200_000):
    """Same counts as :func:`compute_collisions` with the candidate search done by a spatial tree, as dolfin does
    (``BoundingBoxTree`` [third party]; here ``scipy.spatial.cKDTree`` over the vertices): every cell asks for the vertices
    inside the cube around its bounding box, then the same bounding-box filter and barycentric predicate decide.
    O((Nv + Nc) log Nv) instead of O(Nv Nc), so the CPU arm of ``bench.py`` can keep the intersection test switched on."""
    from scipy.spatial import cKDTree

    nv, d = coords.shape
    tree = cKDTree(coords)
    counts = np.zeros(nv, dtype=np.int64)
    for c0 in range(0, cells.shape[0], chunk):
        cc = cells[c0:c0 + chunk]
        x = coords[cc]
        lo, hi = x.min(axis=1), x.max(axis=1)
        centre = 0.5 * (lo + hi)
        radius = 0.5 * (hi - lo).max(axis=1) + 2e-12
        lists = tree.query_ball_point(centre, radius, p=np.inf, workers=-1)
        lens = np.fromiter((len(l) for l in lists), dtype=np.int64, count=len(lists))
        ci = np.repeat(np.arange(cc.shape[0]), lens)
        vi = np.fromiter((v for l in lists for v in l), dtype=np.int64, count=int(lens.sum()))
        p = coords[vi]
        keep = np.all((p >= lo[ci] - 1e-12) & (p <= hi[ci] + 1e-12), axis=1)
        ci, vi, p = ci[keep], vi[keep], p[keep]
        J = np.transpose(x[:, 1:, :] - x[:, :1, :], (0, 2, 1))
        Ji = np.linalg.inv(J)
        lam = np.einsum("cij,cj->ci", Ji[ci], p - x[ci, 0, :])
        lam0 = 1.0 - lam.sum(axis=1)
        inside = np.all(lam >= -tol, axis=1) & (lam0 >= -tol)
        own = np.any(cc[ci] == vi[:, None], axis=1)
        counts += np.bincount(vi[inside | own], minlength=nv)
    return counts.astype(np.int32)


def intersection_test(coords, cells, tol=1e-13):
    """``IntersectionTester.test`` (``mesh_testing.py:161-195``): collisions == valence everywhere."""
    collisions = compute_collisions if coords.shape[0] <= 2000 else compute_collisions_tree
    return bool(np.all(collisions(coords, cells, tol) == vertex_valence(cells, coords.shape[0])))
