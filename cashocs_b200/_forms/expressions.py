"""Coefficient expressions of the supported form family.

The reference lets users write arbitrary UFL expressions of ``SpatialCoordinate`` that FFC
JIT-compiles into the element kernels (``cashocs/_forms/general_form_handler.py:101-140``).
The device kernels take such coefficients as *polynomials in x* (monomial table in kernel
parameter space) together with a quadrature rule that integrates them exactly, so the element
integrals agree with any exact FFC/FIAT rule to round-off.
"""

from __future__ import annotations

import functools

import numpy as np
from scipy.special import roots_jacobi

from cashocs_b200 import _exceptions
from cashocs_b200 import _lib


class Polynomial:
    """Multivariate polynomial ``sum_t c_t x^ex y^ey z^ez`` with a tiny algebra (+, -, *, **)."""

    def __init__(self, terms: dict[tuple[int, int, int], float] | None = None) -> None:
        self.terms = {k: float(v) for k, v in (terms or {}).items() if v != 0.0}

    # -- constructors
    @staticmethod
    def constant(c: float) -> "Polynomial":
        return Polynomial({(0, 0, 0): c})

    @staticmethod
    def coordinate(i: int) -> "Polynomial":
        e = [0, 0, 0]
        e[i] = 1
        return Polynomial({tuple(e): 1.0})

    # -- algebra
    @staticmethod
    def _lift(o) -> "Polynomial":
        return o if isinstance(o, Polynomial) else Polynomial.constant(float(o))

    def __add__(self, o):
        o = self._lift(o)
        t = dict(self.terms)
        for k, v in o.terms.items():
            t[k] = t.get(k, 0.0) + v
        return Polynomial(t)

    __radd__ = __add__

    def __neg__(self):
        return Polynomial({k: -v for k, v in self.terms.items()})

    def __sub__(self, o):
        return self + (-self._lift(o))

    def __rsub__(self, o):
        return self._lift(o) - self

    def __mul__(self, o):
        o = self._lift(o)
        t: dict = {}
        for (a, b, c), v in self.terms.items():
            for (d, e, f), w in o.terms.items():
                k = (a + d, b + e, c + f)
                t[k] = t.get(k, 0.0) + v * w
        return Polynomial(t)

    __rmul__ = __mul__

    def __pow__(self, n: int):
        if int(n) != n or n < 0:
            raise _exceptions.InputError("Polynomial", "power", "only non-negative integer powers")
        r = Polynomial.constant(1.0)
        for _ in range(int(n)):
            r = r * self
        return r

    @property
    def degree(self) -> int:
        return max((sum(k) for k in self.terms), default=0)

    def derivative(self, i: int) -> "Polynomial":
        t: dict = {}
        for k, v in self.terms.items():
            if k[i] > 0:
                e = list(k)
                e[i] -= 1
                t[tuple(e)] = t.get(tuple(e), 0.0) + v * k[i]
        return Polynomial(t)

    def __call__(self, x: np.ndarray) -> np.ndarray:
        """Evaluate at points ``x[..., d]`` (host, for tests / interpolation)."""
        d = x.shape[-1]
        out = np.zeros(x.shape[:-1])
        for (a, b, c), v in self.terms.items():
            m = v * x[..., 0] ** a * x[..., 1] ** b
            if d == 3:
                m = m * x[..., 2] ** c
            elif c:
                continue
            out = out + m
        return out

    def struct(self) -> _lib.cb_poly:
        if len(self.terms) > _lib.POLY_MAXT:
            raise _exceptions.InputError("Polynomial", "terms", f"more than {_lib.POLY_MAXT} monomials")
        p = _lib.cb_poly()
        p.nterms = len(self.terms)
        for t, ((a, b, c), v) in enumerate(sorted(self.terms.items())):
            p.coef[t] = v
            p.ex[t], p.ey[t], p.ez[t] = a, b, c
        return p


def spatial_coordinate(dim: int) -> tuple[Polynomial, ...]:
    """``x = SpatialCoordinate(mesh)`` for polynomial coefficient expressions."""
    return tuple(Polynomial.coordinate(i) for i in range(dim))


@functools.lru_cache(maxsize=None)
def simplex_quadrature(dim: int, degree: int):
    """Collapsed Gauss-Jacobi rule, exact for ``degree``; barycentric points + weights (sum 1)."""
    m = max(degree, 0) // 2 + 1
    pts = []
    if degree in (4, 5):
        # symmetric rules with half the points of the collapsed 3^d rule (exactness to 1e-16 is checked in
        # tests/test_host.py): 7-point Radon rule on the triangle, 14-point degree-5 rule on the tetrahedron
        import itertools

        if dim == 2:
            s15 = 15.0 ** 0.5
            orbits = [((1.0 / 3.0,) * 3, 9.0 / 40.0)]
            for a, w_ in (((6.0 - s15) / 21.0, (155.0 - s15) / 1200.0), ((6.0 + s15) / 21.0, (155.0 + s15) / 1200.0)):
                orbits += [(perm, w_) for perm in sorted(set(itertools.permutations((a, a, 1.0 - 2.0 * a))))]
            lam = np.array([list(o[0]) + [0.0] for o in orbits])
        elif dim == 3:
            orbits = []
            for a, w_ in ((0.3108859192633006, 0.11268792571801585), (0.09273525031089122, 0.07349304311636196)):
                orbits += [(perm, w_) for perm in sorted(set(itertools.permutations((a, a, a, 1.0 - 3.0 * a))))]
            b = 0.04550370412564965
            orbits += [(perm, 0.042546020777081466) for perm in sorted(set(itertools.permutations((b, b, 0.5 - b, 0.5 - b))))]
            lam = np.array([list(o[0]) for o in orbits])
        else:
            raise _exceptions.InputError("simplex_quadrature", "dim", "dim must be 2 or 3")
        w = np.array([o[1] for o in orbits])
        return lam, w / w.sum()
    if dim == 2:
        xa, wa = roots_jacobi(m, 1.0, 0.0)
        xb, wb = roots_jacobi(m, 0.0, 0.0)
        for a, wa_ in zip(xa, wa):
            u = 0.5 * (a + 1.0)
            for b, wb_ in zip(xb, wb):
                v = 0.5 * (b + 1.0)
                px, py = u, v * (1.0 - u)
                pts.append(((1.0 - px - py, px, py, 0.0), wa_ * wb_))
    elif dim == 3:
        xa, wa = roots_jacobi(m, 2.0, 0.0)
        xb, wb = roots_jacobi(m, 1.0, 0.0)
        xc, wc = roots_jacobi(m, 0.0, 0.0)
        for a, wa_ in zip(xa, wa):
            u = 0.5 * (a + 1.0)
            for b, wb_ in zip(xb, wb):
                v = 0.5 * (b + 1.0)
                for c, wc_ in zip(xc, wc):
                    t = 0.5 * (c + 1.0)
                    px, py, pz = u, v * (1.0 - u), t * (1.0 - u) * (1.0 - v)
                    pts.append(((1.0 - px - py - pz, px, py, pz), wa_ * wb_ * wc_))
    else:
        raise _exceptions.InputError("simplex_quadrature", "dim", "dim must be 2 or 3")
    lam = np.array([p[0] for p in pts])
    w = np.array([p[1] for p in pts])
    return lam, w / w.sum()


def quadrature_struct(dim: int, degree: int) -> _lib.cb_quad:
    lam, w = simplex_quadrature(dim, degree)
    if len(w) > _lib.QUAD_MAXQ:
        raise _exceptions.InputError("quadrature", "degree", f"needs {len(w)} > {_lib.QUAD_MAXQ} points")
    q = _lib.cb_quad()
    q.nq = len(w)
    for i in range(len(w)):
        for a in range(4):
            q.lam[i][a] = lam[i, a]
        q.w[i] = w[i]
    return q
