"""Form descriptors: the restricted UFL form family of the hot path as plain Python objects.

The reference derives its systems symbolically with UFL (``cashocs/_forms/general_form_handler.py``):
state ``derivative(F, p, test)`` split by ``ufl.system`` (``:101-140``), adjoint
``derivative(Lagrangian, y, test)`` with homogenised BCs (``:170-244``).  UFL is not available
here and symbolic differentiation is outside the hot path, so users state *which* member of the
supported family they mean; the descriptors carry exactly the data the element kernels need.
``fenics_adapter`` recognises these forms in UFL input when FEniCS is importable.

Supported family:
    F(u, p) = kappa * grad(u).grad(p) dx + c * u p dx + c3 * u^3 p dx - f p dx ,   u = g on Dirichlet markers
    J(u)    = int alpha_u * u dx  +  1/2 int (u - u_d)^2 dx  +  alpha/2 int control^2 dx
"""

from __future__ import annotations

import dataclasses

import torch

from cashocs_b200 import _exceptions
from cashocs_b200._forms import expressions


class Function:
    """A finite-element function: a device vector plus its value size (1 = scalar, d = vector).

    ``degree`` = 1 (P1, dof == vertex; the only choice for vector functions, the deformation space of
    the reference is CG1) or 2 (scalar P2: vertex dofs then edge dofs, see ``function_space.P2Space``).
    """

    def __init__(self, mesh, value_size: int = 1, name: str = "f", degree: int = 1) -> None:
        self.mesh = mesh
        self.value_size = int(value_size)
        self.name = name
        self.degree = int(degree)
        if self.degree not in (1, 2):
            raise _exceptions.InputError("cashocs_b200.Function", "degree", "CG degree must be 1 or 2")
        if self.degree == 2 and self.value_size != 1:
            raise _exceptions.InputError("cashocs_b200.Function", "degree", "P2 functions are scalar (states / adjoints)")
        self.space = mesh.p2_space() if self.degree == 2 else None
        n = self.space.ndofs if self.degree == 2 else mesh.num_vertices
        # ghost entries (distributed meshes) live at the tail
        self.vector = torch.zeros(n * self.value_size, dtype=torch.float64, device=mesh.device)

    def owned(self) -> torch.Tensor:
        if self.degree == 2:
            return self.vector[: self.space.n_owned_dofs]
        return self.vector[: self.mesh.n_owned_vertices * self.value_size]

    def assign(self, other) -> None:
        src = other.vector if isinstance(other, Function) else other
        self.vector.copy_(src)

    def copy(self) -> "Function":
        f = Function(self.mesh, self.value_size, self.name, self.degree)
        f.vector.copy_(self.vector)
        return f

    def interpolate(self, expr) -> None:
        """Nodal interpolation of a polynomial / callable of the node coordinates."""
        if self.degree == 2:
            self.vector.copy_(self.space.interpolate(expr))
            return
        x = self.mesh.coords
        if isinstance(expr, expressions.Polynomial):
            out = torch.zeros(x.shape[0], dtype=torch.float64, device=x.device)
            for (a, b, c), v in expr.terms.items():
                m = v * x[:, 0] ** a * x[:, 1] ** b
                if x.shape[1] == 3:
                    m = m * x[:, 2] ** c
                out += m
            self.vector.copy_(out)
        else:
            self.vector.copy_(expr(x).reshape(-1))


@dataclasses.dataclass
class DirichletBC:
    """``DirichletBC(V, value, boundaries, marker)`` for a (sub)space of a P1 space.

    ``component`` = None constrains all components (``V``), an int only ``V.sub(component)``.
    """

    value: float
    markers: tuple
    component: int | None = None

    def __post_init__(self):
        if isinstance(self.markers, int):
            self.markers = (self.markers,)
        self.markers = tuple(int(m) for m in self.markers)


def create_dirichlet_bcs(value: float, markers, component: int | None = None) -> list[DirichletBC]:
    """``cashocs.create_dirichlet_bcs`` (``cashocs/_utils/forms.py:214-271``): one BC per marker."""
    if isinstance(markers, int):
        markers = [markers]
    return [DirichletBC(value, (int(m),), component) for m in markers]


def bc_flag_arrays(mesh, bcs: list[DirichletBC], value_size: int, space=None):
    """Flag / value arrays over all dofs of a P1 space (or of the scalar P2 ``space``); later BCs
    overwrite earlier ones."""
    n = (space.ndofs if space is not None else mesh.num_vertices) * value_size
    flag = torch.zeros(n, dtype=torch.uint8, device=mesh.device)
    val = torch.zeros(n, dtype=torch.float64, device=mesh.device)
    any_bc = False
    for bc in bcs:
        verts = space.boundary_dofs(bc.markers) if space is not None else mesh.boundary_vertices(bc.markers)
        if verts.numel() == 0:
            continue
        any_bc = True
        comps = range(value_size) if bc.component is None else [bc.component]
        for c in comps:
            dofs = verts * value_size + c
            flag[dofs] = 1
            val[dofs] = float(bc.value)
    return (flag, val) if any_bc else (None, None)


@dataclasses.dataclass
class PoissonForm:
    """``F = kappa*inner(grad(u), grad(p))*dx + reaction*u*p*dx - source*p*dx``.

    ``source`` is a :class:`Polynomial` in the spatial coordinates (shape problems: it moves
    with the domain), a :class:`Function` (a P1 field, e.g. the control) or ``None``.
    """

    source: object = None
    kappa: float = 1.0
    reaction: float = 0.0
    source_scale: float = 1.0
    # semilinear member of the family: + cubic * u**3 * p * dx (tests/test_nonlinear_solvers.py:28-61,
    # tests/test_optimal_control.py:364-376); solved by Newton's method, the state config must have is_linear = False
    cubic: float = 0.0

    @property
    def is_self_adjoint(self) -> bool:
        return True

    @property
    def is_nonlinear(self) -> bool:
        return self.cubic != 0.0


@dataclasses.dataclass
class IntegralFunctional:
    """``J = int (c_state*u + c_track/2 (u-u_d)^2) dx + c_control/2 int control^2 dx + c_volume int 1 dx + c_surface int 1 ds``
    (``cashocs.IntegralFunctional``, ``cashocs/_optimization/cost_functional.py:146-206``); the last two are the
    geometric terms of shape problems (``Constant(alpha) * dx + Constant(alpha) * ds`` in
    ``demos/documented/shape_optimization/regularization/demo_regularization.py``)."""

    c_state: float = 0.0
    c_track: float = 0.0
    desired_state: Function | None = None
    c_control: float = 0.0
    c_volume: float = 0.0
    c_surface: float = 0.0

    def __post_init__(self):
        if self.c_track != 0.0 and self.desired_state is None:
            raise _exceptions.InputError("IntegralFunctional", "desired_state", "tracking term needs u_d")


@dataclasses.dataclass
class ScalarTrackingFunctional:
    """``weight / 2 (int integrand - tracking_goal)^2`` (``cashocs.ScalarTrackingFunctional``,
    ``cashocs/_optimization/cost_functional.py:209-316``); the integrand is an :class:`IntegralFunctional` of the state
    and the geometry (``c_state``, ``c_track`` -- e.g. ``u*u*dx`` is ``c_track = 2`` with a zero ``desired_state`` --,
    ``c_volume``, ``c_surface``)."""

    integrand: IntegralFunctional
    tracking_goal: float
    weight: float = 1.0

    def __post_init__(self):
        if self.integrand.c_control != 0.0:
            raise _exceptions.InputError("ScalarTrackingFunctional", "integrand",
                                         "integrands of the control are not supported on the device path")


_COEFFICIENTS = ("c_state", "c_track", "c_volume", "c_surface")


class CostFunctionalList:
    """A list of cost functionals, ``J = sum_i J_i`` (``cost_functional_form`` may be a list,
    ``cashocs/_optimization/optimization_problem.py:150-160``), seen by the adjoint problem, the shape derivative and
    the geometric terms as ONE :class:`IntegralFunctional` with *effective* coefficients: the derivative of a scalar
    tracking term is its integrand's derivative times ``weight (int integrand - goal)``, so after every state solve
    :meth:`update` folds those factors into ``c_state / c_track / c_volume / c_surface``.  All quadratic state terms
    must share one ``desired_state``."""

    def __init__(self, functionals) -> None:
        self._evaluator = None
        self._set(list(functionals))

    def _set(self, functionals) -> None:
        self.functionals = functionals  # in the caller's order (desired_weights refers to it)
        self.integral = [f for f in functionals if isinstance(f, IntegralFunctional)]
        self.tracking = [f for f in functionals if isinstance(f, ScalarTrackingFunctional)]
        if len(self.integral) + len(self.tracking) != len(functionals):
            raise _exceptions.InputError("cashocs_b200.CostFunctionalList", "cost_functional_form",
                                         "IntegralFunctional and ScalarTrackingFunctional terms are supported")
        terms = self.integral + [t.integrand for t in self.tracking]
        targets = [f.desired_state for f in terms if f.c_track != 0.0]
        if any(t is not targets[0] for t in targets[1:]):
            raise _exceptions.InputError("cashocs_b200.CostFunctionalList", "desired_state",
                                         "all quadratic state terms must use the same desired_state Function")
        self.desired_state = targets[0] if targets else None
        self.c_control = float(sum(f.c_control for f in self.integral))
        self.base = {k: float(sum(getattr(f, k) for f in self.integral)) for k in _COEFFICIENTS}
        self.has_surface_terms = any(f.c_surface != 0.0 for f in terms)
        self.has_volume_terms = any(f.c_volume != 0.0 for f in terms)
        for k in _COEFFICIENTS:
            setattr(self, k, self.base[k])

    def term_values(self, ints, control_value: float = 0.0) -> list[float]:
        """The value of every functional of the list; ``control_value`` = 1/2 int control^2."""
        iu, it, vol, surf = ints
        out = []
        for f in self.functionals:
            g = f.integrand if isinstance(f, ScalarTrackingFunctional) else f
            val = g.c_state * iu + g.c_track * it + g.c_volume * vol + g.c_surface * surf + g.c_control * control_value
            out.append(0.5 * f.weight * (val - f.tracking_goal) ** 2 if isinstance(f, ScalarTrackingFunctional) else val)
        return out

    def scale(self, factors) -> None:
        """``functional.scale(factor)`` for every term (``cost_functional.py:195-202,293-300``): integral functionals
        are multiplied by their factor, scalar tracking functionals get it as their weight.  The caller's objects are
        left alone, the list holds scaled copies."""
        scaled = []
        for f, factor in zip(self.functionals, factors):
            if isinstance(f, ScalarTrackingFunctional):
                scaled.append(dataclasses.replace(f, weight=float(factor)))
            else:
                scaled.append(dataclasses.replace(f, **{k: getattr(f, k) * float(factor)
                                                        for k in _COEFFICIENTS + ("c_control",)}))
        self._set(scaled)

    def bind(self, evaluator) -> None:
        """``evaluator()`` returns (int u, 1/2 int (u - u_d)^2, volume, boundary measure) for the current state."""
        self._evaluator = evaluator

    def tracking_values(self, ints=None):
        iu, it, vol, surf = ints if ints is not None else self._evaluator()
        return [t.integrand.c_state * iu + t.integrand.c_track * it + t.integrand.c_volume * vol
                + t.integrand.c_surface * surf for t in self.tracking]

    def update(self) -> None:
        """Effective coefficients for the current state (call after every state solve)."""
        eff = dict(self.base)
        for t, val in zip(self.tracking, self.tracking_values()):
            m = t.weight * (val - t.tracking_goal)
            for k in _COEFFICIENTS:
                eff[k] += m * getattr(t.integrand, k)
        for k in _COEFFICIENTS:
            setattr(self, k, eff[k])

    def value(self) -> float:
        """``sum_i J_i`` without the control term (added by the reduced cost functional)."""
        ints = self._evaluator()
        iu, it, vol, surf = ints
        b = self.base
        val = b["c_state"] * iu + b["c_track"] * it + b["c_volume"] * vol + b["c_surface"] * surf
        for t, v in zip(self.tracking, self.tracking_values(ints)):
            val += 0.5 * t.weight * (v - t.tracking_goal) ** 2
        return val


def as_cost_functional(cost, force_list: bool = False):
    """A single :class:`IntegralFunctional` stays what it is; lists and scalar tracking terms become a
    :class:`CostFunctionalList` (also a single functional that is to be scaled, ``force_list``)."""
    items = list(cost) if isinstance(cost, (list, tuple)) else [cost]
    if len(items) == 1 and isinstance(items[0], IntegralFunctional) and not force_list:
        return items[0]
    return CostFunctionalList(items)
