"""Reduced cost functional (``cashocs/_optimization/cost_functional.py:44-88,146-206``)."""

from __future__ import annotations

import torch

from cashocs_b200 import _lib
from cashocs_b200._forms import forms


class ReducedCostFunctional:
    """J(u) evaluated through a state solve + one fused cell-loop reduction kernel."""

    def __init__(self, mesh, cost, states, state_problem, control=None, mass_scalar_product=None) -> None:
        self.mesh, self.cost, self.states = mesh, cost, states
        self.state_problem = state_problem
        self.control = control
        self.mass_scalar_product = mass_scalar_product
        self._out = torch.zeros(1, dtype=torch.float64, device=mesh.device)
        self.shape_regularization = None  # set by ShapeOptimizationProblem (cost_functional.py:84-86)
        if isinstance(cost, forms.CostFunctionalList):
            if states[0].space is not None:
                from cashocs_b200 import _exceptions

                raise _exceptions.InputError("cashocs_b200.ReducedCostFunctional", "cost_functional_form",
                                             "lists of functionals / scalar tracking terms are supported for P1 states")
            cost.bind(self._integrand_values)
            state_problem.after_solve = cost.update

    def scale(self, desired_weights) -> None:
        """``_scale_cost_functional`` (``cashocs/_optimization/optimization_problem.py:791-824``): every term of the
        list is scaled so that its value on the initial iterate is ``|desired_weights[i]|`` (a term that vanishes there
        is multiplied by its weight instead)."""
        from cashocs_b200 import _exceptions
        from cashocs_b200 import log

        c = self.cost
        if not isinstance(c, forms.CostFunctionalList) or len(desired_weights) != len(c.functionals):
            raise _exceptions.InputError("cashocs_b200.OptimizationProblem", "desired_weights",
                                         "Length of desired_weights and cost_functional does not match.")
        log.warning("You are using the automatic scaling functionality of cashocs."
                    "This may lead to unexpected results if you try to scale the cost "
                    "functional yourself or if you supply custom forms.")
        self.state_problem.solve()
        control_value = 0.0
        if c.c_control != 0.0 and self.control is not None:
            control_value = 0.5 * self.mass_scalar_product(self.control.vector, self.control.vector)
        values = c.term_values(self._integrand_values(), control_value)
        factors = []
        for i, (w, val) in enumerate(zip(desired_weights, values)):
            if abs(val) <= 1e-15:
                val = 1.0
                log.info(f"Term {i:d} of the cost functional vanishes for the initial iteration. Multiplying this term "
                         "with the factor you supplied in desired weights.")
            factors.append(abs(float(w) / val))
        c.scale(factors)
        c.update()
        self.initial_function_values, self.scaling_factors = values, factors

    def _integrand_values(self):
        """(int u, 1/2 int (u - u_d)^2, volume, boundary measure) at the current state, for CostFunctionalList."""
        c, u = self.cost, self.states[0].vector
        iu = self._integral(u, 1.0, 0.0, None)
        it = self._integral(u, 0.0, 1.0, c.desired_state.vector) if c.desired_state is not None else 0.0
        reg = self.shape_regularization
        vol = reg.compute_volume() if (c.has_volume_terms and reg is not None) else 0.0
        surf = reg.compute_surface() if (c.has_surface_terms and reg is not None) else 0.0
        if (c.has_volume_terms or c.has_surface_terms) and reg is None:
            from cashocs_b200 import _exceptions

            raise _exceptions.InputError("cashocs_b200.IntegralFunctional", "c_volume",
                                         "the geometric terms int 1 dx / int 1 ds belong to shape optimisation problems")
        return iu, it, vol, surf

    def _integral(self, u, c_u, c_track, ud) -> float:
        mesh = self.mesh
        scratch = mesh.reduce_scratch()
        space = self.states[0].space
        if space is not None:  # P2 state
            import ctypes as C

            view = space.view()
            _lib.check(
                _lib.load().cb_p2_functional(_lib.stream_ptr(), C.byref(view), float(c_u), _lib.ptr(u), float(c_track),
                                             _lib.ptr(ud), _lib.ptr(self._out), _lib.ptr(scratch), scratch.numel()),
                "cb_p2_functional",
            )
            if mesh.dist is not None:
                mesh.dist.allreduce(self._out, "sum")
            return float(self._out.item())
        _lib.check(
            _lib.load().cb_functional(_lib.stream_ptr(), mesh.dim, mesh.n_owned_cells, _lib.ptr(mesh.coords),
                                      _lib.ptr(mesh.cells), float(c_u), _lib.ptr(u), float(c_track), _lib.ptr(ud),
                                      _lib.ptr(self._out), _lib.ptr(scratch), scratch.numel()),
            "cb_functional",
        )
        if mesh.dist is not None:
            mesh.dist.allreduce(self._out, "sum")
        return float(self._out.item())

    def evaluate(self) -> float:
        """``ReducedCostFunctional.evaluate`` (``cost_functional.py:65-88``)."""
        self.state_problem.solve()
        c = self.cost
        reg = self.shape_regularization
        if isinstance(c, forms.CostFunctionalList):
            val = c.value()
        else:
            ud = c.desired_state.vector if c.c_track != 0.0 else None
            val = self._integral(self.states[0].vector, c.c_state, c.c_track, ud)
            if reg is not None and reg.is_active:
                val += reg.geometric_cost_terms()
        if c.c_control != 0.0 and self.control is not None:
            # alpha/2 int u^2 = alpha/2 u^T M u
            val += 0.5 * c.c_control * self.mass_scalar_product(self.control.vector, self.control.vector)
        if self.shape_regularization is not None and self.shape_regularization.is_active:
            val += self.shape_regularization.compute_objective()
        elif getattr(c, "has_volume_terms", getattr(c, "c_volume", 0.0) != 0.0) or \
                getattr(c, "has_surface_terms", getattr(c, "c_surface", 0.0) != 0.0):
            from cashocs_b200 import _exceptions

            raise _exceptions.InputError("cashocs_b200.IntegralFunctional", "c_volume",
                                         "the geometric terms int 1 dx / int 1 ds belong to shape optimisation problems")
        return val
