"""Adjoint system (``cashocs/_pde_problems/adjoint_problem.py``), linear branch ``:124-191``.

Adjoint equation ``derivative(Lagrangian, y, test) = 0`` with homogenised state BCs
(``cashocs/_forms/general_form_handler.py:170-244``).  For the self-adjoint Poisson family the
operator equals the state operator; the reference re-assembles it anyway,
here the state matrix is reused when the BC dof sets coincide -- results are unchanged.
"""

from __future__ import annotations

import torch

from cashocs_b200 import _exceptions
from cashocs_b200._forms import forms
from cashocs_b200._pde_problems import pde_problem
from cashocs_b200._utils import linalg


class AdjointProblem(pde_problem.PDEProblem):
    def __init__(self, mesh, state_forms, bcs_list, cost, states, adjoints, state_problem, ksp_options,
                 linear_solver=None) -> None:
        super().__init__(linear_solver)
        self.mesh = mesh
        self.state_forms = state_forms
        self.cost = cost
        self.states, self.adjoints = states, adjoints
        self.state_problem = state_problem
        self.ksp_options = ksp_options
        self.number_of_solves = 0
        self.iterations = [0] * len(adjoints)
        # homogenised BCs: same dofs, value 0
        self.bc_arrays = []
        self.spaces = [p.space for p in adjoints]
        for bcs, sp in zip(bcs_list, self.spaces):
            hom = [forms.DirichletBC(0.0, bc.markers, bc.component) for bc in bcs]
            self.bc_arrays.append(forms.bc_flag_arrays(mesh, hom, 1, space=sp))
        for y, p in zip(states, adjoints):
            if y.degree != p.degree:
                raise _exceptions.InputError("cashocs_b200.AdjointProblem", "adjoints",
                                             "state and adjoint must use the same CG degree")
        self.b_tensors = [torch.zeros(sp.n_owned_dofs if sp is not None else mesh.n_owned_vertices, dtype=torch.float64,
                                      device=mesh.device) for sp in self.spaces]
        self.A_tensors = [None] * len(adjoints)
        nwork = max(sp.ndofs if sp is not None else mesh.num_vertices for sp in self.spaces)
        self._rhs_work = torch.zeros(nwork, dtype=torch.float64, device=mesh.device)

    def _rhs_nodal(self, i: int) -> torch.Tensor:
        """Nodal field r with  -dJ/dy[v] = int r v dx :  r = -(c_state + c_track (y - y_d))."""
        r = self._rhs_work
        r.fill_(-self.cost.c_state)
        if self.cost.c_track != 0.0:
            r.add_(self.states[i].vector - self.cost.desired_state.vector, alpha=-self.cost.c_track)
        return r

    def solve(self):
        self.state_problem.solve()
        if not self.has_solution:
            for i in reversed(range(len(self.adjoints))):
                form = self.state_forms[i]
                flag, val = self.bc_arrays[i]
                reuse = form.is_self_adjoint and self.state_problem.A_tensors[i] is not None
                if self.spaces[i] is not None:  # P2: -dJ/dy[v] = -c_state int v - c_track int (y - y_d) v
                    kw = {}
                    if self.cost.c_track != 0.0:
                        r = self._rhs_work[: self.spaces[i].ndofs]
                        torch.sub(self.states[i].vector, self.cost.desired_state.vector, out=r)
                        kw = dict(f_nodal=r, f_nodal_scale=-self.cost.c_track)
                    A, b = linalg.assemble_p2_system(
                        self.spaces[i], form.kappa, form.reaction, f_const=-self.cost.c_state, bc_flag=flag,
                        bc_val=val, A=self.A_tensors[i], b=self.b_tensors[i], matrix=not reuse, **kw)
                    if reuse:
                        A = self.state_problem.A_tensors[i]
                    else:
                        self.A_tensors[i] = A
                    res = self.linear_solver.solve(self.adjoints[i].vector, A=A, b=b, ksp_options=self.ksp_options[i])
                    self.iterations[i] = res.its
                    if self.spaces[i].dist is not None:
                        self.spaces[i].dist.halo_exchange(self.adjoints[i].vector, 1)
                    continue
                A, b = linalg.assemble_scalar_system(
                    self.mesh, form.kappa, form.reaction, f_nodal=self._rhs_nodal(i), f_nodal_scale=1.0,
                    bc_flag=flag, bc_val=val, A=self.A_tensors[i], b=self.b_tensors[i], matrix=not reuse,
                )
                if reuse:
                    A = self.state_problem.A_tensors[i]
                else:
                    self.A_tensors[i] = A
                res = self.linear_solver.solve(self.adjoints[i].vector, A=A, b=b, ksp_options=self.ksp_options[i])
                self.iterations[i] = res.its
                if self.mesh.dist is not None:
                    self.mesh.dist.halo_exchange(self.adjoints[i].vector, 1)
            self.has_solution = True
            self.number_of_solves += 1
        return self.adjoints
