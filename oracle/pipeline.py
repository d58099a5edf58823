"""Oracle pipeline: one shape-gradient evaluation, Armijo update, Taylor test, control gradient.

Test infrastructure only (see ``oracle/__init__.py``).

Restates, for the Poisson model problem of the reference's demos/tests
(``demos/documented/shape_optimization/shape_poisson/demo_shape_poisson.py``,
``tests/test_shape_optimization.py:97-118``):

* ``StateProblem.solve`` linear branch (``cashocs/_pde_problems/state_problem.py:142-172``),
* ``AdjointProblem.solve`` (``cashocs/_pde_problems/adjoint_problem.py:124-191``),
* ``Stiffness.compute`` (``cashocs/_forms/shape_form_handler.py:188-235``),
* ``ShapeFormHandler.update_scalar_product`` / ``scalar_product`` (``:714-780``),
* ``ShapeGradientProblem.solve`` (``cashocs/_pde_problems/shape_gradient_problem.py:117-173``),
* ``_MeshHandler.move_mesh`` (``cashocs/geometry/mesh_handler.py:264-303``),
* ``ShapeVariableAbstractions.update_optimization_variables``
  (``cashocs/_optimization/shape_optimization/shape_variable_abstractions.py:91-135``),
* the Armijo rule (``cashocs/_optimization/line_search/armijo_line_search.py:97-202``,
  ``line_search.py:152-214,251-279``),
* ``shape_gradient_test`` (``cashocs/verification.py:182-302``),
* ``ControlGradientProblem.solve`` (``cashocs/_pde_problems/control_gradient_problem.py:96-125``).
"""

from __future__ import annotations

import collections
import dataclasses
import types

import numpy as np

from oracle import fem
from oracle import krylov
from oracle import p2
from oracle import quality


# ----------------------------------------------------------------------------- source term
def demo_f(x):
    """f = 2.5 (x + 0.4 - y^2)^2 + x^2 + y^2 - 1 (``demo_shape_poisson.py``, tests ``:116``)."""
    X, Y = x[..., 0], x[..., 1]
    return 2.5 * (X + 0.4 - Y**2) ** 2 + X**2 + Y**2 - 1.0


def demo_grad_f(x):
    X, Y = x[..., 0], x[..., 1]
    g = np.zeros_like(x)
    t = X + 0.4 - Y**2
    g[..., 0] = 5.0 * t + 2.0 * X
    g[..., 1] = -10.0 * Y * t + 2.0 * Y
    return g


DEMO_F_DEGREE = 4

DEFAULT_KSP = {"ksp_type": "preonly", "pc_type": "lu"}


@dataclasses.dataclass
class ShapeConfig:
    """The ``[ShapeGradient]`` / ``[MeshQuality]`` / ``[LineSearch]`` keys the hot path reads
    (defaults ``cashocs/io/config.py:593-741``; values here = ``tests/config_sop.ini``)."""

    shape_bdry_def: tuple = (1,)
    shape_bdry_fix: tuple = ()
    shape_bdry_fix_x: tuple = ()
    shape_bdry_fix_y: tuple = ()
    shape_bdry_fix_z: tuple = ()
    fixed_dimensions: tuple = ()  # [ShapeGradient] fixed_dimensions (shape_form_handler.py:693-712)
    lambda_lame: float = 1.428571428571429
    damping_factor: float = 0.2
    mu_def: float = 0.35714285714285715
    mu_fix: float = 0.35714285714285715
    use_sqrt_mu: bool = False
    # distance-based mu (shape_form_handler.py:141-186,216-235; geometry/boundary_distance.py:191-282)
    use_distance_mu: bool = False
    dist_min: float = 1.0
    dist_max: float = 1.0
    mu_min: float = 1.0
    mu_max: float = 1.0
    boundaries_dist: tuple = ()
    smooth_mu: bool = False
    inhomogeneous: bool = False          # scale the form by 1 / (|K| / max|K|)^exponent (shape_form_handler.py:628-646)
    inhomogeneous_exponent: float = 1.0
    update_inhomogeneous: bool = False
    volume_change: float = float("inf")
    angle_change: float = float("inf")
    tol_lower: float = 0.0
    tol_upper: float = 1e-15
    measure: str = "skewness"
    qtype: str = "min"
    quantile: float = 0.1
    test_for_intersections: bool = True
    beta_armijo: float = 2.0
    epsilon_armijo: float = 1e-4
    initial_stepsize: float = 1.0
    safeguard_stepsize: bool = True
    # [Regularization] (cashocs/_forms/shape_regularization.py:131-222,318-510); surface / curvature need facet terms
    factor_volume: float = 0.0
    target_volume: float = 0.0
    use_initial_volume: bool = False
    factor_barycenter: float = 0.0
    target_barycenter: tuple = (0.0, 0.0, 0.0)
    use_initial_barycenter: bool = False
    factor_surface: float = 0.0          # :224-316
    factor_curvature: float = 0.0        # :512-646 (boundary mass solve for the mean-curvature vector)
    use_relative_scaling: bool = False   # every active term's factor is divided by its initial value (`scale`)
    target_surface: float = 0.0
    use_initial_surface: bool = False
    line_search: str = "armijo"          # [LineSearch] method: armijo | basic | polynomial
    # p-Laplace projection of the shape derivative (shape_gradient_problem.py:329-427)
    use_p_laplacian: bool = False
    p_laplacian_power: int = 2
    p_laplacian_stabilization: float = 0.0
    # re-extension of the gradient deformation from its boundary values (shape_gradient_problem.py:175-223)
    reextend_from_boundary: bool = False
    reextension_mode: str = "surface"    # or "normal": only the normal part of the boundary values (:225-299)
    polynomial_model: str = "cubic"
    factor_low: float = 0.1
    factor_high: float = 0.5


class PoissonShapeProblem:
    """-Laplace u = f, u = 0 on ``bc_markers``; J = int u dx; elasticity Riesz projection."""

    def __init__(
        self,
        mesh,
        f=demo_f,
        grad_f=demo_grad_f,
        f_degree=DEMO_F_DEGREE,
        bc_markers=(1,),
        config: ShapeConfig | None = None,
        state_ksp=None,
        adjoint_ksp=None,
        gradient_ksp=None,
        mu_ksp=None,
        state_degree: int = 1,
        c_state: float = 1.0,
        kappa: float = 1.0,
        reaction: float = 0.0,
        c_track: float = 0.0,
        u_d=None,
        use_pull_back: bool = True,
        c_volume: float = 0.0,
        c_surface: float = 0.0,
        scalar_tracking=(),
    ):
        """State ``kappa grad u . grad p + reaction u p - f p`` (= 0), cost
        ``J = c_state int u + c_track / 2 int (u - u_d)^2`` with a P1 (vertex-valued) ``u_d`` that is transported with the
        mesh -- the ``J_func`` case of ``tests/test_shape_optimization.py:247-299``; its gradient enters the shape
        derivative through the pull-back of ``shape_form_handler.py:507-531`` unless ``use_pull_back`` is off."""
        self.mesh = mesh
        self.c_state = float(c_state)  # J = c_state * int u dx (tests with J = Constant(0)*dx use 0)
        self.kappa, self.reaction, self.c_track = float(kappa), float(reaction), float(c_track)
        self.u_d = None if u_d is None else np.asarray(u_d, dtype=float)
        self.use_pull_back = bool(use_pull_back)
        # geometric terms of the cost functional itself, J += c_volume int 1 dx + c_surface int 1 ds
        # (demos/documented/shape_optimization/regularization/demo_regularization.py: Constant(alpha)*dx + Constant(alpha)*ds)
        self.c_volume, self.c_surface = float(c_volume), float(c_surface)
        # ScalarTrackingFunctional terms weight / 2 (int integrand - goal)^2 (cost_functional.py:209-316), the integrand from
        # the same family: dicts with c_state, c_track, c_volume, c_surface, goal, weight.  Their derivative is the
        # integrand's derivative times weight (int integrand - goal), i.e. effective coefficients that follow the state.
        self.scalar_tracking = [{**dict(c_state=0.0, c_track=0.0, c_volume=0.0, c_surface=0.0, weight=1.0), **t} for t in scalar_tracking]
        self._base = dict(c_state=self.c_state, c_track=self.c_track, c_volume=self.c_volume, c_surface=self.c_surface)
        if self.u_d is None and any(t["c_track"] != 0.0 for t in self.scalar_tracking):
            self.u_d = np.zeros(mesh.num_vertices)
        if self.scalar_tracking and state_degree == 2:
            raise ValueError("oracle: scalar tracking functionals are written for P1 states")
        if self.c_track != 0.0 and self.u_d is None:
            raise ValueError("oracle: the tracking term needs u_d")
        general = (self.kappa != 1.0 or self.reaction != 0.0 or self.c_track != 0.0)
        if state_degree == 2 and (c_state != 1.0 or general):
            raise ValueError("oracle: the P2 shape derivative is written for J = int u, kappa = 1, no reaction")
        self.state_degree = int(state_degree)  # CG degree of the state / adjoint space (configs[4]: 2)
        self.coords = mesh.coords  # moved in place by move_mesh
        self.cells = mesh.cells
        self.d = mesh.dim
        self.nv = mesh.num_vertices
        self.f, self.grad_f, self.f_degree = f, grad_f, f_degree
        self.cfg = config or ShapeConfig()
        self.state_ksp = state_ksp or DEFAULT_KSP
        self.adjoint_ksp = adjoint_ksp or DEFAULT_KSP
        self.gradient_ksp = gradient_ksp or DEFAULT_KSP
        # the reference hard-codes cg + boomeramg, rtol 1e-16, max_it 100 for mu
        # (shape_form_handler.py:113-120); the oracle uses its direct solve unless told otherwise
        self.mu_ksp = mu_ksp or DEFAULT_KSP
        self.state_bc = mesh.boundary_vertices(bc_markers)
        self.vcell_dofs = fem.vector_cell_dofs(self.cells, self.d)
        self.shape_bc_dofs = self._shape_bc_dofs()
        self.valence = quality.vertex_valence(self.cells, self.nv)
        self.old_coords = None
        self.its = {}
        self.u = np.zeros(self.nv)
        self.p = np.zeros(self.nv)
        if self.state_degree == 2:
            self.p2_dofs, self.p2_edges = p2.cell_dofs(self.cells, self.nv)
            self.n_state = self.nv + len(self.p2_edges)
            self.state_bc = p2.boundary_dofs(mesh, bc_markers, self.p2_edges)
            self.u = np.zeros(self.n_state)
            self.p = np.zeros(self.n_state)
        self.mu = np.ones(self.nv)
        self.A_elas = None
        self.current_mesh_quality = quality.compute_mesh_quality(
            self.coords, self.cells, self.cfg.qtype, self.cfg.measure, self.cfg.quantile
        )
        # targets of the geometric regularisation are fixed at construction (shape_regularization.py:141-146,334-339)
        self.target_volume = self.volume() if self.cfg.use_initial_volume else self.cfg.target_volume
        tb = list(self.cfg.target_barycenter) + [0.0] * 3
        self.target_barycenter = self.barycenter() if self.cfg.use_initial_barycenter else np.asarray(tb[:3], dtype=float)
        self.target_surface = self.surface() if self.cfg.use_initial_surface else self.cfg.target_surface
        self.reg_mu = {k: float(getattr(self.cfg, "factor_" + k)) for k in ("volume", "surface", "barycenter", "curvature")}
        if self.cfg.use_relative_scaling:
            self._scale_regularization()
        self.update_scalar_product()

    # -- geometric regularisation: volume and barycenter (shape_regularization.py)
    def volume(self):
        return float(np.sum(self._geom()[1]))

    def barycenter(self):
        """[3] (z = 0 in 2-D), ``_compute_barycenter_list`` (``:486-510``)."""
        _, vol, _ = self._geom()
        out = np.zeros(3)
        out[: self.d] = (vol[:, None] * self.coords[self.cells].mean(axis=1)).sum(axis=0) / vol.sum()
        return out

    def boundary_facets(self):
        """Facets belonging to exactly one cell (the ``ds`` measure)."""
        if getattr(self, "_bfacets", None) is None:
            d = self.d
            sub = np.concatenate([np.delete(self.cells, k, axis=1) for k in range(d + 1)], axis=0)
            uniq, counts = np.unique(np.sort(sub, axis=1), axis=0, return_counts=True)
            self._bfacets = uniq[counts == 1]
        return self._bfacets

    def _facet_measures(self):
        X = self.coords[self.boundary_facets()]  # [nf, d, d]
        if self.d == 2:
            return np.linalg.norm(X[:, 1] - X[:, 0], axis=1)
        return 0.5 * np.linalg.norm(np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), axis=1)

    def surface(self):
        """``_compute_surface`` (``shape_regularization.py:307-316``)."""
        return float(np.sum(self._facet_measures()))

    def surface_derivative_vector(self):
        """int_Gamma t_div(V, n) ds for V = phi_a e_i (``:63-74,248-269``): the gradient of the facet measures with
        respect to the vertex positions, by central differences of the exact measure formula would be an option --
        here the closed forms: segment (x1 - x0)/|F|; triangle 1/2 (n_hat x opposite edge)."""
        F = self.boundary_facets()
        X = self.coords[F]
        out = np.zeros((self.nv, self.d))
        if self.d == 2:
            e = X[:, 1] - X[:, 0]
            t = e / np.linalg.norm(e, axis=1)[:, None]
            np.add.at(out, F[:, 1], t)
            np.add.at(out, F[:, 0], -t)
        else:
            n = np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0])
            nh = n / np.linalg.norm(n, axis=1)[:, None]
            for a in range(3):
                b, c = (a + 1) % 3, (a + 2) % 3
                np.add.at(out, F[:, a], 0.5 * np.cross(X[:, b] - X[:, c], nh))
        return out.reshape(-1)

    # -- curvature regularisation (shape_regularization.py:512-646)
    def _boundary_frames(self):
        """For every boundary facet: its vertices F [nf, d], the adjacent cell's vertices C [nf, d+1], the tangential
        gradients tg [nf, d+1, d] of that cell's hat functions, ``t_grad(phi, n) = grad phi - (grad phi . n) n``
        (``:49-60``; the reference evaluates ``grad`` in the adjacent cell), the projector P = t_grad(x, n) = I - n n^T
        [nf, d, d] and the facet measures."""
        d = self.d
        F = self.boundary_facets()
        sub = np.concatenate([np.sort(np.delete(self.cells, k, axis=1), axis=1) for k in range(d + 1)], axis=0)
        owner = np.tile(np.arange(len(self.cells)), d + 1)
        key = lambda a: np.ascontiguousarray(a).view([("", a.dtype)] * d).reshape(-1)  # noqa: E731
        order = np.argsort(key(sub))
        pos = np.searchsorted(key(sub)[order], key(np.sort(F, axis=1)))
        cell = owner[order[pos]]
        C = self.cells[cell]
        _, _, G = fem.cell_geometry(self.coords, self.cells)
        G = G[cell]  # [nf, d+1, d]
        # the hat function of the vertex opposite to the facet is constant zero on it: its gradient is normal
        opp = np.array([[v not in set(f) for v in c] for c, f in zip(C, F)])
        nrm = -G[opp]
        nrm = nrm / np.linalg.norm(nrm, axis=1)[:, None]
        tg = G - np.einsum("fad,fd->fa", G, nrm)[:, :, None] * nrm[:, None, :]
        P = np.eye(d)[None] - nrm[:, :, None] * nrm[:, None, :]
        return F, C, tg, P, self._facet_measures()

    def boundary_mass_matrix(self):
        """``a_curvature = inner(trial, test) * ds`` per component (``:536-542``) on all vertices, ``ident_zeros``
        (``:601``) for the interior rows; the exact P1 facet mass matrix |F| (1 + delta_ab) / (d (d + 1))."""
        import scipy.sparse as sps

        d = self.d
        F = self.boundary_facets()
        m = self._facet_measures()
        loc = (np.ones((d, d)) + np.eye(d)) / (d * (d + 1))
        rows = np.repeat(F, d, axis=1).reshape(-1)
        cols = np.tile(F, (1, d)).reshape(-1)
        vals = (m[:, None, None] * loc[None]).reshape(-1)
        M = sps.coo_matrix((vals, (rows, cols)), shape=(self.nv, self.nv)).tocsr()
        interior = np.setdiff1d(np.arange(self.nv), np.unique(F))
        return M + sps.coo_matrix((np.ones(len(interior)), (interior, interior)), shape=(self.nv, self.nv)).tocsr()

    def compute_curvature(self):
        """``_compute_curvature`` (``:596-613``): ``M_Gamma kappa = l_curvature`` with
        ``l_curvature[V] = int inner(t_grad(x, n), t_grad(V, n)) ds`` (``:543-549``); returns kappa [nv, d]."""
        import scipy.sparse.linalg as spla

        F, C, tg, P, m = self._boundary_frames()
        d = self.d
        b = np.zeros((self.nv, d))
        # inner(P, e_i (x) tg_a) = (P tg_a)_i
        np.add.at(b, C.reshape(-1), (m[:, None, None] * np.einsum("fij,faj->fai", P, tg)).reshape(-1, d))
        M = self.boundary_mass_matrix().tocsc()
        self.kappa_curvature = np.column_stack([spla.spsolve(M, b[:, i]) for i in range(d)])
        self._boundary_mass = M
        return self.kappa_curvature

    def curvature_objective(self):
        """``1/2 mu int kappa . kappa ds`` (``:577-594``)."""
        k = self.compute_curvature()
        F = self.boundary_facets()
        interior = np.setdiff1d(np.arange(self.nv), np.unique(F))
        kb = k.copy()
        kb[interior] = 0.0
        M = self._boundary_mass
        return 0.5 * self.reg_mu["curvature"] * float(sum((M @ kb[:, i]) @ kb[:, i] for i in range(self.d)))

    def curvature_derivative_vector(self):
        """``compute_shape_derivative`` (``:554-575``) for V = phi_a e_i with the current kappa:
        ``mu int inner((I - (P + P^T)) t_grad(V), t_grad(kappa)) + 1/2 t_div(V) t_div(kappa) ds``."""
        F, C, tg, P, m = self._boundary_frames()
        d = self.d
        kap = self.kappa_curvature[C]                                # [nf, d+1, d]
        tK = np.einsum("fbi,fbj->fij", kap, tg)                      # t_grad(kappa) = sum_b kappa_b (x) tg_b
        R = np.eye(d)[None] - (P + np.transpose(P, (0, 2, 1)))
        # t_grad(V) = e_i (x) tg_a: inner(R (e_i (x) tg_a), tK) = (R^T tK tg_a)_i ; t_div(V) = tg_a[i]
        first = np.einsum("fji,fjk,fak->fai", R, tK, tg)
        second = 0.5 * np.einsum("fii->f", tK)[:, None, None] * tg
        out = np.zeros((self.nv, d))
        np.add.at(out, C.reshape(-1), (m[:, None, None] * (first + second)).reshape(-1, d))
        return self.reg_mu["curvature"] * out.reshape(-1)

    def _scale_regularization(self):
        """``scale`` of every term (``shape_regularization.py:194-211,289-305,461-484,621-643``): the factor is divided by
        the term's value on the initial geometry unless that vanishes (< 1e-15)."""
        initial = {
            "volume": 0.5 * (self.volume() - self.target_volume) ** 2,
            "surface": 0.5 * (self.surface() - self.target_surface) ** 2,
            "barycenter": 0.5 * float(np.sum((self.barycenter() - self.target_barycenter) ** 2)),
        }
        if self.reg_mu["curvature"] > 0.0:
            self.reg_mu["curvature"], keep = 1.0, self.reg_mu["curvature"]
            initial["curvature"] = self.curvature_objective()
            self.reg_mu["curvature"] = keep
        for k, v in initial.items():
            if self.reg_mu[k] > 0.0 and abs(v) >= 1e-15:
                self.reg_mu[k] /= abs(v)

    def regularization_objective(self):
        c = types.SimpleNamespace(factor_volume=self.reg_mu["volume"], factor_surface=self.reg_mu["surface"],
                                  factor_barycenter=self.reg_mu["barycenter"], factor_curvature=self.reg_mu["curvature"])
        val = 0.0
        if c.factor_surface > 0.0:
            val += 0.5 * c.factor_surface * (self.surface() - self.target_surface) ** 2
        if c.factor_volume > 0.0:
            val += 0.5 * c.factor_volume * (self.volume() - self.target_volume) ** 2
        if c.factor_barycenter > 0.0:
            val += 0.5 * c.factor_barycenter * float(np.sum((self.barycenter() - self.target_barycenter) ** 2))
        if c.factor_curvature > 0.0:
            val += self.curvature_objective()
        return val

    def regularization_derivative_elements(self):
        """Element vectors [Nc, d+1, d] of the regularisation's shape derivative for V = phi_a e_i:
        d(vol)[V] = int div V, d(int x_k)[V] = int (V_k + x_k div V); the barycenter term is the reference's
        form verbatim (``:353-410``), including its ``+ bc_k / vol * div V``."""
        c, d = self.cfg, self.d
        _, vol, G = self._geom()
        dvol = vol[:, None, None] * G                                              # |K| G_a[i]
        be = np.zeros_like(dvol)
        if self.reg_mu["volume"] > 0.0:
            be += self.reg_mu["volume"] * (self.volume() - self.target_volume) * dvol
        if self.c_volume != 0.0:
            be += self.c_volume * dvol
        if self.reg_mu["barycenter"] > 0.0:
            v, bc = self.volume(), self.barycenter()
            xbar = self.coords[self.cells].mean(axis=1)                              # [Nc, d]
            for k in range(d):
                dxk = dvol * xbar[:, k][:, None, None]                              # int_K x_k div V
                dxk[:, :, k] += (vol / (d + 1))[:, None]                            # int_K V_k
                be += self.reg_mu["barycenter"] * (bc[k] - self.target_barycenter[k]) * (bc[k] / v * dvol + dxk / v)
        return be

    # -- boundary conditions of the deformation space (shape_form_handler.py:551-588)
    def _shape_bc_dofs(self):
        d, c = self.d, self.cfg
        dofs = []
        fix = self.mesh.boundary_vertices(c.shape_bdry_fix)
        for k in range(d):
            dofs.append(d * fix + k)
        for k, mk in enumerate((c.shape_bdry_fix_x, c.shape_bdry_fix_y, c.shape_bdry_fix_z)[:d]):
            dofs.append(d * self.mesh.boundary_vertices(mk) + k)
        for k in c.fixed_dimensions:  # the whole component is removed from the scalar product
            dofs.append(d * np.arange(self.nv) + int(k))
        return np.unique(np.concatenate(dofs)).astype(np.int64) if dofs else np.zeros(0, np.int64)

    def _geom(self):
        return fem.cell_geometry(self.coords, self.cells)

    # -- state / adjoint (a3, a4)
    def solve_state(self):
        _, vol, G = self._geom()
        if self.state_degree == 2:
            Ke = p2.stiffness_elements(vol, G)
            be = p2.load_elements_callable(self.coords, self.cells, vol, self.f, self.f_degree)
            A, b = fem.assemble_system(Ke, be, self.p2_dofs, self.n_state, self.state_bc, 0.0)
        else:
            Ke = self.kappa * fem.stiffness_elements(vol, G) + self.reaction * fem.mass_elements(vol, self.d)
            be = fem.load_elements_callable(self.coords, self.cells, vol, self.f, self.f_degree + 1)
            A, b = fem.assemble_system(Ke, be, self.cells, self.nv, self.state_bc, 0.0)
        self.u, reason, its, _ = krylov.solve(A, b, self.state_ksp)
        self._check(reason)
        self.its["state"] = its
        self.A_state, self.b_state = A, b
        self._update_effective_cost_coefficients()
        return self.u

    def _integrand_values(self):
        """(int u, 1/2 int (u - u_d)^2, volume, boundary measure) for the current state and mesh."""
        _, vol, _ = self._geom()
        iu = float(np.sum(vol * self.u[self.cells].mean(axis=1)))
        it = 0.0
        if self.u_d is not None:
            w = (self.u - self.u_d)[self.cells]
            it = 0.5 * float(np.einsum("ca,cab,cb->", w, fem.mass_elements(vol, self.d), w))
        need_surf = self._base["c_surface"] != 0.0 or any(t["c_surface"] != 0.0 for t in self.scalar_tracking)
        return iu, it, float(np.sum(vol)), (self.surface() if need_surf else 0.0)

    def _tracking_values(self):
        iu, it, vol, surf = self._integrand_values()
        return [t["c_state"] * iu + t["c_track"] * it + t["c_volume"] * vol + t["c_surface"] * surf for t in self.scalar_tracking]

    def _update_effective_cost_coefficients(self):
        """c_x <- c_x of the integral terms + sum_k weight_k (I_k - goal_k) c_x,k: what the adjoint equation and the shape
        derivative see (the derivative of a scalar tracking term is its integrand's derivative times that factor)."""
        if not self.scalar_tracking:
            return
        eff = dict(self._base)
        for t, val in zip(self.scalar_tracking, self._tracking_values()):
            m = t["weight"] * (val - t["goal"])
            for k in eff:
                eff[k] += m * t[k]
        self.c_state, self.c_track, self.c_volume, self.c_surface = (eff[k] for k in ("c_state", "c_track", "c_volume", "c_surface"))

    def solve_adjoint(self):
        _, vol, G = self._geom()
        if self.state_degree == 2:
            Ke = p2.stiffness_elements(vol, G)
            be = -self.c_state * vol[:, None] * p2.reference_tensors(self.d)[2][None, :]  # -dJ/du[v] = -c int v
            A, b = fem.assemble_system(Ke, be, self.p2_dofs, self.n_state, self.state_bc, 0.0)
        else:
            Ke = self.kappa * fem.stiffness_elements(vol, G) + self.reaction * fem.mass_elements(vol, self.d)
            be = -self.c_state * fem.load_elements_p1(vol, self.cells, np.ones(self.nv), self.d)  # -dJ/du[v] = -c int v
            if self.c_track != 0.0:  # - c_track int (u - u_d) v
                be = be - self.c_track * fem.load_elements_p1(vol, self.cells, self.u - self.u_d, self.d)
            A, b = fem.assemble_system(Ke, be, self.cells, self.nv, self.state_bc, 0.0)
        self.p, reason, its, _ = krylov.solve(A, b, self.adjoint_ksp)
        self._check(reason)
        self.its["adjoint"] = its
        self.b_adjoint = b
        return self.p

    @staticmethod
    def _check(reason):
        if reason < 0:
            raise RuntimeError(f"PETScKSPError({reason})")  # cashocs/_utils/linalg.py:543-544

    def cost(self):
        """``IntegralFunctional.evaluate`` (``cashocs/_optimization/cost_functional.py:160-168``)."""
        _, vol, _ = self._geom()
        reg = self.regularization_objective()  # cost_functional.py:84-86
        b = self._base
        if b["c_volume"] != 0.0:
            reg += b["c_volume"] * self.volume()
        if b["c_surface"] != 0.0:
            reg += b["c_surface"] * self.surface()
        if self.state_degree == 2:
            return b["c_state"] * p2.integral(self.u, self.p2_dofs, vol, self.d) + reg
        J = b["c_state"] * float(np.sum(vol * self.u[self.cells].mean(axis=1)))
        if b["c_track"] != 0.0:
            w = (self.u - self.u_d)[self.cells]
            J += 0.5 * b["c_track"] * float(np.einsum("ca,cab,cb->", w, fem.mass_elements(vol, self.d), w))
        for t, val in zip(self.scalar_tracking, self._tracking_values()):  # cost_functional.py:268-283
            J += 0.5 * t["weight"] * (val - t["goal"]) ** 2
        return J + reg

    # -- Lame mu (a5)
    def compute_mu(self):
        c = self.cfg
        if c.use_distance_mu:
            self.distance = boundary_distance_eikonal(self.mesh, self.coords, self.cells, c.boundaries_dist, its=self.its)
            self.mu = distance_mu(self.distance, c.dist_min, c.dist_max, c.mu_min, c.mu_max, c.smooth_mu)
            return self.mu
        if abs(c.mu_def - c.mu_fix) / c.mu_fix > 1e-2:
            _, vol, G = self._geom()
            Ke = fem.stiffness_elements(vol, G)
            fix = self.mesh.boundary_vertices(c.shape_bdry_fix)
            deff = self.mesh.boundary_vertices(c.shape_bdry_def)
            g = np.zeros(self.nv)
            g[fix] = c.mu_fix
            g[deff] = c.mu_def  # bcs_mu: fix first, def appended -> def applied last wins
            dofs = np.unique(np.concatenate([fix, deff]))
            A, b = fem.assemble_system(Ke, np.zeros((self.cells.shape[0], self.d + 1)), self.cells,
                                       self.nv, dofs, g[dofs])
            mu, reason, its, _ = krylov.solve(A, b, self.mu_ksp)
            self._check(reason)
            self.its["mu"] = its
            if c.use_sqrt_mu:
                mu = np.sqrt(np.abs(mu))
            self.mu = mu
        else:
            self.mu = np.full(self.nv, c.mu_fix)
        return self.mu

    # -- elasticity scalar product (a6, a9)
    def update_scalar_product(self):
        self.compute_mu()
        _, vol, G = self._geom()
        Ae = fem.elasticity_elements(vol, G, self.mu, self.cells, self.cfg.lambda_lame,
                                     self.cfg.damping_factor)
        if self.cfg.inhomogeneous:
            # volumes are computed at construction and only refreshed with update_inhomogeneous (:721-735)
            if getattr(self, "_volumes", None) is None or self.cfg.update_inhomogeneous:
                self._volumes = vol / vol.max()
            Ae = Ae / (self._volumes ** self.cfg.inhomogeneous_exponent)[:, None, None]
        self._Ae_elas = Ae  # the re-extension system eliminates other dofs from the same element tensors
        A, _ = fem.assemble_system(Ae, None, self.vcell_dofs, self.d * self.nv,
                                   self.shape_bc_dofs, 0.0)
        self.A_elas = A
        return A

    def _reextension_bc_dofs(self):
        """``_setup_bcs_extension`` (shape_form_handler.py:590-611): every component on every boundary the
        [ShapeGradient] section names, plus the components removed by ``fixed_dimensions`` (:706-712)."""
        d, c = self.d, self.cfg
        markers = (tuple(c.shape_bdry_def) + tuple(c.shape_bdry_fix) + tuple(c.shape_bdry_fix_x)
                   + tuple(c.shape_bdry_fix_y) + tuple(c.shape_bdry_fix_z))
        bdry = self.mesh.boundary_vertices(markers)
        dofs = [(d * bdry[:, None] + np.arange(d)[None, :]).reshape(-1)]
        for k in c.fixed_dimensions:
            dofs.append(d * np.arange(self.nv) + int(k))
        return np.unique(np.concatenate(dofs)).astype(np.int64)

    def compute_normal_deformation(self, g):
        """``_compute_normal_deformation`` (shape_gradient_problem.py:225-299): two L2 projections on the boundaries the
        [ShapeGradient] section names -- the scalar ``phi`` with ``int phi v ds = int (g . n) v ds`` for all P1 ``v``, then
        the vector ``w`` with ``int w . v ds = int phi (n . v) ds``; rows without a boundary facet are identities
        (``ident_zeros``), so ``w`` vanishes in the interior.  ``n`` is the outward unit normal.
        A boundary facet carries a marker when all its vertices do (facet tags are not stored with the oracle's meshes)."""
        import scipy.sparse as sps
        import scipy.sparse.linalg as spla

        d, c = self.d, self.cfg
        markers = (tuple(c.shape_bdry_def) + tuple(c.shape_bdry_fix) + tuple(c.shape_bdry_fix_x)
                   + tuple(c.shape_bdry_fix_y) + tuple(c.shape_bdry_fix_z))
        tagged = np.zeros(self.nv, dtype=bool)
        tagged[self.mesh.boundary_vertices(markers)] = True
        allF, C, _, _, all_m = self._boundary_frames()
        sel = np.all(tagged[allF], axis=1)
        F, C, m = allF[sel], C[sel], all_m[sel]
        # outward unit normal: minus the (normalised) gradient of the hat function of the opposite vertex
        _, _, G = fem.cell_geometry(self.coords, C)
        opp = np.array([[v not in set(f) for v in c] for c, f in zip(C, F)])
        n = -G[opp]
        n = n / np.linalg.norm(n, axis=1)[:, None]
        loc = (np.ones((d, d)) + np.eye(d)) / (d * (d + 1))
        Me = m[:, None, None] * loc[None]                                   # facet mass matrices [nf, d, d]
        rows, cols = np.repeat(F, d, axis=1).reshape(-1), np.tile(F, (1, d)).reshape(-1)
        M = sps.coo_matrix((Me.reshape(-1), (rows, cols)), shape=(self.nv, self.nv)).tocsr()
        free = np.setdiff1d(np.arange(self.nv), np.unique(F))
        M = (M + sps.coo_matrix((np.ones(len(free)), (free, free)), shape=(self.nv, self.nv))).tocsc()
        lu = spla.splu(M)
        gF = g.reshape(-1, d)[F]                                            # [nf, d(vertex), d(component)]
        b1 = np.zeros(self.nv)
        np.add.at(b1, F.reshape(-1), np.einsum("fab,fbi,fi->fa", Me, gF, n).reshape(-1))
        phi = lu.solve(b1)
        b2 = np.zeros((self.nv, d))
        np.add.at(b2, F.reshape(-1), np.einsum("fab,fb,fi->fai", Me, phi[F], n).reshape(-1, d))
        w = np.column_stack([lu.solve(b2[:, i]) for i in range(d)])
        return w.reshape(-1)

    def reextend_gradient_from_boundaries(self, g):
        """``ShapeGradientProblem.reextend_gradient_from_boundaries`` (shape_gradient_problem.py:175-223), mode
        ``surface``: solve the homogeneous elasticity equation of the scalar product with the boundary values of the
        gradient deformation as Dirichlet data (symmetric elimination with those values: ``setup_assembler`` with
        ``bcs_extension``, shape_form_handler.py:343-367) and take the solution as the new gradient deformation."""
        if self.cfg.reextension_mode == "normal":  # :186-191: only the normal part of the boundary values is kept
            g = self.compute_normal_deformation(g)
        elif self.cfg.reextension_mode != "surface":
            raise ValueError("reextension_mode is surface or normal")
        dofs = self._reextension_bc_dofs()
        A, b = fem.assemble_system(self._Ae_elas.copy(), np.zeros(self.vcell_dofs.shape), self.vcell_dofs,
                                   self.d * self.nv, dofs, g[dofs])
        x, reason, its, _ = krylov.solve(A, b, self.gradient_ksp, block_size=self.d, coords=self.coords)
        self._check(reason)
        self.its["reextension"] = its
        x[dofs] = g[dofs]  # apply_reextension_bcs (shape_form_handler.py:843-862)
        return x

    def p_laplace_product(self, a, b):
        """``ShapeFormHandler.scalar_product`` with ``use_p_laplacian`` (``shape_form_handler.py:761-770,785-815``):
        the p-Laplace form with the gradient replaced by ``a`` and the test function by ``b`` -- nonlinear in ``a``."""
        d, c = self.d, self.cfg
        p = int(c.p_laplacian_power)
        _, vol, G = self._geom()
        ae, be = a.reshape(-1, d)[self.cells], b.reshape(-1, d)[self.cells]
        Ha = np.einsum("cai,caj->cij", ae, G)
        Hb = np.einsum("cai,caj->cij", be, G)
        s = np.einsum("cij,cij->c", Ha, Ha)
        kappa = np.ones_like(s) if p == 2 else s ** (0.5 * (p - 2))
        w = vol * self.mu[self.cells].mean(axis=1) * (c.p_laplacian_stabilization + kappa)
        M = fem.mass_elements(vol, d)
        return float(np.sum(w * np.einsum("cij,cij->c", Ha, Hb))
                     + c.damping_factor * np.einsum("cab,cai,cbi->", M, ae, be))

    def scalar_product(self, a, b):
        if self.cfg.use_p_laplacian and not self.cfg.fixed_dimensions:
            return self.p_laplace_product(a, b)
        return float((self.A_elas @ a) @ b)

    # -- shape derivative (a7)
    def shape_derivative_elements(self):
        d = self.d
        if self.state_degree == 2:
            return p2.shape_derivative_poisson(self.coords, self.cells, self.p2_dofs, self.u, self.p, self.f,
                                               self.grad_f, self.f_degree)
        _, vol, G = self._geom()
        ue, pe = self.u[self.cells], self.p[self.cells]
        gu = np.einsum("ca,cai->ci", ue, G)
        gp = np.einsum("ca,cai->ci", pe, G)
        lam, w = fem.quadrature(d, self.f_degree + 1)
        xq = np.einsum("qa,cad->cqd", lam, self.coords[self.cells])
        fq = self.f(xq)
        gfq = self.grad_f(xq)[..., :d]
        pq = np.einsum("qa,ca->cq", lam, pe)
        int_u = self.c_state * vol * ue.mean(axis=1)
        int_fp = vol * np.einsum("cq,cq,q->c", fq, pq, w, optimize=True)
        int_gf_phi_p = vol[:, None, None] * np.einsum("cqi,cq,q,qa->cai", gfq, pq, w, lam, optimize=True)
        gugp = np.einsum("ci,ci->c", gu, gp)
        Ggu = np.einsum("cai,ci->ca", G, gu)
        Ggp = np.einsum("cai,ci->ca", G, gp)
        # Lagrangian density integrated over the cell (multiplies div V = G_a[i]); SURVEY.md 9.2 generalised to
        # kappa, a reaction term and the tracking functional
        Me = fem.mass_elements(vol, d)
        dens = int_u + self.kappa * vol * gugp - int_fp
        if self.reaction != 0.0:
            dens = dens + self.reaction * np.einsum("ca,cab,cb->c", ue, Me, pe)
        be = (
            - self.kappa * vol[:, None, None] * (Ggu[:, :, None] * gp[:, None, :] + Ggp[:, :, None] * gu[:, None, :])
            - int_gf_phi_p
        )
        if self.c_track != 0.0:
            we = ue - self.u_d[self.cells]
            dens = dens + 0.5 * self.c_track * np.einsum("ca,cab,cb->c", we, Me, we)
            if self.use_pull_back:  # - c_track int (u - u_d) (d_i u_d) phi_a   (u_d is P1: its gradient is cellwise constant)
                gud = np.einsum("ca,cai->ci", self.u_d[self.cells], G)
                be = be - self.c_track * np.einsum("cab,cb->ca", Me, we)[:, :, None] * gud[:, None, :]
        be = be + G * dens[:, None, None]
        return be.reshape(self.cells.shape[0], -1)

    def shape_derivative_vector(self):
        vec = self._cell_shape_derivative_vector()
        if self.reg_mu["surface"] > 0.0 or self.c_surface != 0.0:
            coef = self.c_surface
            if self.reg_mu["surface"] > 0.0:
                coef += self.reg_mu["surface"] * (self.surface() - self.target_surface)
            extra = coef * self.surface_derivative_vector()
            extra[self.shape_bc_dofs] = 0.0
            vec = vec + extra
        if self.cfg.factor_curvature > 0.0:
            self.compute_curvature()  # update_geometric_quantities is not needed: the form holds kappa itself
            extra = self.curvature_derivative_vector()
            extra[self.shape_bc_dofs] = 0.0
            vec = vec + extra
        return vec

    def _cell_shape_derivative_vector(self):
        be = self.shape_derivative_elements()
        if self.cfg.factor_volume > 0.0 or self.cfg.factor_barycenter > 0.0 or self.c_volume != 0.0:
            be = be + self.regularization_derivative_elements().reshape(be.shape)
        if self.shape_bc_dofs.size:
            _, be = fem.apply_bcs_elementwise(None, be, self.vcell_dofs, self.shape_bc_dofs, 0.0,
                                              self.d * self.nv)
        return fem.assemble_vector(be, self.vcell_dofs, self.d * self.nv)

    # -- a8
    def compute_shape_gradient(self):
        self.solve_state()
        self.solve_adjoint()
        self.dJ = self.shape_derivative_vector()
        if self.cfg.use_p_laplacian and not self.cfg.fixed_dimensions:  # shape_gradient_problem.py:133-139
            g = self.p_laplace_projection(self.dJ)
        else:
            g, reason, its, _ = krylov.solve(self.A_elas, self.dJ, self.gradient_ksp, block_size=self.d, coords=self.coords)
            self._check(reason)
            self.its["gradient"] = its
            g[self.shape_bc_dofs] = 0.0  # apply_shape_bcs (shape_form_handler.py:825-841)
        if self.cfg.reextend_from_boundary:
            g = self.reextend_gradient_from_boundaries(g)
        self.gradient = g
        self.gradient_norm_squared = self.scalar_product(g, g)
        return g

    # -- p-Laplace projector (shape_gradient_problem.py:329-427)
    def p_laplace_forms(self, U, p):
        """Element residual [Nc, (d+1) d] and Jacobian [Nc, (d+1) d, (d+1) d] of
        ``mu (eps + (grad U : grad U)^((p-2)/2)) grad U : grad W + delta U . W`` (``:370-386``; the Jacobian is
        ``fenics.derivative`` of it, ``nonlinear_solvers/snes.py:121``)."""
        d, n = self.d, self.d + 1
        c = self.cfg
        _, vol, G = self._geom()
        Ue = U.reshape(-1, d)[self.cells]                               # [Nc, d+1, d]
        H = np.einsum("cai,caj->cij", Ue, G)                            # d_j U_i
        s = np.einsum("cij,cij->c", H, H)
        w0 = vol * self.mu[self.cells].mean(axis=1)
        kappa = np.ones_like(s) if p == 2 else s ** (0.5 * (p - 2))
        w1 = w0 * (c.p_laplacian_stabilization + kappa)
        with np.errstate(divide="ignore", invalid="ignore"):
            w2 = np.where(s > 0.0, w0 * (p - 2) * s ** (0.5 * (p - 4)), 0.0) if p != 2 else np.zeros_like(s)
        HG = np.einsum("cij,caj->cai", H, G)                            # (grad U G_a)_i
        M = fem.mass_elements(vol, d)
        eye = np.eye(d)
        Fe = w1[:, None, None] * HG + c.damping_factor * np.einsum("cab,cbi->cai", M, Ue)
        Je = np.einsum("c,cab,ij->caibj", w1, np.einsum("cak,cbk->cab", G, G), eye, optimize=True)
        Je += np.einsum("c,cai,cbj->caibj", w2, HG, HG, optimize=True)
        Je += c.damping_factor * np.einsum("cab,ij->caibj", M, eye)
        return Fe.reshape(-1, n * d), Je.reshape(-1, n * d, n * d)

    def p_laplace_projection(self, dJ, snes_rtol=1e-8, snes_atol=1e-50, snes_stol=1e-8, snes_max_it=50):
        """``_PLaplaceProjector.solve`` (``:396-427``): for p = 2 .. p_target Newton's method with full steps (SNES
        ``newtonls`` + ``basic`` line search, PETSc's default tolerances) on ``F_p(U) = dJ``, each p starting from the
        previous solution; homogeneous shape boundary conditions."""
        nd = self.d * self.nv
        U = np.zeros(nd)
        bc = self.shape_bc_dofs
        self.its["p_laplace_newton"] = []
        for p in range(2, int(self.cfg.p_laplacian_power) + 1):
            def residual(U):
                Fe, Je = self.p_laplace_forms(U, p)
                F = fem.assemble_vector(Fe, self.vcell_dofs, nd) - dJ
                F[bc] = 0.0
                return F, Je

            F, Je = residual(U)
            f0 = fnorm = np.linalg.norm(F)
            n_it = 0
            while fnorm >= snes_atol and n_it < snes_max_it:
                A, _ = fem.assemble_system(Je, None, self.vcell_dofs, nd, bc, 0.0)
                y, reason, _, _ = krylov.solve(A, F, self.gradient_ksp, block_size=self.d, coords=self.coords)
                self._check(reason)
                U = U - y
                U[bc] = 0.0
                F, Je = residual(U)
                fnorm = np.linalg.norm(F)
                n_it += 1
                if fnorm <= snes_rtol * f0 or np.linalg.norm(y) < snes_stol * np.linalg.norm(U):
                    break
            else:
                if fnorm >= snes_atol:
                    raise RuntimeError("p-Laplace Newton iteration did not converge")
            self.its["p_laplace_newton"].append(n_it)
        return U

    # -- mesh handling (a12, a13, a15, a16)
    def move_mesh(self, deformation):
        """``_MeshHandler.move_mesh``: a-priori test -> move -> intersection test -> quality."""
        V = deformation.reshape(self.nv, self.d)
        ok, _, _ = quality.a_priori_test(self.coords, self.cells, V, self.cfg.volume_change)
        if not ok:
            return False
        self.old_coords = self.coords.copy()
        self.coords += V
        success = True
        if self.cfg.test_for_intersections:
            success = self.intersection_test()
            if not success:
                self.revert_transformation()
        self.current_mesh_quality = quality.compute_mesh_quality(
            self.coords, self.cells, self.cfg.qtype, self.cfg.measure, self.cfg.quantile
        )
        return success

    def intersection_test(self):
        return quality.intersection_test(self.coords, self.cells)

    def revert_transformation(self):
        self.coords[:, :] = self.old_coords
        self.old_coords = None

    def compute_decreases(self, direction, stepsize):
        c = self.cfg
        if c.angle_change == float("inf"):
            return 0
        fro = quality.gradient_frobenius_max(self.coords, self.cells, direction.reshape(self.nv, self.d))
        return int(np.maximum(np.ceil(np.log(c.angle_change / stepsize / fro) / np.log(1 / c.beta_armijo)), 0.0))

    def update_optimization_variables(self, direction, stepsize):
        """Returns (accepted stepsize, list of decisions) -- decisions are what CUDA must reproduce."""
        decisions = []
        while True:
            if self.move_mesh(stepsize * direction):
                if self.current_mesh_quality < self.cfg.tol_lower:
                    decisions.append(("quality_reject", stepsize))
                    stepsize /= self.cfg.beta_armijo
                    self.revert_transformation()
                    continue
                decisions.append(("accept", stepsize))
                break
            decisions.append(("move_reject", stepsize))
            stepsize /= self.cfg.beta_armijo
        return stepsize, decisions

    def evaluate_cost(self):
        self.solve_state()
        return self.cost()


def armijo_step(prob: PoissonShapeProblem, direction, stepsize, J_cur, iteration=0,
                has_curvature_info=False, max_trials=60, newton_like=False, polynomial=False):
    """One Armijo line search (``armijo_line_search.py:96-202``); returns (stepsize, J_new, log of decisions).
    ``newton_like`` (L-BFGS) drops the relative stepsize criterion (``:79-83``).  ``polynomial`` replaces the
    backtracking by the quadratic / cubic interpolation of ``polynomial_line_search.py:228-335``."""
    c = prob.cfg
    log = []
    f_vals, alpha_vals = collections.deque(), collections.deque()
    ndec = prob.compute_decreases(direction, stepsize)
    stepsize /= c.beta_armijo**ndec
    if c.safeguard_stepsize and iteration == 0:
        nrm = np.sqrt(prob.scalar_product(direction, direction))
        stepsize = float(np.minimum(stepsize, 100.0 / (1.0 + nrm)))
    dinf = float(np.max(np.abs(direction)))
    step0 = getattr(prob, "_armijo_stepsize_initial", c.initial_stepsize)
    for _ in range(max_trials):
        # armijo_line_search.py:75-94: the step became too small, absolutely or relative to the first accepted one
        if stepsize * dinf <= 1e-9 or (not newton_like and stepsize / step0 <= 1e-9):
            log.append(("fail_stepsize", stepsize))
            return None, J_cur, log
        dm = prob.scalar_product(prob.gradient, direction)
        stepsize, dec = prob.update_optimization_variables(direction, stepsize)
        log.extend(dec)
        J_new = prob.evaluate_cost()
        alpha_vals.append(stepsize)
        f_vals.append(J_new)
        if getattr(prob, "is_control", False):
            # control problems measure the decrease with the actual update (armijo_line_search.py:204-223,
            # control_variable_abstractions.py:74-96): g . (u_new - u_old)
            decrease = prob.scalar_product(prob.gradient, prob.u - prob.u_temp)
        else:
            decrease = dm * stepsize
        if not hasattr(prob, "armijo_margins"):
            prob.armijo_margins = []
        prob.armijo_margins.append(J_new - (J_cur + c.epsilon_armijo * decrease))
        if J_new < J_cur + c.epsilon_armijo * decrease:
            log.append(("armijo_accept", stepsize))
            # armijo_line_search.py:169-184: a step that leaves the mesh below tol_upper asks for remeshing (a no-op
            # without `[Mesh] remesh`) and skips the update of the reference stepsize
            needs_remesh = (not getattr(prob, "is_control", False)
                            and prob.current_mesh_quality < prob.cfg.tol_upper)
            if iteration == 0 and not needs_remesh:
                prob._armijo_stepsize_initial = stepsize
            prob.update_scalar_product()  # post_line_search (line_search.py:247-249)
            return stepsize, J_new, log
        log.append(("armijo_reject", stepsize))
        if polynomial:
            stepsize = _polynomial_stepsize(c, J_cur, decrease, f_vals, alpha_vals, stepsize)
        else:
            stepsize /= c.beta_armijo
        prob.revert_transformation()
    return None, J_cur, log


def _polynomial_stepsize(c, f_current, decrease_measure, f_vals, alpha_vals, stepsize):
    """``polynomial_line_search.py:228-335``: minimiser of the quadratic (one trial value known) or cubic (two)
    model through f(0), f'(0) and the trial values, clamped to [factor_low, factor_high] * last trial."""
    if len(f_vals) == 1:
        alpha = -(decrease_measure * stepsize**2) / (2.0 * (f_vals[0] - f_current - decrease_measure * alpha_vals[0]))
    else:
        a0, a1 = alpha_vals[0], alpha_vals[1]
        coeffs = (1 / (a0**2 * a1**2 * (a1 - a0))
                  * np.array([[a0**2, -(a1**2)], [-(a0**3), a1**3]])
                  @ np.array([f_vals[1] - f_current - decrease_measure * a1, f_vals[0] - f_current - decrease_measure * a0]))
        a, b = coeffs
        alpha = float((-b + np.sqrt(b**2 - 3 * a * decrease_measure)) / (3 * a))
    alpha = min(max(alpha, c.factor_low * alpha_vals[-1]), c.factor_high * alpha_vals[-1]) if alpha == alpha else alpha
    if (c.polynomial_model == "quadratic" and len(f_vals) == 1) or (c.polynomial_model == "cubic" and len(f_vals) == 2):
        f_vals.popleft()
        alpha_vals.popleft()
    return float(alpha)


def basic_step(prob, direction, stepsize):
    """``BasicLineSearch.search`` (``basic_line_search.py:58-141``): the constant stepsize of the config, reduced
    only by the mesh tests (shape problems) or when the state equation cannot be solved."""
    log = []
    while True:
        stepsize, dec = prob.update_optimization_variables(direction, stepsize)
        log.extend(dec)
        try:
            J_new = prob.evaluate_cost()
            return stepsize, J_new, log
        except RuntimeError:
            stepsize /= prob.cfg.beta_armijo
            prob.revert_transformation()


def line_search(prob, direction, stepsize, J_cur, iteration=0, has_curvature_info=False, newton_like=False):
    """Dispatch on ``[LineSearch] method``; returns (accepted stepsize or None, trial stepsize of the next call)."""
    c = prob.cfg
    if c.line_search == "basic":
        step, _, _ = basic_step(prob, direction, c.initial_stepsize)
        prob.update_scalar_product()
        return step, c.initial_stepsize
    polynomial = c.line_search == "polynomial"
    step, _, log = armijo_step(prob, direction, stepsize, J_cur, iteration=iteration, has_curvature_info=has_curvature_info,
                               newton_like=newton_like, polynomial=polynomial)
    if not hasattr(prob, "decision_log"):
        prob.decision_log = []
    prob.decision_log.extend(log)  # every decision of the run, in order (what the device must reproduce)
    if step is None:
        return None, stepsize
    if has_curvature_info:
        return step, step
    # armijo_line_search.py:196-197 / polynomial_line_search.py:222-223: no curvature information -> enlarge
    return step, (step / c.factor_high if polynomial else step * c.beta_armijo)


def gradient_descent(prob: PoissonShapeProblem, rtol=1e-2, atol=0.0, max_iter=50):
    """``GradientDescentMethod.run`` (``optimization_algorithms/gradient_descent.py:28-60``)."""
    stepsize = prob.cfg.initial_stepsize
    hist = []
    gnorm0 = None
    for it in range(max_iter + 1):
        g = prob.compute_shape_gradient()
        gnorm = np.sqrt(prob.gradient_norm_squared)
        if gnorm0 is None:
            gnorm0 = gnorm
        J = prob.cost()
        hist.append((it, J, gnorm / gnorm0, prob.current_mesh_quality, stepsize))
        if gnorm <= atol + rtol * gnorm0:
            return True, hist
        if it == max_iter:
            break
        step, stepsize = line_search(prob, -g, stepsize, J, iteration=it)
        if step is None:
            return False, hist
    return False, hist


def lbfgs(prob: PoissonShapeProblem, rtol=1e-2, atol=0.0, max_iter=50, memory=3, scaling=True, periodic_restart=0,
          damped=False):
    """``LBFGSMethod.run`` (``optimization_algorithms/l_bfgs.py:91-135``) with the two-loop recursion, the
    periodic restart of ``check_restart`` (``:312-334``) and Powell's damping (``:284-304``)."""
    S, Y, RHO = [], [], []
    periodic_its = 0
    stepsize = prob.cfg.initial_stepsize
    hist = []
    g = prob.compute_shape_gradient()
    gnorm = gnorm0 = np.sqrt(prob.gradient_norm_squared)
    it = 0
    while True:
        J = prob.cost()
        hist.append((it, J, gnorm / gnorm0, prob.current_mesh_quality, stepsize))
        if gnorm <= atol + rtol * gnorm0:
            return True, hist
        if it >= max_iter:
            return False, hist
        # two-loop; with box constraints the recursion acts on the inactive set and the active part of the
        # gradient is added back (l_bfgs.py:182-222)
        box = getattr(prob, "has_box_constraints", False)
        if box:
            prob.compute_active_sets()
        inactive = prob.restrict_to_inactive_set if box else (lambda a: a)
        q = g.copy()
        if S:
            q = inactive(q)
        alphas = []
        for s, y, rho in zip(reversed(S), reversed(Y), reversed(RHO)):
            a = rho * prob.scalar_product(s, q)
            alphas.append(a)
            q = q - a * y
        if scaling and S:
            q = q * (prob.scalar_product(Y[-1], S[-1]) / prob.scalar_product(Y[-1], Y[-1]))
        if S:
            q = inactive(q)
        for (s, y, rho), a in zip(zip(S, Y, RHO), reversed(alphas)):
            b = rho * prob.scalar_product(y, q)
            q = q + (a - b) * s
        if S and box:
            q = inactive(q) + prob.restrict_to_active_set(g)
        d = -q
        has_curv = len(S) > 0
        restarted = False
        if periodic_restart > 0:
            if periodic_its < periodic_restart:
                periodic_its += 1
            else:
                d = -g
                periodic_its = 0
                has_curv = False
                restarted = True
                S, Y, RHO = [], [], []
        if prob.scalar_product(g, d) >= 0:  # check_for_ascent (optimization_algorithm.py:315-331)
            d = -g
            S, Y, RHO = [], [], []
            has_curv = False
        trial = 1.0 if has_curv else stepsize
        if restarted:  # line_search.py:176-181: a restarted solver starts from the configured stepsize again
            trial = prob.cfg.initial_stepsize
        step, stepsize = line_search(prob, d, trial, J, iteration=it, has_curvature_info=has_curv, newton_like=True)
        if step is None:
            return False, hist
        g_prev = g
        g = prob.compute_shape_gradient()
        gnorm = np.sqrt(prob.gradient_norm_squared)
        it += 1
        if memory > 0 and not gnorm <= atol + rtol * gnorm0:
            if box:
                prob.compute_active_sets()
            y = inactive(g - g_prev)
            s = inactive(step * d)
            curv = prob.scalar_product(y, s)
            if damped:
                # H y with the current memory (l_bfgs.py:331-354), then s blended with it so that y . r > 0
                hy = y.copy()
                al = []
                for s_, y_, rho in zip(reversed(S), reversed(Y), reversed(RHO)):
                    a = rho * prob.scalar_product(s_, hy)
                    al.append(a)
                    hy = hy - a * y_
                if scaling and S:
                    hy = hy * (prob.scalar_product(Y[-1], S[-1]) / prob.scalar_product(Y[-1], Y[-1]))
                for (s_, y_, rho), a in zip(zip(S, Y, RHO), reversed(al)):
                    hy = hy + (a - rho * prob.scalar_product(y_, hy)) * s_
                gamma = prob.scalar_product(y, hy)
                phi = 1.0 if curv >= 0.2 * gamma else (0.8 * gamma) / (gamma - curv)
                r = phi * s + (1.0 - phi) * hy
                S.append(r), Y.append(y), RHO.append(1.0 / prob.scalar_product(y, r))
            elif curv <= 0:
                S, Y, RHO = [], [], []
            else:
                S.append(s), Y.append(y), RHO.append(1.0 / curv)
            if len(S) > memory:
                S.pop(0), Y.pop(0), RHO.pop(0)


def ncg(prob: PoissonShapeProblem, method="DY", rtol=1e-2, atol=0.0, max_iter=50, periodic_restart=False,
        periodic_its=10, relative_restart=False, restart_tol=0.25):
    """``NonlinearCGMethod.run`` (``optimization_algorithms/ncg.py:71-103``): beta of Fletcher-Reeves,
    Polak-Ribiere, Hestenes-Stiefel, Dai-Yuan or Hager-Zhang (``:105-169``), restarts (``:188-212``),
    ascent check (``optimization_algorithm.py:315-331``); no curvature information for the line search."""
    sp = prob.scalar_product
    stepsize = prob.cfg.initial_stepsize
    hist = []
    g = np.zeros(prob.nv if getattr(prob, "is_control", False) else prob.d * prob.nv)
    d = np.zeros_like(g)
    memory = 0
    gnorm0 = None
    for it in range(max_iter + 1):
        g_prev = g
        g = prob.compute_shape_gradient()
        gnorm = np.sqrt(prob.gradient_norm_squared)
        if gnorm0 is None:
            gnorm0 = gnorm
        beta = 0.0
        if it > 0:
            y = g - g_prev
            m = method.casefold()
            if m == "fr":
                beta = sp(g, g) / sp(g_prev, g_prev)
            elif m == "pr":
                beta = sp(g, y) / sp(g_prev, g_prev)
            elif m == "hs":
                beta = sp(g, y) / sp(y, d)
            elif m == "dy":
                beta = sp(g, g) / sp(d, y)
            elif m == "hz":
                dy = sp(d, y)
                y2 = sp(y, y)
                beta = sp(y - 2.0 * y2 / dy * d, g) / dy
            else:
                raise ValueError(method)
        J = prob.cost()
        hist.append((it, J, gnorm / gnorm0, prob.current_mesh_quality, stepsize))
        if gnorm <= atol + rtol * gnorm0:
            return True, hist
        if it == max_iter:
            break
        d = -g + beta * d
        if periodic_restart:
            if memory < periodic_its:
                memory += 1
            else:
                d = -g
                memory = 0
        if relative_restart and abs(sp(g, g_prev)) / gnorm**2 >= restart_tol:
            d = -g
        if getattr(prob, "has_box_constraints", False):
            d = prob.project_ncg_search_direction(d)
        if sp(g, d) >= 0:
            d = -g
        step, stepsize = line_search(prob, d, stepsize, J, iteration=it)
        if step is None:
            return False, hist
    return False, hist


def boundary_distance_eikonal(mesh, coords, cells, boundary_idcs=(), tol=1e-1, max_iter=10, ksp=None, its=None):
    """``compute_boundary_distance_eikonal`` (``cashocs/geometry/boundary_distance.py:191-282``): ``-Laplace u = 1``,
    then sweeps ``int grad u . grad v = int grad u_prev / |grad u_prev| . grad v`` with ``u = 0`` on the boundaries
    ``boundary_idcs`` (all boundary vertices if empty) until ``sqrt(int (|grad u| - 1)^2)`` fell by ``tol``."""
    d, nv = coords.shape[1], coords.shape[0]
    if len(boundary_idcs) > 0:
        bc = mesh.boundary_vertices(tuple(boundary_idcs))
    else:
        sub = np.concatenate([np.delete(cells, k, axis=1) for k in range(d + 1)], axis=0)
        uniq, counts = np.unique(np.sort(sub, axis=1), axis=0, return_counts=True)
        bc = np.unique(uniq[counts == 1])
    _, vol, G = fem.cell_geometry(coords, cells)
    Ke = fem.stiffness_elements(vol, G)
    ksp = ksp or DEFAULT_KSP

    def solve(be):
        A, b = fem.assemble_system(Ke, be, cells, nv, bc, 0.0)
        u, reason, n_its, _ = krylov.solve(A, b, ksp)
        PoissonShapeProblem._check(reason)
        if its is not None:
            its["distance"] = its.get("distance", 0) + n_its
        return u

    def grad(u):
        return np.einsum("ca,cai->ci", u[cells], G)

    def residual(u):
        return np.sqrt(np.sum(vol * (np.linalg.norm(grad(u), axis=1) - 1.0) ** 2))

    u = solve(fem.load_elements_p1(vol, cells, np.ones(nv), d))
    res_0 = residual(u)
    for _ in range(max_iter):
        gu = grad(u)
        # cells with all vertices on the boundary have grad u = 0: their 0 / 0 only reaches boundary rows, which the
        # Dirichlet elimination overwrites (as in the reference)
        with np.errstate(divide="ignore", invalid="ignore"):
            direction = gu / np.linalg.norm(gu, axis=1)[:, None]
        u = solve(vol[:, None] * np.einsum("ci,cai->ca", direction, G))
        if residual(u) <= res_0 * tol:
            break
    return u


def distance_mu(dist, dist_min, dist_max, mu_min, mu_max, smooth=False):
    """The ``mu_expression`` of ``Stiffness._setup_mu_computation`` (``shape_form_handler.py:150-186``) interpolated
    to CG1, i.e. evaluated at the vertices."""
    w, dm = dist_max - dist_min, mu_max - mu_min
    with np.errstate(divide="ignore", invalid="ignore"):
        if not smooth:
            mid = mu_min + (dist - dist_min) / w * dm
        else:
            mid = (mu_min + dm / w * (dist - dist_min) - dm / w**2 * (dist - dist_min) * (dist - dist_max)
                   - 2.0 * dm / w**3 * (dist - dist_min) * (dist - dist_max) ** 2)
    return np.where(dist <= dist_min, mu_min, np.where(dist <= dist_max, mid, mu_max))


def truncated_newton(prob, inner="cg", rtol=1e-2, atol=0.0, max_iter=50, inner_rtol=1e-7, inner_atol=0.0, inner_max_it=20):
    """``NewtonMethod.run`` (``optimization_algorithms/newton.py:57-101``) with the inner Krylov solvers of
    ``HessianProblem`` (``_pde_problems/hessian_problems.py:237-431``: ``newton_solve``, ``cg``, ``cr`` on the reduced
    Hessian of ``reduced_hessian_application``: identity on the active set, ``H`` restricted to the inactive set).
    Defaults of the inner solver = ``tests/config_ocp.ini:24-28``.  Returns (converged, history)."""
    sp = prob.scalar_product
    act, inact = prob.restrict_to_active_set, prob.restrict_to_inactive_set

    def reduced_hessian(h):
        """:237-265."""
        return act(h) + inact(prob.hessian_application(inact(h)))

    def split_product(a, b):
        """The reference evaluates products of the CG / CR recurrences set by set (:304-310, :357-366)."""
        return sp(act(a), act(b)) + sp(inact(a), inact(b))

    def newton_step(g):
        prob.compute_active_sets()
        delta = np.zeros_like(g)
        r = -g
        d = r.copy()
        eps0 = np.sqrt(sp(r, r))
        if inner == "cg":  # :282-343
            rs_old = sp(r, r)
            for _ in range(inner_max_it):
                q = reduced_hessian(d)
                # <p_A, p_A> + <p_I, q_I> (:304-310)
                alpha = rs_old / (sp(act(d), act(d)) + sp(inact(d), inact(q)))
                delta = delta + alpha * d
                r = r - alpha * q
                rs_new = sp(r, r)
                if np.sqrt(rs_new) < inner_atol + inner_rtol * eps0:
                    break
                d = r + (rs_new / rs_old) * d
                rs_old = rs_new
            return delta
        s = reduced_hessian(r)  # conjugate residuals, :345-431
        q = s.copy()
        rar = split_product(r, s)
        for i in range(inner_max_it):
            alpha = rar / split_product(q, q)
            delta = delta + alpha * d
            r = r - alpha * q
            if np.sqrt(sp(r, r)) < inner_atol + inner_rtol * eps0 or i == inner_max_it - 1:
                break
            s = reduced_hessian(r)
            rar_new = split_product(r, s)
            beta = rar_new / rar
            d = r + beta * d
            q = s + beta * q
            rar = rar_new
        return delta

    hist = []
    gnorm0 = None
    stepsize = 1.0  # line_search.py:92-93
    for it in range(max_iter + 1):
        g = prob.compute_gradient()
        gnorm = np.sqrt(prob.gradient_norm_squared)
        if gnorm0 is None:
            gnorm0 = gnorm
        if gnorm <= atol + rtol * gnorm0:
            hist.append((it, prob.cost(), gnorm / gnorm0, None, stepsize))
            return True, hist
        if it == max_iter:
            hist.append((it, prob.cost(), gnorm / gnorm0, None, stepsize))
            break
        d = newton_step(g)
        has_curv = True
        if sp(g, d) >= 0:  # check_for_ascent (optimization_algorithm.py:315-331)
            d, has_curv = -g, False
        J = prob.cost()
        hist.append((it, J, gnorm / gnorm0, None, stepsize))
        step, nxt = line_search(prob, d, 1.0 if has_curv else stepsize, J, iteration=it, has_curvature_info=has_curv,
                                newton_like=True)
        if step is None and has_curv:  # newton.py:80-92: fall back to the negative gradient once
            step, nxt = line_search(prob, -g, nxt, J, iteration=it, has_curvature_info=False, newton_like=True)
        if step is None:
            return False, hist
        stepsize = nxt
    return False, hist


def control_gradient_test(prob, rng=None, h=None):
    """Taylor test of ``cashocs/verification.py:38-179`` for control problems; returns the last convergence rate."""
    rng = rng or np.random.RandomState(300696)
    u0 = prob.u.copy()
    if h is None:
        h = rng.rand(prob.nv)
    prob.solve_state()
    J0 = prob.cost()
    g = prob.compute_gradient()
    dJh = prob.scalar_product(g, h)
    eps_list = [1e-2 / 2**i for i in range(4)]
    res = []
    for eps in eps_list:
        prob.u = u0 + eps * h
        prob.solve_state()
        res.append(abs(prob.cost() - J0 - eps * dJh))
    prob.u = u0
    rates = [np.log(res[i] / res[i - 1]) / np.log(eps_list[i] / eps_list[i - 1]) for i in range(1, 4)]
    return rates[-1], rates, res


def shape_gradient_test(prob: PoissonShapeProblem, rng=None, h=None):
    """Taylor test of ``cashocs/verification.py:182-302``; returns the last convergence rate."""
    rng = rng or np.random.RandomState(300696)
    if h is None:
        h = rng.rand(prob.d * prob.nv)
    h = h.copy()
    h[prob.shape_bc_dofs] = 0.0
    J0 = prob.evaluate_cost()
    g = prob.compute_shape_gradient()
    dJh = prob.scalar_product(g, h)
    length = float(prob.coords.max() - prob.coords.min())
    eps_list = [length * 1e-4 / 2**i for i in range(4)]
    res = []
    for eps in eps_list:
        if prob.move_mesh(eps * h):
            J1 = prob.evaluate_cost()
            res.append(abs(J1 - J0 - eps * dJh))
            prob.revert_transformation()
        else:
            res.append(float("inf"))
    rates = [np.log(res[i] / res[i - 1]) / np.log(eps_list[i] / eps_list[i - 1]) for i in range(1, 4)]
    return rates[-1], rates, res


# ----------------------------------------------------------------------------- optimal control (a10)
class PoissonControlProblem:
    """F = grad y . grad p - u p;  J = 1/2 |y - y_d|^2 + alpha/2 |u|^2  (tests/test_pde_problems.py:59-84)."""

    def __init__(self, mesh, y_d, u, alpha=1e-6, bc_markers=(1, 2), state_ksp=None, adjoint_ksp=None,
                 riesz_ksp=None, control_constraints=None, control_bcs=None, cubic=0.0):
        self.mesh, self.coords, self.cells = mesh, mesh.coords, mesh.cells
        # semilinear state equation grad y . grad p + cubic y^3 p - u p (tests/test_optimal_control.py:364-376),
        # solved by Newton's method (oracle/nonlinear.py) when cubic != 0
        self.cubic = float(cubic)
        # Dirichlet conditions on the control (optimal_control_problem.py:212-230,306-311): (markers, value); the
        # Riesz problem gets the homogenised conditions, the control itself the values (set by the solve loops)
        self.control_bc_dofs = np.zeros(0, dtype=np.int64)
        if control_bcs is not None:
            markers, value = control_bcs
            self.control_bc_dofs = mesh.boundary_vertices(markers)
            u = np.array(u, dtype=float, copy=True)
            u[self.control_bc_dofs] = value
        self.d, self.nv = mesh.dim, mesh.num_vertices
        self.y_d, self.u, self.alpha = y_d, u, alpha
        # box constraints u_a <= u <= u_b (optimal_control/box_constraints.py:29-182); floats or nodal arrays
        lo, hi = control_constraints if control_constraints is not None else (-np.inf, np.inf)
        self.u_a = np.broadcast_to(np.asarray(lo, dtype=float), (self.nv,)).copy()
        self.u_b = np.broadcast_to(np.asarray(hi, dtype=float), (self.nv,)).copy()
        if not np.all(self.u_a < self.u_b):
            raise ValueError("The lower bound must always be smaller than the upper bound for the control_constraints.")
        self.has_box_constraints = not (self.u_a.max() == -np.inf and self.u_b.min() == np.inf)
        self.bc = mesh.boundary_vertices(bc_markers)
        self.state_ksp = state_ksp or DEFAULT_KSP
        self.adjoint_ksp = adjoint_ksp or DEFAULT_KSP
        self.riesz_ksp = riesz_ksp or DEFAULT_KSP
        self.its = {}

    def _system(self, rhs_vertex, sign):
        _, vol, G = fem.cell_geometry(self.coords, self.cells)
        Ke = fem.stiffness_elements(vol, G)
        be = sign * fem.load_elements_p1(vol, self.cells, rhs_vertex, self.d)
        return fem.assemble_system(Ke, be, self.cells, self.nv, self.bc, 0.0)

    def solve_state(self):
        if self.cubic != 0.0:
            from oracle import nonlinear

            _, vol, _ = fem.cell_geometry(self.coords, self.cells)
            load = fem.load_elements_p1(vol, self.cells, self.u, self.d)
            self.y, self.its["newton"] = nonlinear.newton_solve(self.coords, self.cells, load, self.cubic, self.bc)
            return self.y
        A, b = self._system(self.u, 1.0)
        self.y, reason, self.its["state"], _ = krylov.solve(A, b, self.state_ksp)
        PoissonShapeProblem._check(reason)
        return self.y

    def solve_adjoint(self):
        if self.cubic != 0.0:  # linearised operator K + 3 c M[y^2] (self-adjoint)
            from oracle import nonlinear

            _, vol, G = fem.cell_geometry(self.coords, self.cells)
            Ke = fem.stiffness_elements(vol, G) + nonlinear.weighted_mass_elements(vol, self.cells, self.y, self.d, self.cubic)
            be = -fem.load_elements_p1(vol, self.cells, self.y - self.y_d, self.d)
            A, b = fem.assemble_system(Ke, be, self.cells, self.nv, self.bc, 0.0)
            self.p, reason, self.its["adjoint"], _ = krylov.solve(A, b, self.adjoint_ksp)
            PoissonShapeProblem._check(reason)
            return self.p
        A, b = self._system(self.y - self.y_d, -1.0)
        self.p, reason, self.its["adjoint"], _ = krylov.solve(A, b, self.adjoint_ksp)
        PoissonShapeProblem._check(reason)
        return self.p

    def mass_matrix(self, riesz=True):
        """L2 mass matrix; ``riesz``: with the homogenised control BCs eliminated symmetrically
        (``control_form_handler.py:82-137``) -- the Riesz matrix and the scalar product of the optimisation loop."""
        _, vol, _ = fem.cell_geometry(self.coords, self.cells)
        if riesz and self.control_bc_dofs.size:
            M, _ = fem.assemble_system(fem.mass_elements(vol, self.d), None, self.cells, self.nv, self.control_bc_dofs, 0.0)
        else:
            M, _ = fem.assemble_system(fem.mass_elements(vol, self.d), None, self.cells, self.nv)
        return M

    def compute_gradient(self):
        """M g = alpha M u - M p  (``control_form_handler.py:209-220``)."""
        self.solve_state()
        self.solve_adjoint()
        _, vol, _ = fem.cell_geometry(self.coords, self.cells)
        be = fem.load_elements_p1(vol, self.cells, self.alpha * self.u - self.p, self.d)
        if self.control_bc_dofs.size:
            _, be = fem.apply_bcs_elementwise(None, be, self.cells, self.control_bc_dofs, 0.0, self.nv)
        b = fem.assemble_vector(be, self.cells, self.nv)
        self.M = self.mass_matrix()
        self.g, reason, self.its["gradient"], _ = krylov.solve(self.M, b, self.riesz_ksp)
        PoissonShapeProblem._check(reason)
        # stationary measure |u - P(u - g)|^2 (control_variable_abstractions.py:163-191); = |g|^2 without constraints
        w = self.u - self.project(self.u - self.g) if self.has_box_constraints else self.g
        self.gradient_norm_squared = float((self.M @ w) @ w)
        return self.g

    def hessian_application(self, h):
        """``HessianProblem.hessian_application`` (``_pde_problems/hessian_problems.py:149-235``) with the forms of
        ``HessianFormHandler.compute_newton_forms`` (``_forms/control_form_handler.py:281-360,425-542``) worked out for
        this Lagrangian: sensitivity ``K y' = M h``, adjoint sensitivity ``K p' = w_1 = M y'`` (L_yy = mass form, L_uy = 0),
        ``hessian_rhs = w_2 + w_3 = alpha M h + M p'`` (L_uu = alpha * mass form, L_yu = 0, w_3 = -d/du[-u p']), then the
        Riesz solve with the control's scalar product -- all with homogenised state / control conditions."""
        if self.cubic != 0.0:
            raise NotImplementedError("second-order forms of the semilinear state equation are not restated")
        A, b = self._system(np.asarray(h, dtype=float), 1.0)
        y_prime, reason, _, _ = krylov.solve(A, b, self.state_ksp)
        PoissonShapeProblem._check(reason)
        A, b = self._system(y_prime, 1.0)
        p_prime, reason, _, _ = krylov.solve(A, b, self.adjoint_ksp)
        PoissonShapeProblem._check(reason)
        _, vol, _ = fem.cell_geometry(self.coords, self.cells)
        be = fem.load_elements_p1(vol, self.cells, self.alpha * h + p_prime, self.d)
        if self.control_bc_dofs.size:
            _, be = fem.apply_bcs_elementwise(None, be, self.cells, self.control_bc_dofs, 0.0, self.nv)
        rhs = fem.assemble_vector(be, self.cells, self.nv)
        if not hasattr(self, "M"):
            self.M = self.mass_matrix()
        out, reason, _, _ = krylov.solve(self.M, rhs, self.riesz_ksp)
        PoissonShapeProblem._check(reason)
        self.no_sensitivity_solves = getattr(self, "no_sensitivity_solves", 0) + 2
        return out

    # -- box constraints (optimal_control/box_constraints.py:184-346)
    def project(self, a):
        return np.maximum(np.minimum(self.u_b, a), self.u_a)

    def compute_active_sets(self):
        self.idx_active = np.flatnonzero((self.u <= self.u_a) | (self.u >= self.u_b))
        self.idx_inactive = np.flatnonzero((self.u > self.u_a) & (self.u < self.u_b))

    def restrict_to_inactive_set(self, a):
        if not self.has_box_constraints:
            return a
        b = np.zeros_like(a)
        b[self.idx_inactive] = a[self.idx_inactive]
        return b

    def restrict_to_active_set(self, a):
        b = np.zeros_like(a)
        if self.has_box_constraints:
            b[self.idx_active] = a[self.idx_active]
        return b

    def project_ncg_search_direction(self, d):
        """``control_variable_abstractions.py:212-243``: no step out of the admissible set."""
        d = d.copy()
        d[((self.u <= self.u_a) & (d < 0.0)) | ((self.u >= self.u_b) & (d > 0.0))] = 0.0
        return d

    def cost(self):
        M = self.mass_matrix(riesz=False)
        e = self.y - self.y_d
        return 0.5 * float((M @ e) @ e) + 0.5 * self.alpha * float((M @ self.u) @ self.u)

    # -- duck-typed surface of the optimisation loops (gradient_descent / lbfgs / ncg / armijo_step):
    #    ControlVariableAbstractions (control_variable_abstractions.py:36-150) without box constraints
    is_control = True
    current_mesh_quality = None

    @property
    def cfg(self):
        if not hasattr(self, "_cfg"):
            self._cfg = ShapeConfig()
        return self._cfg

    @property
    def gradient(self):
        return self.g

    @property
    def nv_dofs(self):
        return self.nv

    def compute_shape_gradient(self):
        return self.compute_gradient()

    def scalar_product(self, a, b):
        if not hasattr(self, "M"):
            self.M = self.mass_matrix()
        return float((self.M @ a) @ b)

    def compute_decreases(self, direction, stepsize):
        return 0

    def update_optimization_variables(self, direction, stepsize):
        self.u_temp = self.u.copy()
        self.u = self.project(self.u_temp + stepsize * direction)
        return stepsize, [("accept", stepsize)]

    def revert_transformation(self):
        self.u = self.u_temp.copy()

    def update_scalar_product(self):
        return None

    def evaluate_cost(self):
        self.solve_state()
        return self.cost()
