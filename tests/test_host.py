"""CPU: host-side logic of the product (no compute calls): C-ABI exports, option parsing,
config handling, polynomial expressions, mesh generators, exceptions."""

import ctypes
import os
import re

import numpy as np
import pytest

import cashocs_b200 as cb
from cashocs_b200 import _lib
from cashocs_b200._forms import expressions
from cashocs_b200._utils import linalg
from oracle import fem, meshes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "cashocs_b200.h")).read()
    declared = set(re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().cb_version() >= 100
    assert _lib.load().cb_reduce_scratch_len() > 0


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.cb_poly) == 8 + 8 * 24 + 3 * 24
    assert ctypes.sizeof(_lib.cb_quad) == 8 + 8 * 4 * 64 + 8 * 64
    assert ctypes.sizeof(_lib.cb_mat) == 64
    assert ctypes.sizeof(_lib.cb_mesh_view) == 136
    assert ctypes.sizeof(_lib.cb_ksp_opts) == 104
    assert ctypes.sizeof(_lib.cb_amg_level) == 296
    assert ctypes.sizeof(_lib.cb_amg) == 56
    assert ctypes.sizeof(_lib.cb_ksp_result) == 40
    # ... and the ctypes mirrors agree with what the compiled library itself reports
    lib = _lib.load()
    mirrors = [_lib.cb_poly, _lib.cb_quad, _lib.cb_mat, _lib.cb_mesh_view, _lib.cb_ksp_opts, _lib.cb_amg_level, _lib.cb_amg,
               _lib.cb_ksp_result, _lib.cb_p2_view]
    for which, mirror in enumerate(mirrors):
        assert lib.cb_struct_size(which) == ctypes.sizeof(mirror), (which, mirror.__name__)
    assert lib.cb_struct_size(99) == -1


def test_host_colouring_through_c_abi():
    m = meshes.regular_mesh(4, 1.0, 1.0, 1.0)
    colour = np.empty(m.num_cells, dtype=np.int32)
    ncol = _lib.load().cb_colour_cells_host(m.num_vertices, m.num_cells, 4, m.cells.ctypes.data, colour.ctypes.data)
    assert 24 <= ncol <= 128
    for c in range(ncol):
        v = m.cells[colour == c].reshape(-1)
        assert np.unique(v).size == v.size


def test_no_cpu_fallback():
    m = cb.regular_mesh(3, device="cpu")
    with pytest.raises(cb.CashocsException, match="GPU only"):
        m.build_pattern()
    with pytest.raises(cb.CashocsException, match="GPU only"):
        cb.compute_mesh_quality(m)


def test_ksp_option_parsing():
    o = linalg.parse_ksp_options({"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-9, "ksp_max_it": 77})
    assert (o.ksp_type, o.pc_type, o.rtol, o.max_it, o.atol, o.dtol) == (0, 1, 1e-9, 77, 1e-50, 1e5)
    o = linalg.parse_ksp_options({"ksp_type": "cg", "pc_type": "jacobi"}, rtol=1e-3, atol=1e-7)
    assert (o.rtol, o.atol) == (1e-3, 1e-7)  # setTolerances overrides (linalg.py:533)
    o = linalg.parse_ksp_options(None)  # direct solver of the reference -> documented CG mapping
    assert (o.ksp_type, o.pc_type, o.rtol) == (0, 1, 1e-13)
    o = linalg.parse_ksp_options(linalg.iterative_ksp_options)  # cg + boomeramg -> device AMG (one GPU)
    assert (o.ksp_type, o.pc_type) == (0, 3) and o.rtol == 1e-14
    o = linalg.parse_ksp_options(linalg.iterative_ksp_options, distributed=True)  # partitioned mesh: distributed AMG
    assert (o.ksp_type, o.pc_type) == (0, 3) and o.rtol == 1e-14
    o = linalg.parse_ksp_options({"ksp_type": "minres", "pc_type": "none", "ksp_monitor": None})
    assert (o.ksp_type, o.pc_type, o.monitor) == (1, 0, 1)
    o = linalg.parse_ksp_options({"ksp_type": "cg", "pc_type": "ksp", "ksp_ksp_type": "chebyshev", "ksp_ksp_max_it": 4, "ksp_pc_type": "jacobi"})
    assert (o.pc_type, o.cheb_degree) == (2, 4)
    o = linalg.parse_ksp_options({"ksp_type": "preonly", "pc_type": "jacobi", "pc_jacobi_type": "diagonal", "ksp_rtol": 1e-16,
                                  "ksp_atol": 1e-20, "ksp_max_it": 1000})  # mesh_testing.py:74-81
    assert (o.ksp_type, o.pc_type) == (2, 1)
    for bad in ({"ksp_type": "gmres"}, {"ksp_type": "cg", "pc_type": "ilu"}, {"ksp_type": "cg", "pc_type": "jacobi", "foo": 1}):
        with pytest.raises(cb.InputError):
            linalg.parse_ksp_options(bad)


def test_config_defaults_and_validation(tmp_path):
    cfg = cb.Config()
    assert cfg.getfloat("MeshQuality", "tol_upper") == 1e-15 and cfg.get("MeshQuality", "measure") == "skewness"
    assert cfg.getfloat("LineSearch", "beta_armijo") == 2.0 and cfg.getfloat("LineSearch", "epsilon_armijo") == 1e-4
    assert cfg.getfloat("MeshQuality", "volume_change") == float("inf")
    assert cfg.getboolean("ShapeGradient", "test_for_intersections")
    cfg.validate_config()
    path = tmp_path / "config.ini"
    path.write_text("[ShapeGradient]\nshape_bdry_def = [1, 2]\nlambda_lame = 1.5\n[MeshQuality]\ntol_lower = 0.01\ntol_upper = 0.02\n")
    cfg = cb.load_config(str(path))
    assert cfg.getlist("ShapeGradient", "shape_bdry_def") == [1, 2] and cfg.getfloat("ShapeGradient", "lambda_lame") == 1.5
    cfg.validate_config()
    cfg.set("MeshQuality", "tol_lower", "0.5")
    cfg.set("LineSearch", "beta_armijo", "0.5")
    cfg.set("MeshQuality", "measure", "bogus")
    cfg.set("StateSystem", "is_linear", "maybe")
    with pytest.raises(cb.ConfigError) as e:
        cfg.validate_config()
    assert len(e.value.config_errors) >= 4
    with pytest.raises(cb.InputError):
        cb.load_config(str(tmp_path / "missing.ini"))


def test_polynomial_algebra_and_quadrature():
    x = cb.spatial_coordinate(3)
    f = 2.5 * (x[0] + 0.4 - x[1] ** 2) ** 2 + x[0] ** 2 + x[1] ** 2 - 1
    assert f.degree == 4
    pts = np.random.RandomState(0).rand(50, 3)
    X, Y = pts[:, 0], pts[:, 1]
    assert np.allclose(f(pts), 2.5 * (X + 0.4 - Y**2) ** 2 + X**2 + Y**2 - 1, rtol=1e-14)
    assert np.allclose(f.derivative(0)(pts), 5 * (X + 0.4 - Y**2) + 2 * X, rtol=1e-13)
    assert np.allclose(f.derivative(1)(pts), -10 * Y * (X + 0.4 - Y**2) + 2 * Y, rtol=1e-13)
    assert f.derivative(2).terms == {}
    s = f.struct()
    assert s.nterms == len(f.terms) <= _lib.POLY_MAXT
    for d in (2, 3):
        lam, w = expressions.simplex_quadrature(d, 5)
        lam0, w0 = fem.quadrature(d, 5)
        assert abs(w.sum() - 1) < 1e-15 and len(w) <= len(w0)  # symmetric 7- / 14-point rules vs the collapsed 3^d rule
        # same integrals as the oracle's independent rule
        g = lambda l: (l[:, 1] ** 2 * l[:, 2] ** 3 + l[:, 0])  # noqa: E731
        assert abs(np.sum(w * g(lam)) - np.sum(w0 * g(lam0))) < 1e-15


def test_host_mesh_generators_match_oracle():
    for args in ((5,), (3, 1.0, 2.0), (3, 1.0, 1.0, 1.0), (2, 1.0, 2.0, 3.0)):
        dm, om = cb.regular_mesh(*args, device="cpu"), meshes.regular_mesh(*args)
        assert np.array_equal(dm.cells.numpy(), om.cells) and np.array_equal(dm.coords.numpy(), om.coords)
        for marker in np.unique(om.facet_tags):
            assert np.array_equal(dm.boundary_vertices([int(marker)]).numpy(), om.boundary_vertices([marker]))
    with pytest.raises(cb.InputError):
        cb.regular_mesh(0, device="cpu")
    with pytest.raises(cb.InputError):
        cb.regular_box_mesh(3, start_z=0.0, device="cpu")


def test_exceptions_mirror_reference():
    e = cb.PETScKSPError(-3)
    assert e.error_code == -3 and "maximum iterations" in str(e)
    assert issubclass(cb.PETScKSPError, cb.CashocsException)
    assert "faulty input" in str(cb.InputError("obj", "param", "msg"))
    assert str(cb.NotConvergedError("solver")).startswith("The solver failed")


def test_bench_reference_arm_runs(tmp_path):
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-budget", "6"],
                         capture_output=True, text=True, check=True, timeout=600)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0
    # the line states what was measured: ms_per_step is the measured time of one sample evaluation, `value` the workload
    # throughput it implies, `config` names the workload exactly like the B200 arm does
    import bench

    assert line["config"] == bench.workload_config(119, 1, "cube")
    s_ = line["sample"]
    assert abs(line["value"] - s_["fraction_of_workload"] / (line["ms_per_step"] * 1e-3)) <= 1e-9 * line["value"]


def test_quadrature_rules_are_exact():
    """Every rule the device kernels receive integrates all monomials up to its degree (relative to |K|):
    int lam^alpha / |K| = d! alpha! / (|alpha| + d)!  -- in particular the 7- / 14-point degree-5 rules."""
    from math import factorial

    from cashocs_b200._forms import expressions

    for dim in (2, 3):
        for degree in range(0, 7):
            lam, w = expressions.simplex_quadrature(dim, degree)
            assert abs(w.sum() - 1.0) < 1e-15 and np.all(w > 0) and np.all(lam[:, : dim + 1] >= 0)
            assert np.allclose(lam[:, : dim + 1].sum(axis=1), 1.0, atol=1e-15)
            for a in range(degree + 1):
                for b in range(degree + 1 - a):
                    for c in (range(degree + 1 - a - b) if dim == 3 else [0]):
                        exact = factorial(dim) * factorial(a) * factorial(b) * factorial(c) / factorial(a + b + c + dim)
                        got = np.sum(w * lam[:, 1] ** a * lam[:, 2] ** b * (lam[:, 3] ** c if dim == 3 else 1.0))
                        assert abs(got - exact) < 5e-16, (dim, degree, a, b, c)
    assert len(expressions.simplex_quadrature(3, 5)[1]) == 14 and len(expressions.simplex_quadrature(2, 5)[1]) == 7


def test_log_profiling_is_off_by_default():
    """cashocs/log.py:361-389: profiling only at TRACE level; otherwise the decorator is a plain call."""
    from cashocs_b200 import log

    calls = []

    @log.profile_execution_time("unit test action")
    def f(x):
        calls.append(x)
        return x + 1

    log.set_log_level(log.WARNING)
    assert f(1) == 2 and calls == [1] and "unit test action" not in log.timings()
    log.set_log_level(log.TRACE)
    try:
        assert f(2) == 3  # no GPU here: still a plain call
    finally:
        log.set_log_level(log.WARNING)
    assert f.__name__ == "f"


@pytest.mark.skipif(not os.path.isdir("/root/reference/demos"), reason="reference tree not present (GPU box)")
def test_reference_config_files_load_unchanged():
    """Drop-in surface: the reference's own config .ini files (tests/ + the documented shape-optimisation and
    optimal-control demos) load and validate unchanged; the only rejections are explicit ConfigErrors for keys that
    select features outside the hot path (p-Laplacian, distance-based mu, shape regularisation)."""
    import glob

    files = (["/root/reference/tests/config_sop.ini", "/root/reference/tests/config_ocp.ini"]
             + sorted(glob.glob("/root/reference/demos/documented/shape_optimization/*/config.ini"))
             + sorted(glob.glob("/root/reference/demos/documented/optimal_control/*/config.ini")))
    assert len(files) >= 25
    rejected = {}
    for f in files:
        try:
            cfg = cb.load_config(f)
            cfg.validate_config()
        except cb.ConfigError as e:
            rejected[f] = str(e)
    assert len(files) - len(rejected) >= 25
    for f, msg in rejected.items():
        # every message is the device layer's own refusal (or a relative file path of the demo directory) --
        # never one of the reference's validation errors
        lines = [ln for ln in msg.splitlines() if ln.startswith(("Key", "The"))]
        assert lines and all("outside the B200 hot path" in ln or "should point to a file" in ln for ln in lines), (f, msg)
    sop_cfg = cb.load_config("/root/reference/tests/config_sop.ini")
    assert sop_cfg.getfloat("ShapeGradient", "lambda_lame") == 1.428571428571429
    assert sop_cfg.getlist("ShapeGradient", "shape_bdry_def") == [1] and sop_cfg.getfloat("MeshQuality", "tol_upper") == 0.02


def test_gmsh_reader():
    """cashocs_b200.io.mesh.read_gmsh (the source format of cashocs-convert): a hand-written 4.1 ASCII file, and --
    where the reference tree is present -- the reference's own mesh fixtures against the committed .npz goldens."""
    from cashocs_b200.io import mesh as mesh_io

    golden = os.path.join(os.path.dirname(__file__), "golden")
    coords, cells, facets, ftags, ctags = mesh_io.read_gmsh(os.path.join(golden, "tiny_square.msh"))
    assert coords.shape == (5, 2) and cells.shape == (4, 3) and np.all(ctags == 7)
    assert np.array_equal(cells, [[0, 1, 4], [1, 2, 4], [2, 3, 4], [0, 3, 4]])
    assert sorted(zip(map(tuple, facets), ftags)) == [((0, 1), 11), ((0, 3), 14), ((1, 2), 12), ((2, 3), 13)]
    m = cb.import_mesh(os.path.join(golden, "tiny_square.msh"), device="cpu")
    assert m.num_cells == 4 and m.boundary_markers == [11, 12, 13, 14]
    assert m.boundary_vertices([11, 13]).tolist() == [0, 1, 2, 3]
    ref = "/root/reference/tests/mesh"
    if os.path.isdir(ref):
        for name, rel in (("unit_circle", "unit_circle/mesh.msh"), ("mesh2d", "mesh.msh"), ("mesh3d", "mesh3.msh")):
            c, ce, f, ft, _ = mesh_io.read_gmsh(os.path.join(ref, rel))
            z = np.load(os.path.join(golden, name + ".npz"))
            assert np.array_equal(c, z["coords"]) and np.array_equal(ce, np.sort(z["cells"], axis=1))
            assert sorted(map(tuple, np.c_[f, ft])) == sorted(map(tuple, np.c_[np.sort(z["facets"], axis=1), z["facet_tags"]]))
    with pytest.raises(cb.InputError):
        mesh_io.read_gmsh(__file__)  # not a gmsh 4.1 ASCII file


def test_rcb_partition_is_a_balanced_permutation():
    """``distributed.rcb_order``: recursive coordinate bisection into any number of parts -- a permutation of the
    vertices, part sizes within one of each other, parts geometrically compact (each part's bounding box is a
    fraction of the domain's)."""
    from cashocs_b200 import distributed

    rng = np.random.RandomState(3)
    pts = rng.rand(1000, 3)
    for parts in (1, 2, 3, 5, 8):
        order, off = distributed.rcb_order(pts, parts)
        assert np.array_equal(np.sort(order), np.arange(1000)) and off[0] == 0 and off[-1] == 1000
        sizes = np.diff(off)
        assert sizes.size == parts and sizes.max() - sizes.min() <= 1
        if parts == 8:
            vols = [np.prod(np.ptp(pts[order[off[r]:off[r + 1]]], axis=0)) for r in range(parts)]
            assert max(vols) < 0.3
    o1, _ = distributed.rcb_order(pts, 8)
    o2, _ = distributed.rcb_order(pts, 8)
    assert np.array_equal(o1, o2)


def test_dense_polynomial_evaluator(tmp_path):
    """The compile-time Horner tables of the quadrature kernels (csrc/common.cuh: dense_from_poly / dense_eval, every
    (degree, z-dependence) instantiation the launchers select) agree with the generic monomial evaluator to round-off;
    host build of the same header, no GPU involved."""
    import shutil
    import subprocess

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "dense_poly")
    subprocess.run([nvcc, "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", os.path.join(root, "include"), "-I",
                    os.path.join(root, "cashocs_b200", "csrc"), "-o", exe,
                    os.path.join(root, "tests", "native", "test_dense_poly.cu")], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "degree-5 rejected 1" in out


def test_cost_functional_list_algebra():
    """``CostFunctionalList`` (lists of cost functionals, ``ScalarTrackingFunctional``, ``desired_weights``;
    cashocs/_optimization/cost_functional.py:209-316, optimization_problem.py:791-824) without a device: the value is
    the sum of the terms, the effective coefficients carry weight * (integrand - goal) of every tracking term, scaling
    multiplies integral functionals and ASSIGNS the weight of tracking functionals."""
    import types

    from cashocs_b200._forms import forms

    ud = types.SimpleNamespace(vector=None)
    J = [cb.IntegralFunctional(c_state=2.0, c_volume=0.5), cb.IntegralFunctional(c_track=2.0, desired_state=ud),
         cb.ScalarTrackingFunctional(cb.IntegralFunctional(c_volume=1.0, c_surface=3.0), 4.0, weight=0.5)]
    c = forms.as_cost_functional(J)
    ints = (1.5, 0.25, 2.0, 1.0)  # int u, 1/2 int (u - u_d)^2, volume, boundary measure
    c.bind(lambda: ints)
    track = 1.0 * 2.0 + 3.0 * 1.0
    assert c.tracking_values() == [track] and c.has_volume_terms and c.has_surface_terms and c.desired_state is ud
    assert abs(c.value() - (2.0 * 1.5 + 0.5 * 2.0 + 2.0 * 0.25 + 0.5 * 0.5 * (track - 4.0) ** 2)) < 1e-15
    c.update()
    m = 0.5 * (track - 4.0)
    assert (c.c_state, c.c_track, c.c_volume, c.c_surface) == (2.0, 2.0, 0.5 + m * 1.0, m * 3.0)
    vals = c.term_values(ints)
    assert vals == [2.0 * 1.5 + 0.5 * 2.0, 2.0 * 0.25, 0.5 * 0.5 * (track - 4.0) ** 2]
    c.scale([2.0, 3.0, 7.0])
    assert c.base == {"c_state": 4.0, "c_track": 6.0, "c_volume": 1.0, "c_surface": 0.0} and c.tracking[0].weight == 7.0
    assert J[0].c_state == 2.0 and J[2].weight == 0.5  # the caller's objects are untouched
    assert forms.as_cost_functional(J[0]) is J[0] and isinstance(forms.as_cost_functional(J[0], force_list=True), forms.CostFunctionalList)
    with pytest.raises(cb.InputError):
        forms.as_cost_functional([J[1], cb.IntegralFunctional(c_track=1.0, desired_state=types.SimpleNamespace(vector=None))])
    with pytest.raises(cb.InputError):
        cb.ScalarTrackingFunctional(cb.IntegralFunctional(c_control=1.0), 1.0)
