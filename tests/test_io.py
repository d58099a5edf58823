"""Output layer (``cashocs/io``): gmsh export of moved meshes, console / history strings, result files.
Host-only logic -- runs without a GPU; the golden vectors were produced by executing the reference's own
pure-Python functions (``tests/golden/make_io_fixtures.py``)."""

import hashlib
import json
import os
import types

import numpy as np
import pytest
import torch

import cashocs_b200 as cb
from cashocs_b200.io import managers
from cashocs_b200.io import mesh as iomesh

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_output_strings_match_reference_goldens():
    records = json.load(open(os.path.join(GOLDEN, "output_strings.json")))
    assert len(records) == 16
    for r in records:
        assert managers.generate_output_str(r["state"], r["precision"]) == r["output"]
        assert managers.generate_summary_str(r["state"], r["precision"]) == r["summary"]


def test_write_out_mesh_matches_reference_golden(tmp_path):
    mesh = cb.import_mesh(os.path.join(GOLDEN, "tiny_square.msh"), device="cpu")
    pts = np.load(os.path.join(GOLDEN, "tiny_square_moved_points.npy"))
    mesh.coords.copy_(torch.from_numpy(pts))
    out = tmp_path / "sub" / "moved.msh"  # parent directories are created (io/mesh.py:548-549)
    cb.io.write_out_mesh(mesh, os.path.join(GOLDEN, "tiny_square.msh"), str(out))
    assert out.read_text() == open(os.path.join(GOLDEN, "tiny_square_moved.msh")).read()
    # and the file round-trips through the reader bit for bit
    coords, cells, facets, tags, _ = iomesh.read_gmsh(str(out))
    assert np.array_equal(coords, pts)
    assert np.array_equal(cells, mesh.cells.numpy())


@pytest.mark.skipif(not os.path.isfile("/root/reference/tests/mesh/mesh3.msh"), reason="reference tree not present")
def test_write_out_mesh_3d_digest_matches_reference(tmp_path):
    src = "/root/reference/tests/mesh/mesh3.msh"
    mesh = cb.import_mesh(src, device="cpu")
    pts = np.random.RandomState(11).normal(size=(mesh.num_vertices, 3))
    mesh.coords.copy_(torch.from_numpy(pts))
    out = tmp_path / "m3.msh"
    iomesh.write_out_mesh(mesh, src, str(out))
    want = open(os.path.join(GOLDEN, "mesh3_moved.msh.sha256")).read().split()[0]
    assert hashlib.sha256(out.read_bytes()).hexdigest() == want
    # meshes built from converted arrays (no node tags) use the reference's "vertex v <-> node v + 1" convention
    del mesh.gmsh_node_tags
    iomesh.write_out_mesh(mesh, src, str(out))
    assert hashlib.sha256(out.read_bytes()).hexdigest() == want


def test_write_out_mesh_rejects_an_incompatible_file(tmp_path):
    mesh = cb.import_mesh(os.path.join(GOLDEN, "unit_circle.npz"), device="cpu")
    with pytest.raises(cb.CashocsException, match="not compatible"):
        iomesh.write_out_mesh(mesh, os.path.join(GOLDEN, "tiny_square.msh"), str(tmp_path / "x.msh"))


class _FakeSolver:
    """The attributes ``OptimizationAlgorithm.optimization_state`` reads, without a device."""


def _fake_problem(tmp_path, **output):
    cfg = cb.Config()
    cfg.set("Output", "result_dir", str(tmp_path / "out") + "/")
    for k, v in output.items():
        cfg.set("Output", k, str(v))
    mesh = cb.import_mesh(os.path.join(GOLDEN, "tiny_square.msh"), device="cpu")
    f = types.SimpleNamespace(vector=torch.arange(5, dtype=torch.float64))
    return types.SimpleNamespace(config=cfg, mesh=mesh, states=[f], adjoints=[f], gradient=f, problem_type="shape")


def test_output_manager_writes_the_reference_files(tmp_path, capsys):
    prob = _fake_problem(tmp_path, save_txt=True, save_mesh=True, save_state=True, verbose=True)
    om = managers.OutputManager(prob)
    assert om.result_dir == str(tmp_path / "out")
    states = [r["state"] for r in json.load(open(os.path.join(GOLDEN, "output_strings.json"))) if r["precision"] == 3]
    s0, s1 = dict(states[0]), dict(states[1])
    om.output(s0)
    prob.mesh.coords.add_(0.25)
    om.output(s1)
    om.post_process(s1)
    om.output_summary(s1)
    out = tmp_path / "out"
    hist = json.load(open(out / "history.json"))
    assert hist["iteration"] == [0, 1] and hist["objective_value"] == [s0["objective_value"], s1["objective_value"]]
    assert set(hist) == set(s0)
    txt = (out / "history.txt").read_text()
    assert txt == (managers.generate_output_str(s0, 3) + "\n" + managers.generate_output_str(s1, 3) + "\n"
                   + managers.generate_summary_str(s1, 3))
    assert "Optimization was successful." in capsys.readouterr().out
    for name in ("checkpoints/mesh/mesh_0.msh", "checkpoints/mesh/mesh_1.msh", "optimized_mesh.msh",
                 "checkpoints/state_0_0.npy", "checkpoints/state_0_1.npy"):
        assert (out / name).is_file(), name
    c0 = iomesh.read_gmsh(str(out / "checkpoints/mesh/mesh_0.msh"))[0]
    c1 = iomesh.read_gmsh(str(out / "optimized_mesh.msh"))[0]
    assert np.array_equal(c1, c0 + 0.25)


def test_output_manager_is_silent_on_disk_when_nothing_is_saved(tmp_path):
    prob = _fake_problem(tmp_path, save_results=False)
    om = managers.OutputManager(prob)
    om.output({"iteration": 0, "objective_value": 1.0, "stepsize": 1.0})
    om.post_process({"iteration": 0})
    assert not (tmp_path / "out").exists()


def test_time_suffix(tmp_path):
    prob = _fake_problem(tmp_path, time_suffix=True)
    om = managers.OutputManager(prob)
    assert om.result_dir.startswith(str(tmp_path / "out") + "_") and om.result_dir.endswith(om.suffix)
    assert os.path.isdir(om.result_dir)


# ------------------------------------------------------------------ config boundary vs the reference's Config class
def _config_cases():
    text = open(os.path.join(GOLDEN, "config_cases.json")).read().replace("{GOLDEN}", GOLDEN)
    return json.loads(text)


def test_config_defaults_equal_the_reference():
    ref = _config_cases()["defaults"]
    ours = cb.Config()
    assert ours.sections() == list(ref)
    for section, keys in ref.items():
        assert dict(ours[section]) == keys, section  # same keys, same order, same default strings


def test_config_validation_matches_the_reference_case_by_case():
    """798 configs through the reference's own ``validate_config``: same outcome (ok / ConfigError / escaping
    ValueError), same messages in the same order, same ``getlist`` results.  The device layer may add its own
    "outside the B200 hot path" errors after the reference's."""
    cases = _config_cases()["cases"]
    assert len(cases) == 798
    for rec in cases:
        cfg = cb.Config()
        for sec, key, val in rec["overrides"]:
            if not cfg.has_section(sec):
                cfg.add_section(sec)
            cfg.set(sec, key, val)
        try:
            cfg.validate_config()
            outcome = "ok"
        except cb.ConfigError:
            outcome = "ConfigError"
        except Exception as e:
            outcome = type(e).__name__
        extra = [e for e in cfg.config_errors if "outside the B200 hot path" in e]
        mirrored = [e for e in cfg.config_errors if "outside the B200 hot path" not in e]
        assert mirrored == rec["errors"], rec["overrides"]
        if extra and rec["outcome"] == "ok":
            assert outcome == "ConfigError", rec["overrides"]
        else:
            assert outcome == rec["outcome"], rec["overrides"]
        if "getlist" in rec:
            sec, key, _ = next(o for o in rec["overrides"] if cfg.config_scheme.get(o[0], {}).get(o[1], {}).get("type") == "list")
            try:
                got = cfg.getlist(sec, key)
            except Exception as e:
                got = "raises " + type(e).__name__
            assert got == rec["getlist"], rec["overrides"]


def test_unsupported_subsystems_are_refused_loudly():
    for sec, key, val in (("StateSystem", "backend", "petsc"), ("Mesh", "remesh", "True"),
                          ("MeshQuality", "remesh_iter", "3")):
        cfg = cb.Config()
        cfg.set(sec, key, val)
        with pytest.raises(cb.ConfigError, match="outside the B200 hot path"):
            cfg.validate_config()


# ------------------------------------------------------------------ the reference's own config tests (tests/test_config.py)
def _sop_like_config():
    """The keys of the reference's tests/config_sop.ini that matter for validation."""
    cfg = cb.Config()
    cfg.set("ShapeGradient", "shape_bdry_def", "[1]")
    cfg.set("MeshQuality", "tol_lower", "0.01")
    cfg.set("MeshQuality", "tol_upper", "0.02")
    return cfg


REFERENCE_CONFIG_ERRORS = [
    # (tests/test_config.py line, overrides, expected message)
    (109, [("Mesh", "remesh", "1.0")], "Key remesh in section Mesh has the wrong type. Required type is bool."),
    (131, [("MeshQuality", "tol_lower", "0.5"), ("MeshQuality", "tol_upper", "0.1")],
     "The value of key tol_upper in section MeshQuality is smaller than the value of key tol_lower in section "
     "MeshQuality, but it should be larger."),
    (144, [("ShapeGradient", "dist_max", "0.5")],
     "The value of key dist_max in section ShapeGradient is smaller than the value of key dist_min in section "
     "ShapeGradient, but it should be larger."),
    (156, [("Regularization", "x_end", "-1.0")],
     "The value of key x_end in section Regularization is smaller than the value of key x_start in section "
     "Regularization, but it should be larger."),
    (168, [("Mesh", "geo_file", "test.msh")],
     "Key geo_file in section Mesh has the wrong file extension, it should end in .geo."),
    (180, [("MeshQuality", "tol_lower", "-1e-1")], "Key tol_lower in section MeshQuality is negative, but it must not be."),
    (192, [("MeshQuality", "tol_upper", "0.0")],
     "Key tol_upper in section MeshQuality is non-positive, but it most be positive."),
    (204, [("MeshQuality", "tol_upper", "2.0")],
     "Key tol_upper in section MeshQuality is larger than one, but it must be smaller."),
    (216, [("MeshQuality", "volume_change", "0.5")],
     "Key volume_change in section MeshQuality is smaller than one, but it must be larger."),
    (228, [("OptimizationRoutine", "gradient_method", "mymethod")],
     "Key gradient_method in section OptimizationRoutine has a wrong value. Possible options are ['direct', 'iterative']."),
    (240, [("Output", "save_mesh", "True")],
     "Key save_mesh in section Output requires key gmsh_file in section Mesh to be present."),
]


@pytest.mark.parametrize("line,overrides,message", REFERENCE_CONFIG_ERRORS, ids=[str(c[0]) for c in REFERENCE_CONFIG_ERRORS])
def test_reference_config_tests(line, overrides, message):
    """tests/test_config.py of the reference, test by test: the same ConfigError with the same message."""
    cfg = _sop_like_config()
    cfg.validate_config()
    for sec, key, val in overrides:
        cfg.set(sec, key, val)
    with pytest.raises(cb.ConfigError) as e:
        cfg.validate_config()
    assert "You have some error(s) in your config file" in str(e.value)
    assert message in str(e.value)


def test_reference_incorrect_config_file(tmp_path):
    """tests/test_config.py:117-126 with a file shaped like tests/test_config.ini: unknown sections, misplaced key."""
    path = tmp_path / "bad.ini"
    path.write_text("[A]\na = 1\n\n[B]\nb = 2\n\n[StateSystem]\nalgorithm = False\n")
    cfg = cb.load_config(str(path))
    with pytest.raises(cb.ConfigError) as e:
        cfg.validate_config()
    assert ("The following section is not valid: <Section: A>\nThe following section is not valid: <Section: B>\n"
            "Key algorithm is not valid for section StateSystem.") in str(e.value)
