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
    for sec, key, val in (("ShapeGradient", "use_p_laplacian", "True"), ("Regularization", "factor_curvature", "1.0"),
                          ("StateSystem", "backend", "petsc"), ("AlgoLBFGS", "damped", "True"),
                          ("ShapeGradient", "global_deformation", "True")):
        cfg = cb.Config()
        cfg.set(sec, key, val)
        with pytest.raises(cb.ConfigError, match="outside the B200 hot path"):
            cfg.validate_config()
