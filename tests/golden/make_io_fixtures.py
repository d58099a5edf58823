"""Golden vectors for the output layer, made by executing the reference's own (pure-Python) functions.

Run in the build container only (``/root/reference`` does not exist on the GPU box):
    python tests/golden/make_io_fixtures.py

``cashocs.io`` cannot be imported here (its package ``__init__`` pulls in FEniCS), but the functions that
define the file formats have no FEniCS dependency: their definitions are cut out of the reference's source
files with ``ast`` and executed unchanged, nothing is copied into this repository.  Produces

* ``output_strings.json``  -- ``generate_output_str`` / ``generate_summary_str`` (``cashocs/io/managers.py:50-142``)
  for a list of optimisation states;
* ``tiny_square_moved.msh`` -- ``parse_file`` (``cashocs/io/mesh.py:434-486``) applied to ``tiny_square.msh`` with
  seeded random vertex positions, and ``mesh3_moved.msh.sha256`` -- the digest of the same for the reference's
  3-D fixture ``tests/mesh/mesh3.msh`` (kept as a digest: the file has 1.5 MB).
"""
import ast
import hashlib
import json
import os
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def cut(path, names):
    """Execute the top-level definitions ``names`` of a reference source file in a fresh namespace."""
    src = open(path, encoding="utf-8").read()
    tree = ast.parse(src)
    keep = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            keep.append(node)
        elif isinstance(node, ast.Assign) and any(getattr(t, "id", None) in names for t in node.targets):
            keep.append(node)
    ns = {"np": np, "__name__": "reference_cut", "annotations": None}
    code = compile(ast.Module(body=[ast.parse("from __future__ import annotations").body[0]] + keep, type_ignores=[]),
                   path, "exec")
    exec(code, ns)
    return ns


def states():
    rng = np.random.RandomState(42)
    out = []
    for it in (0, 1, 9, 10, 11, 123):
        s = {"stepsize": float(10.0 ** rng.uniform(-6, 1)), "no_state_solves": 3 * it + 1, "no_adjoint_solves": it + 1,
             "mesh_quality": float(rng.uniform(0.1, 1.0)), "iteration": it,
             "objective_value": float(rng.normal() * 10.0 ** rng.randint(-8, 5)),
             "gradient_norm": float(10.0 ** rng.uniform(-9, 2)), "gradient_norm_initial": 3.25,
             "relative_norm": float(10.0 ** rng.uniform(-9, 0))}
        out.append(s)
    control = dict(out[2])
    del control["mesh_quality"]
    out.append(control)
    nan_state = dict(out[3])
    nan_state["gradient_norm"] = float("nan")
    nan_state["relative_norm"] = float("nan")
    out.append(nan_state)
    return out


def main():
    mgr = cut(f"{REF}/cashocs/io/managers.py", {"output_mapping", "generate_output_str", "generate_summary_str"})
    records = []
    for s in states():
        db = types.SimpleNamespace(parameter_db=types.SimpleNamespace(optimization_state=s))
        for precision in (3, 6):
            records.append({"state": s, "precision": precision, "output": mgr["generate_output_str"](db, precision),
                            "summary": mgr["generate_summary_str"](db, precision)})
    with open(os.path.join(HERE, "output_strings.json"), "w", encoding="utf-8") as fh:
        json.dump(records, fh, indent=1)

    io = cut(f"{REF}/cashocs/io/mesh.py", {"create_point_representation", "parse_file"})
    rng = np.random.RandomState(7)
    pts = rng.uniform(-2.0, 2.0, size=(5, 2)) * 10.0 ** rng.randint(-3, 3, size=(5, 1))
    np.save(os.path.join(HERE, "tiny_square_moved_points.npy"), pts)
    io["parse_file"](os.path.join(HERE, "tiny_square.msh"), os.path.join(HERE, "tiny_square_moved.msh"), pts, 2)

    # 3-D: the reference's own fixture, all nodes used, tags 1..N
    src = f"{REF}/tests/mesh/mesh3.msh"
    n = None
    with open(src) as fh:
        for line in fh:
            if line == "$Nodes\n":
                n = int(fh.readline().split()[1])
                break
    pts3 = np.random.RandomState(11).normal(size=(n, 3))
    tmp = "/tmp/mesh3_moved.msh"
    io["parse_file"](src, tmp, pts3, 3)
    digest = hashlib.sha256(open(tmp, "rb").read()).hexdigest()
    with open(os.path.join(HERE, "mesh3_moved.msh.sha256"), "w") as fh:
        fh.write(f"{digest}  seed=11 normal(size=({n}, 3))\n")
    print("wrote", len(records), "records;", "mesh3 nodes", n, digest)


if __name__ == "__main__":
    main()
