"""Golden vectors for the config boundary, made by running the reference's own ``Config`` class.

Run in the build container only (``/root/reference`` does not exist on the GPU box):
    python tests/golden/make_config_fixture.py

``cashocs/io/config.py`` depends on nothing but ``cashocs/_exceptions.py`` (both pure Python), so the two modules
are loaded from where they lie under a stub ``cashocs`` package -- no FEniCS needed, nothing copied.  Every case
is a list of ``[section, key, value]`` overrides on top of the defaults; recorded are the outcome of
``validate_config()`` (``ok`` or the exception type), the collected ``config_errors`` and, for list-typed keys,
the outcome of ``getlist``.  Written to ``config_cases.json``; ``defaults`` holds the reference's default of every key.
"""
import importlib.util
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/cashocs"


def load_reference_config():
    pkg = types.ModuleType("cashocs")
    pkg.__path__ = []
    sys.modules["cashocs"] = pkg

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    pkg._exceptions = load("cashocs._exceptions", f"{REF}/_exceptions.py")
    io = types.ModuleType("cashocs.io")
    io.__path__ = []
    sys.modules["cashocs.io"] = io
    return load("cashocs.io.config", f"{REF}/io/config.py")


PROBES = {
    "bool": ["maybe", "True", "false", "1", "yes"],
    "int": ["-1", "0", "1", "2", "1.5", "abc"],
    "float": ["-1.0", "0.0", "0.5", "1.0", "2.0", "abc", "inf", "1e-3"],
    "list": ["[1, 2]", "1,2", "[a;b]", "[1.5, 2]", "[]", '["x", "y"]', "['x']", "[1, [2, 3]]"],
}


def cases(ref):
    scheme = ref.Config().config_scheme
    out = [[]]
    for sec, keys in scheme.items():
        for key, entry in keys.items():
            typ = entry["type"]
            if "possible_options" in entry:
                vals = ["nonsense"] + [o.upper() for o in entry["possible_options"]]
            elif "file" in entry.get("attributes", []):
                vals = ["/nonexistent/a." + entry["file_extension"], "/nonexistent/a.txt", os.path.join(HERE, "tiny_square.msh")]
            elif typ == "str":
                vals = ["./some/where/", "x"]
            else:
                vals = PROBES[typ]
            out += [[[sec, key, v]] for v in vals]
    out += [
        [["Nonsense", "key", "1"]],
        [["Mesh", "nonsense", "1"]],
        [["ShapeGradient", "nonsense", "1"], ["Output", "nonsense", "x"]],
        [["MeshQuality", "tol_lower", "0.5"], ["MeshQuality", "tol_upper", "0.1"]],
        [["MeshQuality", "tol_lower", "0.1"], ["MeshQuality", "tol_upper", "0.1"]],
        [["MeshQuality", "tol_lower", "0.1"], ["MeshQuality", "tol_upper", "0.5"]],
        [["LineSearch", "factor_low", "0.6"]],
        [["LineSearch", "factor_low", "0.5"]],
        [["ShapeGradient", "dist_min", "2.0"]],
        [["ShapeGradient", "dist_min", "0.5"], ["ShapeGradient", "dist_max", "0.5"]],
        [["Regularization", "x_start", "1.0"]],
        [["Regularization", "y_start", "3.0"], ["Regularization", "y_end", "2.0"]],
        [["Mesh", "remesh", "True"]],
        [["Mesh", "remesh", "True"], ["Mesh", "gmsh_file", os.path.join(HERE, "tiny_square.msh")]],
        [["Output", "save_mesh", "True"]],
        [["Output", "save_mesh", "TRUE"], ["Mesh", "gmsh_file", "/nonexistent/b.msh"]],
        [["OptimizationRoutine", "rtol", "2.0"], ["OptimizationRoutine", "atol", "-1"], ["OptimizationRoutine", "max_iter", "-3"]],
        [["AlgoCG", "cg_method", "XY"], ["AlgoLBFGS", "bfgs_memory_size", "-2"], ["Output", "precision", "0"]],
    ]
    return out


def main():
    ref = load_reference_config()
    exc = sys.modules["cashocs._exceptions"]
    records = []
    for overrides in cases(ref):
        cfg = ref.Config()
        for sec, key, val in overrides:
            if not cfg.has_section(sec):
                cfg.add_section(sec)
            cfg.set(sec, key, val)
        try:
            cfg.validate_config()
            outcome = "ok"
        except exc.ConfigError:
            outcome = "ConfigError"
        except Exception as e:  # the reference lets ValueError escape for unparsable numbers
            outcome = type(e).__name__
        rec = {"overrides": overrides, "outcome": outcome, "errors": list(cfg.config_errors)}
        for sec, key, val in overrides:
            if sec in cfg.config_scheme and cfg.config_scheme[sec].get(key, {}).get("type") == "list":
                try:
                    rec["getlist"] = cfg.getlist(sec, key)
                except Exception as e:
                    rec["getlist"] = "raises " + type(e).__name__
        records.append(rec)
    base = ref.Config()
    defaults = {s: dict(base[s]) for s in base.sections()}
    with open(os.path.join(HERE, "config_cases.json"), "w", encoding="utf-8") as fh:
        # the fixtures directory is written as a placeholder so the file does not depend on where the repo lies
        fh.write(json.dumps({"defaults": defaults, "cases": records}, indent=0).replace(HERE, "{GOLDEN}"))
    n = {k: sum(r["outcome"] == k for r in records) for k in sorted({r["outcome"] for r in records})}
    print(len(records), "cases", n)


if __name__ == "__main__":
    main()
