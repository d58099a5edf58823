"""Config ``.ini`` handling compatible with ``cashocs.load_config`` (``cashocs/io/config.py``).

Reads the same files, provides the same defaults for every key the shape-gradient hot path
reads (``cashocs/io/config.py:593-741``), ``getlist`` via JSON (``:764-784``) and a validation
pass that collects errors into :class:`ConfigError` (``:786-1065``).  Sections that only
steer subsystems outside SURVEY.md section 8 (remeshing, topology optimisation, output files)
are parsed but not interpreted.
"""

from __future__ import annotations

import configparser
import json
import pathlib

from cashocs_b200 import _exceptions

_DEFAULTS = {
    "Mesh": {"remesh": "False", "show_gmsh_output": "False"},
    "StateSystem": {
        "is_linear": "False", "newton_rtol": "1e-11", "newton_atol": "1e-13", "newton_iter": "50",
        "picard_iteration": "False", "picard_rtol": "1e-10", "picard_atol": "1e-12", "picard_iter": "50",
        "picard_verbose": "False", "backend": "cashocs",
    },
    "OptimizationRoutine": {
        "algorithm": "none", "rtol": "1e-3", "atol": "0.0", "max_iter": "100", "soft_exit": "False",
        "gradient_tol": "1e-9", "gradient_method": "direct",
    },
    "LineSearch": {
        "method": "armijo", "epsilon_armijo": "1e-4", "beta_armijo": "2.0", "initial_stepsize": "1.0",
        "safeguard_stepsize": "True", "fail_if_not_converged": "False",
    },
    "ShapeGradient": {
        "lambda_lame": "0.0", "damping_factor": "0.0", "mu_def": "1.0", "mu_fix": "1.0",
        "use_sqrt_mu": "False", "use_p_laplacian": "False", "use_pull_back": "True",
        "use_distance_mu": "False", "inhomogeneous": "False", "update_inhomogeneous": "False",
        "inhomogeneous_exponent": "1.0", "fixed_dimensions": "[]", "shape_bdry_def": "[]",
        "shape_bdry_fix": "[]", "shape_bdry_fix_x": "[]", "shape_bdry_fix_y": "[]", "shape_bdry_fix_z": "[]",
        "degree_estimation": "True", "test_for_intersections": "True", "reextend_from_boundary": "False",
    },
    "Regularization": {
        "factor_volume": "0.0", "target_volume": "0.0", "use_initial_volume": "False",
        "factor_surface": "0.0", "target_surface": "0.0", "use_initial_surface": "False",
        "factor_curvature": "0.0", "factor_barycenter": "0.0", "target_barycenter": "[0.0, 0.0, 0.0]",
        "use_initial_barycenter": "False", "use_relative_scaling": "False",
    },
    "AlgoLBFGS": {"bfgs_memory_size": "5", "use_bfgs_scaling": "True", "bfgs_periodic_restart": "0", "damped": "False"},
    "AlgoCG": {"cg_method": "DY", "cg_periodic_restart": "False", "cg_periodic_its": "10",
               "cg_relative_restart": "False", "cg_restart_tol": "0.25"},
    "MeshQuality": {
        "tol_lower": "0.0", "tol_upper": "1e-15", "measure": "skewness", "type": "min", "quantile": "0.0",
        "volume_change": "inf", "angle_change": "inf", "remesh_iter": "0",
    },
    "Output": {"save_results": "True", "verbose": "False", "save_txt": "False", "save_state": "False",
               "save_adjoint": "False", "save_gradient": "False", "save_mesh": "False",
               "result_dir": "./results", "precision": "3", "time_suffix": "False"},
}

# (section, key) -> (type, constraints)
_SCHEMA = {
    ("StateSystem", "is_linear"): ("bool", {}),
    ("StateSystem", "picard_iteration"): ("bool", {}),
    ("OptimizationRoutine", "rtol"): ("float", {"less_than_one": True, "non_negative": True}),
    ("OptimizationRoutine", "atol"): ("float", {"non_negative": True}),
    ("OptimizationRoutine", "max_iter"): ("int", {"non_negative": True}),
    ("OptimizationRoutine", "gradient_tol"): ("float", {"less_than_one": True, "positive": True}),
    ("OptimizationRoutine", "gradient_method"): ("str", {"options": ["direct", "iterative"]}),
    ("OptimizationRoutine", "algorithm"): ("str", {"options": [
        "gd", "gradient_descent", "bfgs", "lbfgs", "cg", "ncg", "nonlinear_cg", "conjugate_gradient",
        "newton", "none"]}),
    ("LineSearch", "method"): ("str", {"options": ["armijo", "polynomial"]}),
    ("LineSearch", "epsilon_armijo"): ("float", {"positive": True, "less_than_one": True}),
    ("LineSearch", "beta_armijo"): ("float", {"larger_than_one": True}),
    ("LineSearch", "initial_stepsize"): ("float", {"positive": True}),
    ("LineSearch", "safeguard_stepsize"): ("bool", {}),
    ("ShapeGradient", "lambda_lame"): ("float", {}),
    ("ShapeGradient", "damping_factor"): ("float", {"non_negative": True}),
    ("ShapeGradient", "mu_def"): ("float", {"positive": True}),
    ("ShapeGradient", "mu_fix"): ("float", {"positive": True}),
    ("ShapeGradient", "use_sqrt_mu"): ("bool", {}),
    ("ShapeGradient", "inhomogeneous"): ("bool", {}),
    ("ShapeGradient", "update_inhomogeneous"): ("bool", {}),
    ("ShapeGradient", "inhomogeneous_exponent"): ("float", {}),
    ("ShapeGradient", "test_for_intersections"): ("bool", {}),
    ("ShapeGradient", "shape_bdry_def"): ("list", {}),
    ("ShapeGradient", "shape_bdry_fix"): ("list", {}),
    ("ShapeGradient", "shape_bdry_fix_x"): ("list", {}),
    ("ShapeGradient", "shape_bdry_fix_y"): ("list", {}),
    ("ShapeGradient", "shape_bdry_fix_z"): ("list", {}),
    ("ShapeGradient", "fixed_dimensions"): ("list", {}),
    ("AlgoLBFGS", "bfgs_memory_size"): ("int", {"non_negative": True}),
    ("AlgoLBFGS", "use_bfgs_scaling"): ("bool", {}),
    ("MeshQuality", "tol_lower"): ("float", {"non_negative": True, "less_than_one": True}),
    ("MeshQuality", "tol_upper"): ("float", {"non_negative": True, "less_than_one": True}),
    ("MeshQuality", "measure"): ("str", {"options": ["skewness", "maximum_angle", "radius_ratios", "condition_number"]}),
    ("MeshQuality", "type"): ("str", {"options": ["min", "minimum", "avg", "average", "quantile"]}),
    ("MeshQuality", "quantile"): ("float", {"non_negative": True, "less_than_one_incl": True}),
    ("MeshQuality", "volume_change"): ("float", {"larger_than_one": True}),
    ("MeshQuality", "angle_change"): ("float", {"positive": True}),
}


class Config(configparser.ConfigParser):
    """Class for handling the config in cashocs (``cashocs/io/config.py:81-1065``)."""

    def __init__(self, config_file: str | None = None) -> None:
        super().__init__()
        self.config_errors: list[str] = []
        self.read_dict(_DEFAULTS)
        if config_file is not None:
            path = pathlib.Path(config_file)
            if not path.is_file():
                raise _exceptions.InputError("cashocs_b200.io.config.Config", "config_file",
                                             "The file you specified does not exist.")
            self.read(config_file)

    def getlist(self, section: str, option: str, **kwargs) -> list:
        """Extract a list via JSON (``cashocs/io/config.py:764-784``)."""
        try:
            out = json.loads(self.get(section, option, **kwargs))
        except json.JSONDecodeError as exc:
            raise _exceptions.ConfigError([f"Key {option} in section {section} is not a valid list.\n"]) from exc
        if not isinstance(out, list):
            out = [out]
        return out

    def validate_config(self) -> None:
        """Collect all problems, then raise one :class:`ConfigError` (``:786-792``)."""
        self.config_errors = []
        for section in self.sections():
            if section not in _DEFAULTS and section not in ("AlgoTNM", "TopologyOptimization", "Debug"):
                self.config_errors.append(f"The following section is not valid: {section}\n")
                continue
            for key in self[section]:
                sch = _SCHEMA.get((section, key))
                if section in _DEFAULTS and key not in _DEFAULTS[section] and sch is None and section in (
                    "LineSearch", "MeshQuality", "OptimizationRoutine", "AlgoLBFGS",
                ):
                    self.config_errors.append(f"Key {key} is not valid for section {section}.\n")
                    continue
                if sch is None:
                    continue
                self._check_key(section, key, *sch)
        lo = self._try_float("MeshQuality", "tol_lower")
        up = self._try_float("MeshQuality", "tol_upper")
        if lo is not None and up is not None and lo > up:
            self.config_errors.append(
                "The value of key tol_upper in section MeshQuality is smaller than the value of key tol_lower"
                " in section MeshQuality, but it should be larger.\n"
            )
        for key, off in (("use_p_laplacian", False), ("use_distance_mu", False), ("reextend_from_boundary", False)):
            try:
                if self.getboolean("ShapeGradient", key) != off:
                    self.config_errors.append(
                        f"Key {key} in section ShapeGradient is outside the B200 hot path (SURVEY.md 8a) and must be False.\n"
                    )
            except ValueError:
                self.config_errors.append(f"Key {key} in section ShapeGradient is the wrong type. Required type is bool.\n")
        for key in ("factor_volume", "factor_surface", "factor_curvature", "factor_barycenter"):
            v = self._try_float("Regularization", key)
            if v is not None and v != 0.0:
                self.config_errors.append(
                    f"Key {key} in section Regularization: shape regularisation is outside the B200 hot path and must be 0.\n"
                )
        if self.config_errors:
            raise _exceptions.ConfigError(self.config_errors)

    def _try_float(self, section, key):
        try:
            return self.getfloat(section, key)
        except ValueError:
            return None

    def _check_key(self, section, key, typ, cons) -> None:
        try:
            if typ == "bool":
                self.getboolean(section, key)
                return
            if typ == "int":
                v = self.getint(section, key)
            elif typ == "float":
                v = self.getfloat(section, key)
            elif typ == "list":
                self.getlist(section, key)
                return
            else:
                v = self.get(section, key)
        except (ValueError, _exceptions.ConfigError):
            self.config_errors.append(f"Key {key} in section {section} is the wrong type. Required type is {typ}.\n")
            return
        if "options" in cons and str(v).casefold() not in cons["options"]:
            self.config_errors.append(
                f"Key {key} in section {section} has a wrong value. Possible options are {cons['options']}.\n"
            )
        if cons.get("non_negative") and v < 0:
            self.config_errors.append(f"Key {key} in section {section} is negative, but it must not be.\n")
        if cons.get("positive") and v <= 0:
            self.config_errors.append(f"Key {key} in section {section} is non-positive, but it most be positive.\n")
        if cons.get("less_than_one") and v >= 1:
            self.config_errors.append(f"Key {key} in section {section} is larger than one, but it must be smaller.\n")
        if cons.get("less_than_one_incl") and v > 1:
            self.config_errors.append(f"Key {key} in section {section} is larger than one, but it must be smaller.\n")
        if cons.get("larger_than_one") and v <= 1:
            self.config_errors.append(f"Key {key} in section {section} is smaller than one, but it must be larger.\n")


def load_config(path: str) -> Config:
    """Loads a config object from a config file (``cashocs/io/config.py:37-49``)."""
    return Config(path)
