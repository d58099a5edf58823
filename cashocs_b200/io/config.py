"""Config ``.ini`` handling compatible with ``cashocs.load_config`` (``cashocs/io/config.py``).

The config file is part of the drop-in boundary: the same files load, every key of the reference has the same
type, default (``cashocs/io/config.py:593-741``) and constraints (``:94-591``), ``getlist`` goes through the same
character whitelist + JSON (``:52-78,761-784``), and ``validate_config`` performs the reference's checks in the
reference's order, collecting the same messages into :class:`ConfigError` (``:786-1065``) -- including the
``ValueError`` the reference lets escape when a numeric key with range attributes cannot be parsed.  This is
pinned by ``tests/golden/config_cases.json``, produced by running the reference's own ``Config`` class
(``tests/golden/make_config_fixture.py``).

After the reference's checks, :meth:`Config.validate_config` adds the restrictions of the device layer: keys
that select subsystems outside the B200 hot path (p-Laplace, distance-based mu,
regularisation, remeshing, ...) are reported instead of being silently ignored.

The key table below is written as ``key | type | default | constraints`` rows per section; constraints:
``>=0`` non-negative, ``>0`` positive, ``<1`` less than one, ``>1`` larger than one, ``in(a,b)`` possible
options (case-folded), ``file(ext)``, ``needs(Section.key)`` (for a bool that is true), ``gt(Section.key)`` /
``ge(Section.key)`` ordering relative to another key.
"""

from __future__ import annotations

import configparser
import json
import pathlib
import re

from cashocs_b200 import _exceptions

_TABLE = """
[Mesh]
gmsh_file | str | | file(msh)
geo_file | str | | file(geo)
remesh | bool | False | needs(Mesh.gmsh_file)
show_gmsh_output | bool | False |

[StateSystem]
is_linear | bool | False |
newton_rtol | float | 1e-11 | <1 >0
newton_atol | float | 1e-13 | >=0
newton_iter | int | 50 | >=0
newton_damped | bool | False |
newton_inexact | bool | False |
newton_verbose | bool | False |
picard_iteration | bool | False |
picard_rtol | float | 1e-10 | >0 <1
picard_atol | float | 1e-12 | >=0
picard_iter | int | 50 | >=0
picard_verbose | bool | False |
backend | str | cashocs | in(cashocs,petsc)
use_adjoint_linearizations | bool | False |

[OptimizationRoutine]
algorithm | str | none | in(gd,gradient_descent,bfgs,lbfgs,nonlinear_cg,ncg,nonlinear_conjugate_gradient,conjugate_gradient,newton,sphere_combination,convex_combination,none)
rtol | float | 1e-3 | <1 >=0
atol | float | 0.0 | >=0
max_iter | int | 100 | >=0
soft_exit | bool | False |
gradient_tol | float | 1e-9 | <1 >0
gradient_method | str | direct | in(direct,iterative)

[LineSearch]
method | str | armijo | in(armijo,polynomial,basic)
epsilon_armijo | float | 1e-4 | >0 <1
beta_armijo | float | 2.0 | >0 >1
initial_stepsize | float | 1.0 | >0
safeguard_stepsize | bool | True |
polynomial_model | str | cubic | in(cubic,quadratic)
factor_high | float | 0.5 | <1 >0 gt(LineSearch.factor_low)
factor_low | float | 0.1 | <1 >0
fail_if_not_converged | bool | False |

[ShapeGradient]
lambda_lame | float | 0.0 |
damping_factor | float | 0.0 | >=0
mu_def | float | 1.0 | >0
mu_fix | float | 1.0 | >0
use_sqrt_mu | bool | False |
use_p_laplacian | bool | False |
p_laplacian_power | int | 2 | >1
p_laplacian_stabilization | float | 0.0 | >=0 <1
use_pull_back | bool | True |
use_distance_mu | bool | False |
mu_min | float | 1.0 | >0
mu_max | float | 1.0 | >0
dist_min | float | 1.0 | >=0
dist_max | float | 1.0 | >=0 ge(ShapeGradient.dist_min)
boundaries_dist | list | [] |
distance_method | str | eikonal | in(eikonal,poisson)
smooth_mu | bool | False |
inhomogeneous | bool | False |
update_inhomogeneous | bool | False |
inhomogeneous_exponent | float | 1.0 | >=0
fixed_dimensions | list | [] |
shape_bdry_def | list | [] |
shape_bdry_fix | list | [] |
shape_bdry_fix_x | list | [] |
shape_bdry_fix_y | list | [] |
shape_bdry_fix_z | list | [] |
shape_volume_fix | list | [] |
degree_estimation | bool | True |
global_deformation | bool | False |
test_for_intersections | bool | True |
reextend_from_boundary | bool | False |
reextension_mode | str | surface | in(surface,normal)

[Regularization]
factor_volume | float | 0.0 | >=0
target_volume | float | 0.0 | >=0
use_initial_volume | bool | False |
factor_surface | float | 0.0 | >=0
target_surface | float | 0.0 | >=0
use_initial_surface | bool | False |
factor_curvature | float | 0.0 | >=0
factor_barycenter | float | 0.0 | >=0
target_barycenter | list | [0.0, 0.0, 0.0] |
use_initial_barycenter | bool | False |
use_relative_scaling | bool | False |
x_start | float | 0.0 |
x_end | float | 1.0 | gt(Regularization.x_start)
y_start | float | 0.0 |
y_end | float | 1.0 | gt(Regularization.y_start)
z_start | float | 0.0 |
z_end | float | 1.0 | gt(Regularization.z_start)

[AlgoTNM]
inner_newton | str | cr | in(cg,cr)
max_it_inner_newton | int | 50 | >=0
inner_newton_rtol | float | 1e-15 | >0 <1
inner_newton_atol | float | 0.0 | >=0

[AlgoLBFGS]
bfgs_memory_size | int | 5 | >=0
use_bfgs_scaling | bool | True |
bfgs_periodic_restart | int | 0 | >=0
damped | bool | False |

[AlgoCG]
cg_method | str | DY | in(fr,pr,hs,dy,hz)
cg_periodic_restart | bool | False |
cg_periodic_its | int | 10 | >=0
cg_relative_restart | bool | False |
cg_restart_tol | float | 0.25 | >0

[MeshQuality]
tol_lower | float | 0.0 | <1 >=0
tol_upper | float | 1e-15 | <1 >0 gt(MeshQuality.tol_lower)
measure | str | skewness | in(skewness,radius_ratios,maximum_angle,condition_number)
type | str | min | in(min,avg,q,minimum,average,quantile)
quantile | float | 0.0 | >=0 <1
volume_change | float | inf | >0 >1
angle_change | float | inf | >0
remesh_iter | int | 0 | >=0

[TopologyOptimization]
angle_tol | float | 1.0 | >0
interpolation_scheme | str | volume | in(angle,volume)
normalize_topological_derivative | bool | False |
re_normalize_levelset | bool | False |
topological_derivative_is_identical | bool | False |
tol_bisection | float | 1e-4 | >=0
max_iter_bisection | int | 100 | >=0

[Output]
save_results | bool | True |
verbose | bool | False |
save_txt | bool | False |
save_state | bool | False |
save_adjoint | bool | False |
save_gradient | bool | False |
save_mesh | bool | False | needs(Mesh.gmsh_file)
result_dir | str | ./results |
precision | int | 3 | >0
time_suffix | bool | False |
single_file | bool | False |

[Debug]
remeshing | bool | False |
restart | bool | False |
"""


def _parse_table(text: str):
    scheme: dict[str, dict[str, dict]] = {}
    defaults: dict[str, dict[str, str]] = {}
    section = ""
    for raw in text.strip().splitlines():
        line = raw.strip()
        if not line:
            continue
        if line.startswith("["):
            section = line[1:-1]
            scheme[section], defaults[section] = {}, {}
            continue
        key, typ, default, cons = (t.strip() for t in line.split("|"))
        entry: dict = {"type": typ}
        for tok in cons.split():
            m = re.fullmatch(r"(\w+)\((.*)\)", tok)
            if m is None:
                entry.setdefault("range", []).append(tok)
            elif m.group(1) == "in":
                entry["options"] = m.group(2).split(",")
            elif m.group(1) == "file":
                entry["file"] = m.group(2)
            else:  # needs / gt / ge
                entry[m.group(1)] = tuple(m.group(2).split("."))
        scheme[section][key] = entry
        if default:
            defaults[section][key] = default
    scheme["DEFAULT"] = {}
    return scheme, defaults


_SCHEME, _DEFAULTS = _parse_table(_TABLE)
_LIST_CHARS = set("[].,-\"'_")


def _check_for_config_list(string: str) -> bool:
    """A python list of numbers / names: the character whitelist of ``cashocs/io/config.py:52-78``."""
    if not all(c.isdigit() or c.isalpha() or c.isspace() or c in _LIST_CHARS for c in string):
        return False
    return string[0] == "[" and string[-1] == "]"


class Config(configparser.ConfigParser):
    """Class for handling the config in cashocs (``cashocs/io/config.py:81-1065``)."""

    def __init__(self, config_file: str | None = None) -> None:
        super().__init__()
        self.config_errors: list[str] = []
        self.config_scheme = _SCHEME
        self.read_dict(_DEFAULTS)
        if config_file is not None:
            if not pathlib.Path(config_file).is_file():
                raise _exceptions.InputError(
                    "cashocs.Config", "config_file",
                    f"Could not find the specified config file {config_file}. "
                    "Please supply a path to an existing configuration file.",
                )
            self.read(config_file)

    def getlist(self, section: str, option: str, **kwargs) -> list:
        """Extract a list via JSON (``cashocs/io/config.py:761-784``)."""
        if self.config_scheme[section][option]["type"] == "list" and _check_for_config_list(self.get(section, option)):
            return json.loads(self.get(section, option, **kwargs))
        raise _exceptions.InputError("Config.getlist", "option",
                                     f"option {option} in section {section} cannot be used as list.")

    def generate_string_representation(self) -> str:
        """``cashocs/io/config.py:794-809``."""
        out = []
        for section in self.sections():
            out.append(f"[{section}]")
            out += [f"{key} = {self[section][key]}" for key in self[section]]
            out.append("")
        return "\n".join(out)

    # ------------------------------------------------------------------ validation
    def validate_config(self) -> None:
        """The reference's checks (``:786-792``), then the device layer's own restrictions."""
        self.config_errors = []
        self._check_sections()
        self._check_keys()
        self._check_supported()
        if self.config_errors:
            raise _exceptions.ConfigError(self.config_errors)

    def _check_sections(self) -> None:
        for name, section in self.items():
            if name not in self.config_scheme:
                self.config_errors.append(f"The following section is not valid: {section}\n")

    def _check_keys(self) -> None:
        for name, section in self.items():
            if name not in self.config_scheme:
                continue
            for key in section.keys():
                if key not in self.config_scheme[name]:
                    self.config_errors.append(f"Key {key} is not valid for section {name}.\n")
                else:
                    self._check_key(name, key, self.config_scheme[name][key])

    def _check_key(self, section: str, key: str, entry: dict) -> None:
        """Type, options, attributes, requirements and ordering of one key, in the reference's order
        (``:819-955``).  Range attributes read the value with ``getfloat`` like the reference does, so an
        unparsable number raises ``ValueError`` out of the validation."""
        err = self.config_errors.append
        typ = entry["type"]
        try:
            if typ == "bool":
                self.getboolean(section, key)
            elif typ == "int":
                self.getint(section, key)
            elif typ == "float":
                self.getfloat(section, key)
            elif typ == "list" and not _check_for_config_list(self.get(section, key)):
                raise ValueError
        except ValueError:
            err(f"Key {key} in section {section} has the wrong type. Required type is {typ}.\n")
        if "options" in entry and self[section][key].casefold() not in entry["options"]:
            err(f"Key {key} in section {section} has a wrong value. Possible options are {entry['options']}.\n")
        if "file" in entry:
            path = self.get(section, key)
            if not pathlib.Path(path).is_file():
                err(f"Key {key} in section {section} should point to a file, but the file does not exist.\n")
            if path.split(".")[-1] != entry["file"]:
                err(f"Key {key} in section {section} has the wrong file extension, it should end in .{entry['file']}.\n")
        rng = entry.get("range", ())
        if ">=0" in rng and self.getfloat(section, key) < 0:
            err(f"Key {key} in section {section} is negative, but it must not be.\n")
        if ">0" in rng and self.getfloat(section, key) <= 0:
            err(f"Key {key} in section {section} is non-positive, but it most be positive.\n")
        if "<1" in rng and self.getfloat(section, key) >= 1:
            err(f"Key {key} in section {section} is larger than one, but it must be smaller.\n")
        if ">1" in rng and self.getfloat(section, key) <= 1:
            err(f"Key {key} in section {section} is smaller than one, but it must be larger.\n")
        if typ == "bool" and "needs" in entry and self[section][key].casefold() == "true":
            rsec, rkey = entry["needs"]
            if not self.has_option(rsec, rkey):
                err(f"Key {key} in section {section} requires key {rkey} in section {rsec} to be present.\n")
        for rel, violated in (("gt", lambda lo, hi: lo >= hi), ("ge", lambda lo, hi: lo > hi)):
            if rel in entry:
                psec, pkey = entry[rel]
                if violated(self.getfloat(psec, pkey), self.getfloat(section, key)):
                    err(f"The value of key {key} in section {section} is smaller than the value of key {pkey} "
                        f"in section {psec}, but it should be larger.\n")

    def _check_supported(self) -> None:
        """Keys that select subsystems outside the B200 hot path must keep their defaults;
        the reference would run those subsystems, the device layer refuses loudly instead of ignoring the key."""

        def flag(section: str, key: str, why: str) -> None:
            self.config_errors.append(
                f"Key {key} in section {section} is outside the B200 hot path (SURVEY.md 8a): {why}.\n")

        def boolean(section, key):
            try:
                return self.getboolean(section, key)
            except ValueError:
                return None

        def number(section, key):
            try:
                return self.getfloat(section, key)
            except ValueError:
                return None

        if boolean("ShapeGradient", "use_distance_mu") and self.get("ShapeGradient", "distance_method") != "eikonal":
            flag("ShapeGradient", "distance_method", "the device path computes boundary distances with eikonal")
        if self.get("ShapeGradient", "shape_volume_fix").replace(" ", "") != "[]":
            flag("ShapeGradient", "shape_volume_fix", "it must be empty")
        if boolean("Mesh", "remesh"):
            flag("Mesh", "remesh", "remeshing needs the gmsh executable, it must be False")
        if (number("MeshQuality", "remesh_iter") or 0) > 0:
            flag("MeshQuality", "remesh_iter", "remeshing needs the gmsh executable, it must be 0")
        if self.get("StateSystem", "backend").casefold() == "petsc":
            flag("StateSystem", "backend", "SNES is not built, use the cashocs backend")


def load_config(path: str) -> Config:
    """Loads a config object from a config file (``cashocs/io/config.py:37-49``)."""
    return Config(path)
