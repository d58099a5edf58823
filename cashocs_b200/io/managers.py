"""Output of the optimisation loop (``cashocs/io/managers.py``, ``cashocs/io/output.py``).

Host-side only: the optimisation state (a handful of floats read back once per iteration) is written to the
console / log, to ``history.txt`` and ``history.json``; meshes are checkpointed as gmsh files through
:func:`cashocs_b200.io.mesh.write_out_mesh`.  The XDMF / HDF5 checkpoints of states, adjoints and gradients
(``XDMFFileManager``, ``managers.py:414-716``) need an HDF5 library, which this image does not have:
``save_state`` / ``save_adjoint`` / ``save_gradient`` write the nodal arrays as ``.npy`` files next to where the
reference puts its ``.xdmf`` / ``.h5`` pairs and say so in the log.
"""

from __future__ import annotations

import datetime
import json
import os

import numpy as np

from cashocs_b200 import log

# key of the optimisation state -> column title (``managers.py:37-47``)
output_mapping = {
    "iteration": "iter",
    "objective_value": "cost function",
    "relative_norm": "rel. grad. norm",
    "gradient_norm": "abs. grad. norm",
    "mesh_quality": "mesh qlty",
    "angle": "angle",
    "stepsize": "step size",
}


def generate_summary_str(state: dict, precision: int) -> str:
    """``managers.py:50-86``."""
    out = ["\n", "Optimization was successful.\n", "Statistics:\n", f"    total iterations: {state['iteration']:4d}\n"]
    for key, title in output_mapping.items():
        if key in state and key != "iteration":
            out.append(f"    final {title}: {state[key]:.{precision}e}\n")
    if "no_state_solves" in state:
        out.append(f"    total number of state systems solved: {state['no_state_solves']:4d}\n")
    if "no_adjoint_solves" in state:
        out.append(f"    total number of adjoint systems solved: {state['no_adjoint_solves']:4d}\n")
    return "".join(out)


def generate_output_str(state: dict, precision: int) -> str:
    """One table row, preceded by the column titles every tenth iteration (``managers.py:89-142``)."""
    iteration = state["iteration"]
    titles, values = [], []
    if iteration % 10 == 0:
        titles.append("\n")
    for key, title in output_mapping.items():
        if key not in state:
            continue
        if key == "iteration":
            if iteration % 10 == 0:
                titles.append("iter,  ")
            values.append(f"{iteration:4d},  ")
            continue
        width = max(len(f"{title},  "), len(f"{state[key]:.{precision}e},  "))
        if iteration % 10 == 0:
            titles.append(f"{title},  ".rjust(width))
        values.append(f"{state[key]:>{width - 3}.{precision}e},  ")
    if iteration % 10 == 0:
        titles.append("\n\n")
    if iteration == 0:
        values.append("\n")
    return "".join(titles) + "".join(values)


class IOManager:
    def __init__(self, problem, result_dir: str) -> None:
        self.problem = problem
        self.config = problem.config
        self.result_dir = result_dir
        dist = getattr(problem.mesh, "dist", None)
        self.rank = dist.rank if dist is not None else 0
        self.world = dist.world if dist is not None else 1

    def barrier(self) -> None:
        if self.world > 1:
            import torch.distributed as td

            td.barrier()

    def output(self, state: dict) -> None:
        pass

    def output_summary(self, state: dict) -> None:
        pass

    def post_process(self, state: dict) -> None:
        pass


class ResultManager(IOManager):
    """Optimisation history -> ``history.json`` (``managers.py:190-233``)."""

    def __init__(self, problem, result_dir: str) -> None:
        super().__init__(problem, result_dir)
        self.save_results = self.config.getboolean("Output", "save_results")
        self.output_dict: dict[str, list] = {}

    def output(self, state: dict) -> None:
        for key, value in state.items():
            self.output_dict.setdefault(key, []).append(value)

    def post_process(self, state: dict) -> None:
        if self.save_results and self.rank == 0:
            with open(f"{self.result_dir}/history.json", "w", encoding="utf-8") as fh:
                json.dump(self.output_dict, fh, indent=4)
        self.barrier()


class ConsoleManager(IOManager):
    """``managers.py:236-285``: ``print`` with ``verbose``, the INFO log otherwise."""

    def __init__(self, problem, result_dir: str, verbose: bool = False) -> None:
        super().__init__(problem, result_dir)
        self.verbose = verbose
        self.precision = self.config.getint("Output", "precision")

    def _emit(self, message: str) -> None:
        if self.verbose:
            if self.rank == 0:
                print(message, flush=True)
        elif self.rank == 0:
            log.info(message)

    def output(self, state: dict) -> None:
        self._emit(generate_output_str(state, self.precision))

    def output_summary(self, state: dict) -> None:
        self._emit(generate_summary_str(state, self.precision))


class FileManager(IOManager):
    """``history.txt`` (``managers.py:288-330``)."""

    def __init__(self, problem, result_dir: str) -> None:
        super().__init__(problem, result_dir)
        self.precision = self.config.getint("Output", "precision")

    def output(self, state: dict) -> None:
        if self.rank == 0:
            with open(f"{self.result_dir}/history.txt", "w" if state["iteration"] == 0 else "a", encoding="utf-8") as fh:
                fh.write(f"{generate_output_str(state, self.precision)}\n")
        self.barrier()

    def output_summary(self, state: dict) -> None:
        if self.rank == 0:
            with open(f"{self.result_dir}/history.txt", "a", encoding="utf-8") as fh:
                fh.write(generate_summary_str(state, self.precision))
        self.barrier()


class MeshManager(IOManager):
    """Mesh checkpoints as gmsh files (``managers.py:373-411``); needs the gmsh file the mesh came from."""

    def _gmsh_file(self) -> str:
        path = getattr(self.problem.mesh, "gmsh_file", None)
        if not path and self.config.has_option("Mesh", "gmsh_file"):
            path = self.config.get("Mesh", "gmsh_file")
        if not path:
            from cashocs_b200 import _exceptions

            raise _exceptions.InputError("cashocs_b200.io.managers.MeshManager", "save_mesh",
                                         "save_mesh needs a mesh imported from a gmsh file ([Mesh] gmsh_file)")
        return path

    def output(self, state: dict) -> None:
        from cashocs_b200.io import mesh as iomesh

        iomesh.write_out_mesh(self.problem.mesh, self._gmsh_file(),
                              f"{self.result_dir}/checkpoints/mesh/mesh_{int(state['iteration'])}.msh")

    def post_process(self, state: dict) -> None:
        from cashocs_b200.io import mesh as iomesh

        iomesh.write_out_mesh(self.problem.mesh, self._gmsh_file(), f"{self.result_dir}/optimized_mesh.msh")


class ArrayFileManager(IOManager):
    """Stand-in for ``XDMFFileManager`` (``managers.py:414-716``) without HDF5: nodal arrays of the owned dofs as
    ``checkpoints/{state,adjoint,gradient}_{i}_{iteration}.npy`` (one file per rank on partitioned meshes)."""

    def __init__(self, problem, result_dir: str) -> None:
        super().__init__(problem, result_dir)
        self.save_state = self.config.getboolean("Output", "save_state")
        self.save_adjoint = self.config.getboolean("Output", "save_adjoint")
        self.save_gradient = self.config.getboolean("Output", "save_gradient")
        log.warning("no HDF5 library in this environment: save_state / save_adjoint / save_gradient write .npy "
                    "arrays instead of XDMF + HDF5 files")

    def _save(self, name: str, functions, iteration: int) -> None:
        suffix = f"_rank{self.rank}" if self.world > 1 else ""
        for i, f in enumerate(functions):
            np.save(f"{self.result_dir}/checkpoints/{name}_{i}_{iteration}{suffix}.npy",
                    f.vector.detach().cpu().numpy())

    def output(self, state: dict) -> None:
        it = int(state["iteration"])
        if self.save_state:
            self._save("state", self.problem.states, it)
            if getattr(self.problem, "problem_type", "shape") == "control":
                self._save("control", self.problem.controls, it)
        if self.save_adjoint:
            self._save("adjoint", self.problem.adjoints, it)
        if self.save_gradient:
            self._save("gradient", [self.problem.gradient], it)


class OutputManager:
    """``cashocs/io/output.py:32-118``."""

    def __init__(self, problem) -> None:
        self.problem = problem
        cfg = problem.config
        self.config = cfg
        self.result_dir = cfg.get("Output", "result_dir").rstrip("/")
        self._silent = False
        if cfg.getboolean("Output", "time_suffix"):
            t = datetime.datetime.now()
            self.suffix = f"{t.year}_{t.month}_{t.day}_{t.hour}_{t.minute}_{t.second}"
            self.result_dir = f"{self.result_dir}_{self.suffix}"
        verbose = cfg.getboolean("Output", "verbose")
        save_txt = cfg.getboolean("Output", "save_txt")
        save_results = cfg.getboolean("Output", "save_results")
        save_fields = any(cfg.getboolean("Output", k) for k in ("save_state", "save_adjoint", "save_gradient"))
        save_mesh = cfg.getboolean("Output", "save_mesh")
        dist = getattr(problem.mesh, "dist", None)
        rank0 = dist is None or dist.rank == 0
        if rank0 and (save_txt or save_results or save_fields or save_mesh):
            os.makedirs(self.result_dir, exist_ok=True)
        if save_fields or save_mesh:
            os.makedirs(f"{self.result_dir}/checkpoints", exist_ok=True)
        self.managers: list[IOManager] = [ConsoleManager(problem, self.result_dir, verbose=verbose)]
        if save_txt:
            self.managers.append(FileManager(problem, self.result_dir))
        if save_fields:
            self.managers.append(ArrayFileManager(problem, self.result_dir))
        result_manager = ResultManager(problem, self.result_dir)
        self.output_dict = result_manager.output_dict
        self.managers.append(result_manager)
        if save_mesh:
            self.managers.append(MeshManager(problem, self.result_dir))

    def output(self, state: dict) -> None:
        if not self._silent:
            for m in self.managers:
                m.output(state)

    def output_summary(self, state: dict) -> None:
        if not self._silent:
            for m in self.managers:
                m.output_summary(state)

    def post_process(self, state: dict) -> None:
        if not self._silent:
            for m in self.managers:
                m.post_process(state)
