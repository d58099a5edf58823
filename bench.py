#!/usr/bin/env python
"""Benchmark: shape-gradient evaluations per second on the synthetic unit-cube workload.

One *step* = one full shape-gradient evaluation: [Lame-mu solve] +
elasticity matrix assembly + state assemble+solve + adjoint assemble+solve + shape-derivative
vector assembly + elasticity Riesz solve + shape BCs + g^T A g + one trial mesh move
(a-priori det(I+grad V) test, coordinate update, intersection test, mesh quality) + revert.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mesh-n CELLS_PER_EDGE] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); the mesh is partitioned into vertex slabs,
halo exchange + all-reduces go over NCCL; `value` is whole-mesh evaluations per second.
Prints ONE JSON line (rank 0).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "shape_gradient_evals_per_s"
UNIT = "evals/s"
# the reference's iterative default (cashocs/_utils/linalg.py:44-52: cg + hypre/boomeramg) at rtol 1e-10;
# `--pc jacobi` selects the one-level variant of SURVEY.md 8d
KSP = {"ksp_type": "cg", "pc_type": "hypre", "pc_hypre_type": "boomeramg", "ksp_rtol": 1e-10, "ksp_atol": 1e-30,
       "ksp_max_it": 20000}
GRADIENT_KSP_EXTRA: dict = {}  # options of the elasticity Riesz solve only (vertex-blocked 3 x 3 operator)
SHAPE = dict(lambda_lame=1.428571428571429, damping_factor=0.2, mu=0.35714285714285715)
ALL_MARKERS = [1, 2, 3, 4, 5, 6]


def workload_name(n: int, n_gpus: int, workload: str = "cube") -> str:
    if workload == "square":
        return (f"unit-square regular_mesh({n}): {2 * n**2} triangles / {(n + 1) ** 2} vertices, P1 Poisson state+adjoint, "
                f"vector-P1 elasticity Riesz shape gradient, cg+{KSP['pc_type']} rtol 1e-10, trial move + quality; {n_gpus} GPU(s)")
    tets = 6 * n**3
    if workload == "control":
        return (f"unit-cube regular_mesh({n},1,1,1): {tets} tets / {(n + 1) ** 3} vertices, distributed-control Poisson problem: "
                f"state + adjoint (cg+{KSP['pc_type']} rtol 1e-10) + L2 mass-matrix Riesz gradient (cg+jacobi); {n_gpus} GPU(s)")
    extra = ", Lame mu from the auxiliary Poisson solve (mu_def 5e2 / mu_fix 1)" if workload == "cube-mu" else ""
    return (f"unit-cube regular_mesh({n},1,1,1): {tets} tets / {(n + 1) ** 3} vertices, P1 Poisson state+adjoint, "
            f"vector-P1 elasticity Riesz shape gradient{extra}, cg+{KSP['pc_type']} rtol 1e-10, trial move + quality; {n_gpus} GPU(s)")


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0) -> None:
        self.gpu_index = gpu_index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), smax.append(float(f[2])), power.append(float(f[3]))
            except ValueError:
                continue
            for name, flag in zip(names, f[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(smax)), "power_w_max": float(np.max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- oracle arm / cpu baseline
_ORACLE_THREADS = 1


def workload_config(n: int, state_degree: int = 1, workload: str = "cube") -> dict:
    """The part of `config` that names the workload -- identical in the B200 arm and in the reference arm."""
    cells, verts = (2 * n**2, (n + 1) ** 2) if workload == "square" else (6 * n**3, (n + 1) ** 3)
    return {"workload": workload_name(n, 1, workload).rsplit(";", 1)[0] + (" [P2 state/adjoint]" if state_degree == 2 else ""),
            "workload_id": workload, "n": n, "tets": cells, "vertices": verts, "state_degree": state_degree,
            "ksp": "cg + pc_type hypre (algebraic multigrid), ksp_rtol 1e-10, ksp_atol 1e-30"}


def make_oracle_problem(n_sample: int, workload: str = "cube"):
    from oracle import meshes, pipeline

    # the reference's iterative default is cg + hypre/boomeramg (cashocs/_utils/linalg.py:44-52); hypre is not
    # installable here, the oracle's multigrid of the same class (oracle/amg.py, scipy) stands in for it, so the
    # CPU arm has mesh-independent iteration counts like the reference and like the device path
    ksp = {"ksp_type": "cg", "pc_type": "hypre", "ksp_rtol": 1e-10, "ksp_atol": 1e-30, "ksp_max_it": 20000}
    if workload == "control":
        om = meshes.regular_mesh(n_sample, 1.0, 1.0, 1.0)
        c = om.coords
        yd = np.sin(2 * np.pi * c[:, 0]) * np.sin(2 * np.pi * c[:, 1]) * np.sin(2 * np.pi * c[:, 2])
        u = np.random.RandomState(300696).rand(om.num_vertices)
        return pipeline.PoissonControlProblem(om, yd, u, alpha=1e-6, bc_markers=tuple(ALL_MARKERS), state_ksp=ksp,
                                              adjoint_ksp=ksp, riesz_ksp={"ksp_type": "cg", "pc_type": "jacobi",
                                                                          "ksp_rtol": 1e-10, "ksp_atol": 1e-30,
                                                                          "ksp_max_it": 1000}), om.num_cells
    dim = 2 if workload == "square" else 3
    om = meshes.regular_mesh(n_sample) if dim == 2 else meshes.regular_mesh(n_sample, 1.0, 1.0, 1.0)
    markers = tuple(ALL_MARKERS[: 2 * dim])
    # the same evaluation as the B200 arm, intersection test included (tree-based candidate search, oracle/quality.py)
    if workload == "cube-mu":
        cfg = pipeline.ShapeConfig(shape_bdry_def=(2,), shape_bdry_fix=(1, 3, 4, 5, 6), lambda_lame=SHAPE["lambda_lame"],
                                   damping_factor=SHAPE["damping_factor"], mu_def=500.0, mu_fix=1.0,
                                   test_for_intersections=True)
    else:
        cfg = pipeline.ShapeConfig(shape_bdry_def=markers, lambda_lame=SHAPE["lambda_lame"],
                                   damping_factor=SHAPE["damping_factor"], mu_def=SHAPE["mu"], mu_fix=SHAPE["mu"],
                                   test_for_intersections=True)
    return pipeline.PoissonShapeProblem(om, bc_markers=markers, config=cfg, state_ksp=ksp, adjoint_ksp=ksp,
                                        gradient_ksp=ksp, mu_ksp=ksp), om.num_cells


def oracle_eval_seconds(n_sample: int, prob=None, workload: str = "cube"):
    """Time ONE evaluation of the CPU oracle on an n_sample mesh; returns (s, its, cells, problem)."""
    global _ORACLE_THREADS
    if prob is None:
        prob, _ = make_oracle_problem(n_sample, workload)
    t0 = time.perf_counter()
    c0 = time.process_time()
    if workload == "control":
        prob.compute_gradient()
    else:
        prob.update_scalar_product()
        g = prob.compute_shape_gradient()
        h = 1.0 / n_sample
        scale = 0.1 * h / max(float(np.max(np.abs(g))), 1e-300)
        if prob.move_mesh(-scale * g):
            prob.revert_transformation()
    wall = time.perf_counter() - t0
    # threads actually kept busy: CPU time of the process / wall time (BLAS threads count, idle cores do not)
    _ORACLE_THREADS = max(1, int(round((time.process_time() - c0) / wall)))
    return wall, dict(prob.its), prob.cells.shape[0], prob


def _cells_of(n: int, workload: str) -> int:
    return 2 * n**2 if workload == "square" else 6 * n**3


def sample_size_for(budget_s: float, n_full: int, workload: str = "cube") -> tuple[int, float]:
    """Largest sample mesh whose evaluation fits `budget_s` seconds on this host, from a calibration evaluation (the
    evaluation is multigrid-preconditioned: its time grows linearly in the number of cells)."""
    n_cal = 64 if workload == "square" else 16
    secs, _, nc, _ = oracle_eval_seconds(n_cal, workload=workload)
    per_cell = secs / nc
    cells = budget_s / per_cell
    n = int((cells / 2.0) ** 0.5) if workload == "square" else int((cells / 6.0) ** (1.0 / 3.0))
    return max(n_cal, min(n, n_full)), secs


def sample_string(n_sample, nc, secs, its, n_full, workload="cube"):
    frac = nc / float(_cells_of(n_full, workload))
    return (f"numpy/scipy oracle (CPU restatement; FEniCS/PETSc cannot be installed here), MEASURED: one full evaluation "
            f"(assembly, cg + smoothed-aggregation AMG standing in for hypre, Krylov its {its}, det test, move, intersection "
            f"test, quality) on regular_mesh({n_sample}) of the workload's mesh family = {nc} cells took {secs:.3f} s; that sample is {frac:.4f} of the "
            f"{_cells_of(n_full, workload)}-cell workload, `value` = {frac:.4f} evaluations of the workload per {secs:.3f} s (EXTRAPOLATION: "
            f"linear in the number of cells; numpy/BLAS may use all {os.cpu_count()} host threads, scipy's sparse kernels are "
            f"single-threaded, `cores` = CPU time / wall time of the sample)")


def cpu_baseline(n_full: int, budget_s: float = 20.0, workload: str = "cube") -> dict:
    n_sample, _ = sample_size_for(budget_s, n_full, workload)
    secs, its, nc, prob = oracle_eval_seconds(n_sample, workload=workload)
    frac = nc / float(_cells_of(n_full, workload))
    out = {"value": frac / secs, "unit": UNIT, "cores": _ORACLE_THREADS, "kind": "port",
           "sample": sample_string(n_sample, nc, secs, its, n_full, workload), "sample_n": n_sample, "sample_seconds": secs,
           "sample_fraction_of_workload": frac}
    out["_oracle_gnorm2"] = float(getattr(prob, "gradient_norm_squared", float("nan")))
    return out


def device_parity_number(n_sample: int, workload: str, oracle_gnorm2: float) -> dict:
    """The bench line's parity number: g^T A g of the device pipeline on the CPU leg's sample mesh, with the solver
    settings that are being timed, next to the oracle's value for the same mesh (rtol 1e-10 on both sides)."""
    sop = build_problem(n_sample, None, 1, workload)
    out = evaluation_step(sop, 1.0 / n_sample)
    dev = float(out[0])
    return {"n": n_sample, "gradient_norm_squared_device": dev, "gradient_norm_squared_oracle": oracle_gnorm2,
            "rel_diff": abs(dev - oracle_gnorm2) / max(abs(oracle_gnorm2), 1e-300)}


def run_reference_arm(args) -> None:
    """The CPU arm: every step is ONE measured evaluation of the oracle on a bounded sample of the workload (a smaller
    cube of the same mesh family, sized so that warm-up + steps end within a few minutes on this host);
    `ms_per_step` is the measured time of such a step, `value` the workload throughput it implies."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total_budget = float(args.cpu_budget)  # seconds for all warm-up + timed steps
    per_step = min(total_budget / max(args.warmup + args.steps, 1), 40.0)
    n_sample, calib = sample_size_for(per_step, args.n, args.workload)
    prob, nc = make_oracle_problem(n_sample, args.workload)
    times = []
    its = {}
    for i in range(args.warmup + args.steps):
        secs, its, nc, prob = oracle_eval_seconds(n_sample, prob, args.workload)
        if i >= args.warmup:
            times.append(secs)
    frac = nc / float(_cells_of(args.n, args.workload))
    step_s = float(np.mean(times))
    value = frac / step_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.n, args.state_degree, args.workload),
        "sample": {"n": n_sample, "tets": nc, "fraction_of_workload": frac, "measured_seconds_per_step": step_s,
                   "krylov_its_state_adjoint_elasticity": [its.get("state"), its.get("adjoint"), its.get("gradient")]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": _ORACLE_THREADS, "kind": "port",
                         "sample": sample_string(n_sample, nc, step_s, its, args.n, args.workload)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- B200 arm
WORKLOADS = {
    "cube": "BASELINE configs[1]/[2]: unit cube, every boundary deformable, constant Lame mu (the default)",
    "cube-mu": "same mesh with the auxiliary mu-Poisson solve: mu_def = 5e2 on face x = 1, mu_fix = 1 on the fixed faces "
               "(shape_form_handler.py:188-235; values of demos/documented/shape_optimization/shape_stokes/config.ini)",
    "square": "unit square regular_mesh(n): 2 n^2 triangles, same pipeline in 2-D",
    "control": "BASELINE configs[3]: distributed-control Poisson problem on the unit cube, one reduced-gradient "
               "evaluation = state + adjoint + mass-matrix Riesz solve (tests/test_pde_problems.py:59-84)",
}


def build_problem(n: int, dist=None, state_degree: int = 1, workload: str = "cube"):
    import cashocs_b200 as cb

    dim = 2 if workload == "square" else 3
    lengths = (1.0, 1.0) if dim == 2 else (1.0, 1.0, 1.0)
    if dist is None:
        mesh = cb.regular_mesh(n) if dim == 2 else cb.regular_mesh(n, 1.0, 1.0, 1.0)
    else:
        from cashocs_b200 import distributed

        mesh = distributed.distributed_regular_mesh(n, dist, lengths=lengths)
    markers = ALL_MARKERS[: 2 * dim]
    cfg = cb.Config()
    cfg.set("StateSystem", "is_linear", "True")
    if workload == "control":
        import torch

        y, p, u, y_d = (cb.Function(mesh) for _ in range(4))
        c = mesh.coords
        y_d.vector.copy_(torch.sin(2 * np.pi * c[:, 0]) * torch.sin(2 * np.pi * c[:, 1]) * torch.sin(2 * np.pi * c[:, 2]))
        gen = np.random.RandomState(300696)  # tests/conftest.py:47-49
        gid = mesh.global_vertex_ids.cpu().numpy() if dist is not None else np.arange(mesh.num_vertices)
        u.vector.copy_(torch.from_numpy(gen.rand((n + 1) ** 3)[gid]).to(mesh.device))
        mksp = dict(KSP)
        mksp["ksp_profile_every"] = 0
        ocp = cb.OptimalControlProblem(cb.PoissonForm(source=u), cb.create_dirichlet_bcs(0.0, markers),
                                       cb.IntegralFunctional(c_track=1.0, desired_state=y_d, c_control=1e-6), y, u, p,
                                       config=cfg, ksp_options=dict(KSP, ksp_profile_every=8), adjoint_ksp_options=dict(KSP),
                                       gradient_ksp_options={"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-10,
                                                             "ksp_atol": 1e-30, "ksp_max_it": 1000})
        return ocp
    x = cb.spatial_coordinate(dim)
    f = 2.5 * (x[0] + 0.4 - x[1] ** 2) ** 2 + x[0] ** 2 + x[1] ** 2 - 1
    cfg.set("ShapeGradient", "lambda_lame", repr(SHAPE["lambda_lame"]))
    cfg.set("ShapeGradient", "damping_factor", repr(SHAPE["damping_factor"]))
    if workload == "cube-mu":
        cfg.set("ShapeGradient", "shape_bdry_def", "[2]")
        cfg.set("ShapeGradient", "shape_bdry_fix", "[1, 3, 4, 5, 6]")
        cfg.set("ShapeGradient", "mu_fix", "1.0")
        cfg.set("ShapeGradient", "mu_def", "500.0")
    else:
        cfg.set("ShapeGradient", "shape_bdry_def", json.dumps(markers))
        cfg.set("ShapeGradient", "mu_fix", repr(SHAPE["mu"]))
        cfg.set("ShapeGradient", "mu_def", repr(SHAPE["mu"]))
    u, p = cb.Function(mesh, degree=state_degree), cb.Function(mesh, degree=state_degree)
    gksp = dict(KSP)
    gksp.update(GRADIENT_KSP_EXTRA)
    gksp["ksp_profile_every"] = 8
    sop = cb.ShapeOptimizationProblem(cb.PoissonForm(source=f), cb.create_dirichlet_bcs(0.0, markers),
                                      cb.IntegralFunctional(c_state=1.0), u, p, config=cfg, ksp_options=dict(KSP),
                                      adjoint_ksp_options=dict(KSP), gradient_ksp_options=gksp)
    return sop


def evaluation_step(sop, h: float):
    """One shape-gradient evaluation + one trial move (then revert, so every step sees the same mesh); for the control
    workload one reduced-gradient evaluation."""
    sop.state_problem.has_solution = False
    sop.adjoint_problem.has_solution = False
    sop.gradient_problem.has_solution = False
    if not hasattr(sop, "mesh_handler") or sop.mesh_handler is None:
        sop.compute_gradient()
        g = sop.gradient.vector
        return sop.form_handler.scalar_product(g, g), True, None
    sop.form_handler.update_scalar_product()
    g = sop.compute_shape_gradient()[0].vector
    gnorm2 = sop.gradient_problem.gradient_norm_squared
    ginf = float(g.abs().max().item())
    if sop.mesh.dist is not None:
        ginf = sop.mesh.dist.allreduce_scalar(ginf, "max")
    step = -0.1 * h / max(ginf, 1e-300)
    moved = sop.mesh_handler.move_mesh(g, step=step)
    quality = sop.mesh_handler.current_mesh_quality
    if moved:
        sop.mesh_handler.revert_transformation()
    return gnorm2, moved, quality


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mesh-n", dest="n", type=int, default=119, help="cells per cube edge (119 -> 10.1M tets, BASELINE configs[1])")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cube", choices=sorted(WORKLOADS), help="; ".join(f"{k}: {v}" for k, v in WORKLOADS.items()))
    ap.add_argument("--pc", default="hypre", choices=["hypre", "jacobi"],
                    help="pc_type of the three Krylov solves (hypre -> device smoothed-aggregation AMG on 1 GPU)")
    ap.add_argument("--state-degree", type=int, default=1, choices=[1, 2],
                    help="CG degree of the state / adjoint space (2 = BASELINE configs[4])")
    ap.add_argument("--graph", type=int, default=0, choices=[-1, 0, 1],
                    help="CUDA-graph replay of the PCG iteration: 0 / 1 on (default), -1 eager launches only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=300.0,
                    help="--impl reference: seconds of CPU work for all warm-up + timed steps (sizes the sample mesh)")
    ap.add_argument("--rbm", default="auto", choices=["auto", "on", "off"],
                    help="rigid-body coarse correction in the elasticity AMG preconditioner (pc_amg_rigid_body_correction); "
                         "auto = on (the library default, also on partitioned meshes)")
    ap.add_argument("--nullspace", default="auto", choices=["auto", "translations", "rigid_body"],
                    help="near-null-space carried by the elasticity AMG hierarchy (pc_amg_nullspace)")
    args = ap.parse_args()
    KSP["pc_type"] = args.pc
    if args.nullspace != "auto" and args.pc == "hypre":
        GRADIENT_KSP_EXTRA["pc_amg_nullspace"] = args.nullspace
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    use_rbm = args.pc == "hypre" and args.rbm != "off"
    if args.pc == "hypre":
        KSP["pc_amg_rigid_body_correction"] = use_rbm  # explicit either way (the library default is on for 1 GPU)
    if args.graph != 0:
        KSP["ksp_use_graph"] = args.graph
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch

    import cashocs_b200 as cb
    from cashocs_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dist = None
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    if world > 1:
        import torch.distributed as td

        from cashocs_b200 import distributed

        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = distributed.DistContext(rank, world)

    lib = _lib.load()
    t_setup = time.perf_counter()
    sop = build_problem(args.n, dist, args.state_degree, args.workload)
    mesh = sop.mesh
    h = 1.0 / args.n
    is_control = args.workload == "control"
    prof_problem = sop.state_problem if is_control else sop.gradient_problem  # whose SpMV launches are event-timed
    prof_bs = 1 if is_control else mesh.dim

    # pinned host mirrors: the reference keeps mesh.coordinates() on the host (deformations.py:73)
    # (control workload: the input of an evaluation is the control vector)
    e2e_input = sop.controls[0].vector if is_control else mesh.coords
    coords_host = e2e_input.cpu().pin_memory()
    grad_host = torch.empty_like(sop.gradient.vector, device="cpu").pin_memory()

    def barrier():
        if dist is not None:
            import torch.distributed as td

            td.barrier()
        torch.cuda.synchronize()

    def timed(nsteps: int, e2e: bool):
        its_acc = []
        spmv_ms = spmv_prof = 0
        barrier()
        l0 = lib.cb_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(nsteps):
            if e2e:
                e2e_input.copy_(coords_host, non_blocking=True)
            out = evaluation_step(sop, h)
            if e2e:
                grad_host.copy_(sop.gradient.vector, non_blocking=True)
                torch.cuda.current_stream().synchronize()
            r = prof_problem.linear_solver.last_result
            spmv_ms += r.spmv_ms
            spmv_prof += r.spmv_profiled
            its_acc.append((sop.state_problem.iterations[0], sop.adjoint_problem.iterations[0], sop.gradient_problem.iterations))
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if dist is not None:
            ms = dist.allreduce_scalar(ms, "max")
        return ms, lib.cb_launch_count() - l0, its_acc, spmv_ms, spmv_prof, out

    timed(max(args.warmup, 3), False)
    setup_s = time.perf_counter() - t_setup  # mesh + pattern + symbolic AMG + warm-up steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, launches, its, spmv_ms, spmv_prof, out = timed(args.steps, False)
    ms_e2e, _, _, _, _, _ = timed(args.steps, True)
    clocks = sampler.stop() if rank == 0 else {}

    value = args.steps / (ms * 1e-3)
    e2e_value = args.steps / (ms_e2e * 1e-3)

    # roofline of the dominant kernel: the block-CSR (3x3) SpMV of the elasticity PCG
    nnzb_total = mesh.nnz if dist is None else int(dist.allreduce_scalar(float(mesh.nnz), "sum"))
    nnzb_rank = mesh.nnz
    nv_rank = mesh.n_owned_vertices
    # (8 bs^2 + 4) nnzb + 16 bs N + 4 N  (3x3 BSR: 76 nnzb + 52 N; scalar CSR: 12 nnz + 20 N)
    alg_bytes = (8.0 * prof_bs * prof_bs + 4.0) * nnzb_rank + 16.0 * prof_bs * nv_rank + 4.0 * nv_rank
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # DRAM traffic of the same kernel from the committed `ncu --set full` capture (per launch; only
    # valid for the workload it was captured on: n = 119 on one GPU)
    traffic = None
    cap = os.path.join(ROOT, "profiles", "r02_ncu_full_k_cg_spmv3.csv")
    if args.n == 119 and world == 1 and args.workload == "cube" and os.path.exists(cap):
        rd = wr = None
        for row in open(cap):
            f = row.strip().split(",")
            if f[0] == "dram__bytes_read.sum":
                rd = float(f[2]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[1]]
            if f[0] == "dram__bytes_write.sum":
                wr = float(f[2]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[1]]
        if rd is not None and wr is not None:
            traffic = rd + wr
    roofline = None
    if spmv_prof > 0:
        avg_ms = spmv_ms / spmv_prof
        achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": f"k_cg_spmv<{prof_bs}> ({'state' if is_control else 'elasticity'} PCG: w = A p, fused p.w)",
                    "achieved": achieved,
                    "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch",
                    "traffic_source": "profiles/r02_ncu_full_k_cg_spmv3.csv (dram__bytes_read.sum + dram__bytes_write.sum, bytes per launch)" if traffic else None,
                    "avg_launch_ms": avg_ms, "launches_profiled": spmv_prof, "algorithmic_bytes_per_launch": alg_bytes,
                    # fine-level launches over A per PCG iteration: w = A p and the residual after pre-smoothing
                    # (the post-smoothing residual goes through A*P, k_spmv_post)
                    "spmv_share_of_step": (avg_ms * sum(i[2] for i in its) *
                                           (2 if args.pc == "hypre" else 1)) / ms}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args.n, args.state_degree, args.workload), **{
                "state_dofs": int(sop.states[0].vector.numel()),
                "block_nnz": nnzb_total, "krylov_its_state_adjoint_elasticity": list(its[-1]),
                "ksp_options": {k: v for k, v in KSP.items()}, "gradient_ksp_extra": dict(GRADIENT_KSP_EXTRA),
                "pc_on_device": (("smoothed-aggregation AMG (fp32 copy of the fine operator for the residual pass, "
                                  "degree-2 smoothing on the coarse levels of vector operators)" +
                                  (" + rigid-body coarse correction for the elasticity system" if (use_rbm and not is_control) else ""))
                                 if args.pc == "hypre" else "jacobi"),
                "l2_policy": "inputs larger than L2 (elasticity matrix %.2f GB, L2 126 MB)" % (76e-9 * nnzb_rank),
                "parallelism": "1 GPU" if dist is None else (
                    f"mesh partitioned into {world} vertex slabs, " +
                    ("halo exchange + Krylov all-reduces one-sided over NVLink peer memory (NCCL for the AMG agglomeration)"
                     if dist.p2p else "NCCL halo + allreduce"))}),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(coords_host.numel() * 8),
                    "d2h_bytes_per_step": int(grad_host.numel() * 8), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "check": {"gradient_norm_squared": out[0], "move_accepted": bool(out[1]), "mesh_quality": out[2]},
            "hbm_peak_gb_rank0": round(torch.cuda.max_memory_allocated() / 1e9, 2),
            "setup_s": round(setup_s, 1),
        }
        if not args.no_cpu_baseline and args.gpus == 1 and args.state_degree == 1:
            base = cpu_baseline(args.n, workload=args.workload)
            oracle_gnorm2 = base.pop("_oracle_gnorm2")
            line["cpu_baseline"] = base
            line["check"]["parity"] = device_parity_number(base["sample_n"], args.workload, oracle_gnorm2)
        print(json.dumps(line))
    if dist is not None:
        import torch.distributed as td

        barrier()
        dist.destroy()
        td.destroy_process_group()


if __name__ == "__main__":
    main()
