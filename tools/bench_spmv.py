#!/usr/bin/env python
"""Micro-benchmark of the block-CSR SpMV on the elasticity matrix of the bench mesh (CUDA events, L2
flushed between launches by the 1.94 GB matrix itself): GB/s of algorithmic bytes."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import cashocs_b200 as cb  # noqa: E402
from cashocs_b200._utils import linalg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mesh-n", dest="n", type=int, default=119)
ap.add_argument("--reps", type=int, default=40)
args = ap.parse_args()
mesh = cb.regular_mesh(args.n, 1.0, 1.0, 1.0)
mesh.build_pattern()
mu = torch.full((mesh.num_vertices,), 0.357, dtype=torch.float64, device="cuda")
E = linalg.assemble_elasticity_matrix(mesh, mu, 1.43, 0.2)
A, _ = linalg.assemble_scalar_system(mesh, 1.0, 1.0, rhs=False)
for name, M, f32 in (("elasticity bs=3", E, False), ("scalar bs=1", A, False), ("elasticity bs=3 fp32 values", E, True),
                     ("scalar bs=1 fp32 values", A, True)):
    bs = M.bs
    mult = M.mult_f32 if f32 else M.mult
    x = torch.rand(M.ncols * bs, dtype=torch.float64, device="cuda")
    y = torch.empty(M.nrows * bs, dtype=torch.float64, device="cuda")
    for _ in range(5):
        mult(x, y)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
    ev[0].record()
    for i in range(args.reps):
        mult(x, y)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.reps))
    bytes_ = ((4.0 if f32 else 8.0) * bs * bs + 4.0) * M.nnzb + 16.0 * bs * M.nrows + 4.0 * M.nrows
    med = ts[len(ts) // 2]
    print(f"{os.environ.get('CASHOCS_B200_LIB', 'default')}: {name}: median {med * 1e3:.1f} us  {bytes_ / med / 1e6:.0f} GB/s  (min {ts[0] * 1e3:.1f} us)", flush=True)
