import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cashocs_b200 as cb
from cashocs_b200.geometry import quality

def tm(f, reps=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): r = f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, r

n = int(sys.argv[1]) if len(sys.argv) > 1 else 119
m = cb.regular_mesh(n, 1., 1., 1.)
m.build_pattern()
V = (torch.rand(m.num_vertices * 3, dtype=torch.float64, device="cuda") - 0.5) * (0.1 / n)
ap = cb.APrioriMeshTester(m); it = cb.IntersectionTester(m); dh = cb.DeformationHandler(m, ap, it)
print("a priori %.2f ms" % tm(lambda: ap.test(V, float("inf")))[0])
print("quality  %.2f ms" % tm(lambda: quality.compute_mesh_quality(m))[0])
print("intersection %.2f ms" % tm(lambda: it.test())[0])
def mv():
    ok = dh.move_mesh(V, validated_a_priori=True, test_for_intersections=False)
    dh.revert_transformation()
print("move+revert %.2f ms" % tm(mv)[0])

# same phases inside the real pipeline (gradient direction)
import bench
bench.KSP["pc_type"] = "jacobi"
sop = bench.build_problem(n)
g = sop.compute_shape_gradient()[0].vector
step = -0.1 / n / float(g.abs().max())
mh = sop.mesh_handler
print("pipeline: a priori %.2f ms" % tm(lambda: mh.a_priori_tester.test(g, float("inf"), step))[0])
def mv2():
    ok = mh.deformation_handler.move_mesh(g, validated_a_priori=True, test_for_intersections=False, step=step)
    t0 = time.perf_counter(); r = mh.a_posteriori_tester.test(); torch.cuda.synchronize(); t1 = time.perf_counter()
    q = quality.compute_mesh_quality(sop.mesh); torch.cuda.synchronize(); t2 = time.perf_counter()
    mh.deformation_handler.revert_transformation()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, r, q
print("pipeline: move (intersection ms, quality ms, ok, q):", tm(mv2)[1])
print("pipeline: full move_mesh %.2f ms" % tm(lambda: (mh.move_mesh(g, step=step), mh.revert_transformation()))[0])
