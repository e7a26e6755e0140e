"""Developer helper: the SDGD tensor-core kernel on a small steady-state problem (for ncu captures)."""
import sys, torch
sys.path.insert(0, '.')
import bench
from deeppicarditeration_b200.engine import LabelEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
wl = bench.WORKLOADS["gbm100d"]
eng = LabelEngine("cuda:0")
eng.set_problem("GBM", 100, 1.0, 1.0, **bench.problem_instance(wl))
eng.set_network(**bench.make_network(wl))
tx = eng.sample_tx(n, seed=0)
eng.set_stage_timing(True)
for i in range(3):
    y = eng.label_gen(tx, M, M, precision="f16", seed=1 + i)
    ms = eng.stage_ms("integral")
torch.cuda.synchronize()
F = bench.flops_per_point(wl, M)
print(f"gbm f16 n={n} M={M}: integral stage {ms:.2f} ms  {n / ms * 1e3:.0f} pts/s  {n * F / ms / 1e9:.1f} TFLOP/s  finite={bool(torch.isfinite(y).all())}")
