"""Developer helper: time the GBM/SDGD and HJB/PISGradNet label paths at the BASELINE config shapes."""
import sys, time, torch
sys.path.insert(0, '.')
from oracle import dpi_oracle as orc
from tests.gpu_util import engine_for
from tests.golden_util import Golden

which = sys.argv[1]
n = int(sys.argv[2])
prec = sys.argv[3] if len(sys.argv) > 3 else 'fp32'
if which == "gbm":
    g = Golden("gbm_100d_mlp64_sdgd")
    eq, M, v = g.eq, 1024, 100
    net = orc.make_mlp(101, [64, 64, 64], seed=0)
    F_row = 3319040
else:
    g = Golden("hjb_100d_pisgrad")
    eq, M, v = g.eq, 4096, 0
    net = orc.make_pisgrad(100, [512] * 4, T=1.0, seed=0, phase_scale=0.3)
    F_row = 3694720
e = engine_for(eq, net, v)
tx = e.sample_tx(n, seed=0)
e.set_stage_timing(True)
for _ in range(2):
    y = e.label_gen(tx, M, M, seed=1, precision=prec)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
K = 3
for _ in range(K):
    y = e.label_gen(tx, M, M, seed=1, precision=prec)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / K
print(f"{which} [{prec}] n={n} M={M}: {ms:.2f} ms/call  {n / ms * 1e3:.1f} pts/s  terminal={e.stage_ms('terminal'):.2f} ms integral={e.stage_ms('integral'):.2f} ms  "
      f"integral {n * (M + 1) * F_row / e.stage_ms('integral') / 1e9:.1f} TFLOP/s  finite={bool(torch.isfinite(y).all())}")
