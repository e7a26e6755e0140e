import sys, time, torch
sys.path.insert(0, '.')
from oracle import dpi_oracle as orc
from tests.gpu_util import engine_for
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
M = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
nx = int(sys.argv[4]) if len(sys.argv) > 4 else 100
eq = orc.Cha(nx, 1.0, 5.0, 1.0)
net = orc.make_mlp(1 + nx, [128] * 4, seed=0)
e = engine_for(eq, net)
tx = e.sample_tx(n, seed=0)
for _ in range(2):
    y = e.label_gen(tx, M, M, precision=prec, seed=1)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
K = 3
for _ in range(K):
    y = e.label_gen(tx, M, M, precision=prec, seed=1)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / K
F = (222976 if nx == 100 else (2 * (177408 + 49280))) * (M + 1)
print(f"{prec} nx={nx} n={n} M={M}: {ms:.2f} ms/call  {n / ms * 1e3:.1f} pts/s  {n * F / ms / 1e9:.1f} TFLOP/s  finite={bool(torch.isfinite(y).all())}")
