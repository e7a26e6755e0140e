"""Developer helper (not part of the product or the bench contract): quick timing + error print on a GPU box."""
import sys, time, torch
sys.path.insert(0, '.')
from oracle import dpi_oracle as orc
from tests.gpu_util import engine_for, draws_dict
from tests.golden_util import Golden
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
M = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
nx = int(sys.argv[4]) if len(sys.argv) > 4 else 100
if prec == "err":
    g = Golden("burgers_100d_mlp128")
    e = engine_for(g.eq, g.net, g.v)
    for pr in ("fp32", "f16"):
        y = e.label_gen(g.tx, g.M_T, g.M_I, precision=pr, draws=draws_dict(g.draws)).double().cpu()
        print(pr, "rel-L2 all %.3e value %.3e grad %.3e integral-only %.3e" % (
            orc.rel_l2(y, g.y), orc.rel_l2(y[:, :1], g.y[:, :1]), orc.rel_l2(y[:, 1:], g.y[:, 1:]),
            orc.rel_l2(y - g.y_terminal, g.y_integral)))
    sys.exit(0)
eq = orc.Cha(nx, 1.0, 5.0, 1.0)
net = orc.make_mlp(1 + nx, [128] * 4, seed=0)
e = engine_for(eq, net)
tx = e.sample_tx(n, seed=0)
e.set_stage_timing(True)
for _ in range(2):
    y = e.label_gen(tx, M, M, precision=prec, seed=1)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
K = 3
st = []
for _ in range(K):
    y = e.label_gen(tx, M, M, precision=prec, seed=1)
    st.append([e.stage_ms(s) for s in ("terminal", "baseline", "integral")])
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / K
F = (222976 if nx == 100 else (2 * (177408 + 49280))) * (M + 1)
print(f"{prec} nx={nx} n={n} M={M}: {ms:.2f} ms/call  {n / ms * 1e3:.1f} pts/s  stages(ms) term/base/integral={st[-1]}  "
      f"integral {n * F / st[-1][2] / 1e9:.1f} TFLOP/s  finite={bool(torch.isfinite(y).all())}")
