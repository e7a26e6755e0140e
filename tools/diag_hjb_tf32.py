"""Developer helper: HJB/PISGradNet labels with the nn_module GEMMs on the tensor cores (TF32) vs the fp32 pipeline on
the same Philox draws, and against the golden fixture on replayed draws."""
import sys, torch
sys.path.insert(0, '.')
from oracle import dpi_oracle as orc
from tests.gpu_util import engine_for, draws_dict
from tests.golden_util import Golden

g = Golden("hjb_100d_pisgrad")
e = engine_for(g.eq, g.net, g.v)
for prec in ("fp32", "tf32"):
    y = e.label_gen(g.tx, g.M_T, g.M_I, precision=prec, draws=draws_dict(g.draws)).double().cpu()
    print(f"golden [{prec}] n={g.n} M_I={g.M_I}: all {orc.rel_l2(y, g.y):.3e} value {orc.rel_l2(y[:, :1], g.y[:, :1]):.3e} "
          f"grad {orc.rel_l2(y[:, 1:], g.y[:, 1:]):.3e}")
e.close()
net = orc.make_pisgrad(100, [512] * 4, T=1.0, seed=0, phase_scale=0.3)
e = engine_for(g.eq, net, 0)
tx = e.sample_tx(48, seed=2)
y32 = e.label_gen(tx, 256, 700, precision="fp32", seed=5).double().cpu()
y19 = e.label_gen(tx, 256, 700, precision="tf32", seed=5).double().cpu()
print(f"512-wide, n=48 M_I=700: tf32 vs fp32 all {orc.rel_l2(y19, y32):.3e} value {orc.rel_l2(y19[:, :1], y32[:, :1]):.3e} "
      f"grad {orc.rel_l2(y19[:, 1:], y32[:, 1:]):.3e} finite={bool(torch.isfinite(y19).all())}")
