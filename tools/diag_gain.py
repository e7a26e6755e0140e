import sys, torch
sys.path.insert(0, '.')
from oracle import dpi_oracle as orc
from tests.gpu_util import engine_for
eq = orc.Cha(100, 1.0, 5.0, 1.0)
for gain in (1.0, 1.3, 2.0, 3.0, 5.0):
    net = orc.make_mlp(101, [128] * 4, seed=2, gain=gain)
    e = engine_for(eq, net)
    tx = e.sample_tx(256, seed=3)
    yz = engine_for(eq, orc.ZeroNet()).label_gen(tx, 1024, 1024, seed=5).double()   # terminal part only
    y32 = e.label_gen(tx, 1024, 1024, precision="fp32", seed=5).double() - yz
    y16 = e.label_gen(tx, 1024, 1024, precision="f16", seed=5).double() - yz
    rel = lambda a, b: float((a - b).norm() / b.norm())
    print(f"gain {gain}: integral part  value rel {rel(y16[:, :1], y32[:, :1]):.3e}  grad rel {rel(y16[:, 1:], y32[:, 1:]):.3e}  "
          f"|I_value| {float(y32[:, 0].abs().mean()):.3e}  max|y16| {float(y16.abs().max()):.3e}")
