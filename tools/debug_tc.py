"""Debug driver: one bf16 label_gen on a golden fixture (deterministic draws) and one in Philox mode."""
import sys
import torch
from oracle import dpi_oracle as orc
from tests.golden_util import Golden
from tests.gpu_util import engine_for, draws_dict

g = Golden("burgers_100d_mlp128")
e = engine_for(g.eq, g.net, g.v)
mode = sys.argv[1] if len(sys.argv) > 1 else "det"
if mode == "det":
    y = e.label_gen(g.tx, g.M_T, g.M_I, precision="f16", draws=draws_dict(g.draws)).double().cpu()
    print("n", g.n, "M_I", g.M_I, "rel", orc.rel_l2(y, g.y), "grad rel", orc.rel_l2(y[:, 1:], g.y[:, 1:]))
else:
    n, M = int(sys.argv[2]), int(sys.argv[3])
    tx = e.sample_tx(n, seed=3)
    y32 = e.label_gen(tx, 64, M, precision="fp32", seed=5).double().cpu()
    y16 = e.label_gen(tx, 64, M, precision="f16", seed=5).double().cpu()
    torch.cuda.synchronize()
    print("n", n, "M", M, "rel", orc.rel_l2(y16, y32), "value rel", orc.rel_l2(y16[:, :1], y32[:, :1]))
