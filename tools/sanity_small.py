"""Developer helper: one small call of every tensor-core path (for compute-sanitizer runs)."""
import sys, torch
sys.path.insert(0, '.')
from oracle import dpi_oracle as orc
from tests.gpu_util import engine_for
from tests.golden_util import Golden

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "burgers"):
    e = engine_for(orc.Cha(100, 1.0, 5.0, 1.0), orc.make_mlp(101, [128] * 4, seed=0))
    y = e.label_gen(e.sample_tx(3, seed=1), 64, 300, precision="f16", seed=2)
    print("burgers f16", bool(torch.isfinite(y).all())); e.close()
if which in ("all", "chunked"):
    e = engine_for(orc.Cha(300, 1.0, 5.0, 1.0), orc.make_mlp(301, [128] * 4, seed=0))
    y = e.label_gen(e.sample_tx(3, seed=1), 32, 400, precision="f16", seed=2)
    print("burgers d=300 f16", bool(torch.isfinite(y).all())); e.close()
if which in ("all", "hjb"):
    g = Golden("hjb_100d_pisgrad")
    e = engine_for(g.eq, orc.make_pisgrad(100, [512] * 4, T=1.0, seed=0, phase_scale=0.3), 0)
    y = e.label_gen(e.sample_tx(3, seed=1), 32, 70, precision="tf32", seed=2)
    print("hjb tf32", bool(torch.isfinite(y).all())); e.close()
if which in ("all", "gbm"):
    g = Golden("gbm_100d_mlp64_sdgd")
    e = engine_for(g.eq, orc.make_mlp(101, [64, 64, 64], seed=0), 100)
    y = e.label_gen(e.sample_tx(3, seed=1), 32, 40, precision="f16", seed=2)
    print("gbm f16", bool(torch.isfinite(y).all())); e.close()
torch.cuda.synchronize()
