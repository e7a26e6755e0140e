import sys, torch
sys.path.insert(0,'.')
from tests.golden_util import Golden
from tests.gpu_util import engine_for, draws_dict
from oracle import dpi_oracle as orc
for name in ("burgers_100d_mlp128","burgers_100d_mlp128_td"):
    g=Golden(name); e=engine_for(g.eq,g.net,g.v)
    y=e.label_gen(g.tx,g.M_T,g.M_I,precision="f16",draws=draws_dict(g.draws),estimate_delta_t=g.delta_t).double().cpu()
    print(name,'label',orc.rel_l2(y,g.y),'grad',orc.rel_l2(y[:,1:],g.y[:,1:]),'integral part',orc.rel_l2(y-g.y_terminal,g.y_integral))
    e.close()
