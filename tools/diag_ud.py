"""Developer helper: per-sample (u, D) of the tensor-core path (debug build) vs the fp64 oracle on replayed draws."""
import os, sys, subprocess, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from oracle import dpi_oracle as orc
    from tests.gpu_util import engine_for
    eq = orc.Cha(100, 1.0, 5.0, 1.0)
    net = orc.make_mlp(101, [128] * 4, seed=2, gain=2.0)
    n, M = 4, 256
    d = orc.draw(eq, n, 16, M, None, torch.Generator().manual_seed(0), round_fp32=True)
    tx = orc.sample_tx(eq, d)
    e = engine_for(eq, net)
    y16 = e.label_gen(tx, 16, M, precision="f16", draws={"zT": d.zT, "u_s": d.u_s, "zS": d.zS, "idx": None})
    torch.cuda.synchronize()
    ud = torch.from_numpy(np.fromfile(os.environ["DPI_TC_UD_FILE"], dtype=np.float32).reshape(n, M, 2)).double()
    t = tx[:, :1].repeat_interleave(M, 0); x = tx[:, 1:].repeat_interleave(M, 0)
    s = d.u_s * (eq.T - t) + t
    Xs = x + torch.sqrt(s - t) * eq.alpha_sqrt * d.zS
    u, du, _ = net.jets(torch.cat([s, Xs], -1), None, order=1)
    u = u.view(n, M); D = du.view(n, M)
    ru = (ud[..., 0] - u); rD = (ud[..., 1] - D)
    print("u : rel-L2 %.3e, mean ratio %.5f, max abs err %.3e (|u| mean %.3e)" % (float(ru.norm()/u.norm()), float((ud[...,0]/u).median()), float(ru.abs().max()), float(u.abs().mean())))
    print("D : rel-L2 %.3e, median ratio %.5f, max abs err %.3e (|D| mean %.3e)" % (float(rD.norm()/D.norm()), float((ud[...,1]/D).median()), float(rD.abs().max()), float(D.abs().mean())))
    idx = rD.abs().view(-1).argsort(descending=True)[:8]
    print("worst D samples: (D_ref, D_tc, u_ref, u_tc, a)")
    a = (torch.sqrt(s - t)).view(-1)
    for i in idx:
        print("   %.4f %.4f %.4f %.4f  a=%.4f" % (float(D.view(-1)[i]), float(ud[..., 1].reshape(-1)[i]), float(u.view(-1)[i]), float(ud[..., 0].reshape(-1)[i]), float(a[i])))
else:
    out = os.path.join(ROOT, "gpurun_out", "ud.bin"); os.makedirs(os.path.dirname(out), exist_ok=True)
    env = dict(os.environ, DPI_B200_LIB=os.path.join(ROOT, "deeppicarditeration_b200", "variants", "lib_ud.so"), DPI_TC_UD_FILE=out)
    print(subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout[-3000:])
