"""Developer helper: bias / noise of the 16-bit tensor-core label path against the fp32 path on the SAME Philox draws,
for a trained Burgers-100d network (20 Picard iterations with fp32 labels)."""
import sys, torch
sys.path.insert(0, '.')
from deeppicarditeration_b200 import equations as eqs
from deeppicarditeration_b200.picard import run_picard, _HA
from deeppicarditeration_b200.data import OnlineDataGenerator
eq = eqs.Cha(nx=100, alpha=1.0, k=5.0, T=1.0)
out = run_picard(eq, [128] * 4, n_iterations=20, precision="fp32", eval_freq=20, log=print)
model = out["model"]
gen = OnlineDataGenerator(eq, model, 100, 21, device="cuda", t_always_uniform=True, n_estimate_terminal=4096,
                          n_estimate_integral=4096, hessian_approximation=_HA(), precision="fp32", seed=777)
e = gen.engine
tx = e.sample_tx(4096, seed=777)
y32 = e.label_gen(tx, 4096, 4096, precision="fp32", seed=5).double()
y16 = e.label_gen(tx, 4096, 4096, precision="f16", seed=5).double()
y32b = e.label_gen(tx, 4096, 4096, precision="fp32", seed=6).double()   # independent MC noise level
t, x = tx[:, :1].double(), tx[:, 1:].double()
u_ex, ux_ex = eq.to(tx.device).u_u_x(t, x)
yex = torch.cat([u_ex, ux_ex], 1)
def rel(a, b): return float((a - b).norm() / b.norm())
print("value col : f16 vs fp32 %.3e | fp32 vs fp32(other seed) %.3e | fp32 vs exact %.3e | f16 vs exact %.3e" % (
    rel(y16[:, :1], y32[:, :1]), rel(y32b[:, :1], y32[:, :1]), rel(y32[:, :1], yex[:, :1]), rel(y16[:, :1], yex[:, :1])))
print("grad cols : f16 vs fp32 %.3e | fp32 vs fp32(other seed) %.3e | fp32 vs exact %.3e | f16 vs exact %.3e" % (
    rel(y16[:, 1:], y32[:, 1:]), rel(y32b[:, 1:], y32[:, 1:]), rel(y32[:, 1:], yex[:, 1:]), rel(y16[:, 1:], yex[:, 1:])))
d = y16 - y32
print("signed mean diff value %.3e (mean |y| %.3e); grad cols mean %.3e (mean |y| %.3e)" % (
    float(d[:, 0].mean()), float(y32[:, 0].abs().mean()), float(d[:, 1:].mean()), float(y32[:, 1:].abs().mean())))
# per-point mean over dims of the gradient difference vs exact gradient (all dims share the same exact value)
gd = d[:, 1:].mean(1)
print("per-point mean grad diff: mean %.3e  rms %.3e ; exact grad per dim mean %.3e" % (
    float(gd.mean()), float(gd.pow(2).mean().sqrt()), float(ux_ex.mean())))
# dependence on t
for lo, hi in ((0.0, 0.2), (0.2, 0.5), (0.5, 0.8), (0.8, 1.0)):
    m = ((t[:, 0] >= lo) & (t[:, 0] < hi))
    print(f"t in [{lo},{hi}): value diff mean %.3e rms %.3e | grad diff per-point-mean mean %.3e rms %.3e" % (
        float(d[m, 0].mean()), float(d[m, 0].pow(2).mean().sqrt()), float(gd[m].mean()), float(gd[m].pow(2).mean().sqrt())))
