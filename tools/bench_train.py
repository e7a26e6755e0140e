"""Developer helper: time the regression step at the BASELINE shape (MLP [101,128x4,1], batch 512, Adam) -- this
library vs a torch restatement of what the reference executes (vmap(jacrev) + backward + torch.optim.Adam) on the
same GPU."""
import sys, time, torch
sys.path.insert(0, '.')
from deeppicarditeration_b200.trainer import RegressionTrainer
from torch.func import jacrev, vmap, functional_call

nx, B, steps = 100, 512, 128
torch.manual_seed(0)
def mk(dtype):
    return torch.nn.Sequential(torch.nn.Linear(1 + nx, 128), torch.nn.ELU(), torch.nn.Linear(128, 128), torch.nn.ELU(),
                               torch.nn.Linear(128, 128), torch.nn.ELU(), torch.nn.Linear(128, 128), torch.nn.ELU(),
                               torch.nn.Linear(128, 1)).to(dtype).cuda()
tx = torch.cat([torch.rand(4096, 1), torch.randn(4096, nx)], 1).cuda()
y = torch.randn(4096, 1 + nx).cuda() * 0.3

model = mk(torch.float32)
tr = RegressionTrainer(model, fixed_weight=1.0, max_batch=B)
def ours():
    tr.fit(tx, y, batch_size=B, n_epochs=steps * B // 4096)
ours(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ours(); e1.record(); torch.cuda.synchronize()
ms_ours = e0.elapsed_time(e1)
print(f"b200 regression step: {ms_ours / steps * 1e3:.1f} us/step, {ms_ours:.2f} ms per {steps} steps (one Picard iteration's 16 epochs over 4096 points)")

for dtype in (torch.float32, torch.float64):
    ref = mk(dtype)
    opt = torch.optim.Adam(ref.parameters())
    txd, yd = tx.to(dtype), y.to(dtype)
    def fwd_aux(x):
        o = ref(x)
        return o, o
    jac = vmap(jacrev(fwd_aux, has_aux=True))
    def ref_steps():
        for s in range(steps):
            i = (s * B) % 4096
            xb, yb = txd[i:i + B], yd[i:i + B]
            opt.zero_grad()
            utx, u = jac(xb)
            ux = utx.view(*xb.shape)[:, 1:]
            loss = torch.mean((u.view(-1, 1) - yb[:, :1]) ** 2) + torch.mean((ux - yb[:, 1:]) ** 2, dim=0).sum()
            loss.backward()
            opt.step()
    ref_steps(); torch.cuda.synchronize()
    e0.record(); ref_steps(); e1.record(); torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1)
    print(f"torch ({dtype}) vmap(jacrev)+backward+Adam on the same GPU: {ms_ref / steps * 1e3:.1f} us/step, {ms_ref:.2f} ms per {steps} steps  -> x{ms_ref / ms_ours:.1f}")
