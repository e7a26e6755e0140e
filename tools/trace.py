"""Developer helper: run the DPI_TC_TRACE variant once and print the CTA-0 timeline summary."""
import os, sys, subprocess, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "gpurun_out", "tc_trace.txt")
os.makedirs(os.path.dirname(out), exist_ok=True)
env = dict(os.environ, DPI_B200_LIB=os.path.join(ROOT, "deeppicarditeration_b200", "variants", "lib_trace.so"),
           DPI_TC_TRACE_FILE=out)
code = """
import sys; sys.path.insert(0, %r)
import torch
from oracle import dpi_oracle as orc
from tests.gpu_util import engine_for
eq = orc.Cha(100, 1.0, 5.0, 1.0); net = orc.make_mlp(101, [128]*4, seed=0)
e = engine_for(eq, net); tx = e.sample_tx(592, seed=0)
for _ in range(2): y = e.label_gen(tx, 64, 4096, precision='bf16', seed=1)
torch.cuda.synchronize(); print('ok')
""" % ROOT
print(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout)
ev = [tuple(map(int, l.split())) for l in open(out)]
ev.sort(key=lambda x: x[1])
t0 = ev[0][1]
# skip the first 20% (warm-up), summarise the epilogue slot-0 timeline
E = [(c, t - t0) for c, t in ev if c // 1000000 == 2]
M = [(c, t - t0) for c, t in ev if c // 1000000 == 1]
P = [(c, t - t0) for c, t in ev if c // 1000000 == 3]
print("events", len(ev), "span clk", ev[-1][1] - t0)
def show(evs, n0, n1):
    for c, t in evs[n0:n1]:
        role, slot, j, tag = c // 1000000, (c // 100000) % 10, (c // 100) % 1000, c % 100
        print(f"  t={t:9d} role={role} slot={slot} j={j:3d} tag={tag:2d}")
mid = len(E) // 2
print("epilogue events (mid-run):"); show(E, mid, mid + 40)
# statistics: epilogue wait durations and pass durations per tag
d = collections.defaultdict(list)
last = {}
for c, t in E:
    slot, j, tag = (c // 100000) % 10, (c // 100) % 1000, c % 100
    key = (slot,)
    if key in last:
        ptag, pt = last[key]
        d[(ptag, tag)].append(t - pt)
    last[key] = (tag, t)
print("epilogue intervals (prev tag -> tag): mean clk, n")
for k in sorted(d): print("  ", k, sum(d[k]) / len(d[k]), len(d[k]))
dm = collections.defaultdict(list); lastm = None
for c, t in M:
    slot, tag = (c // 100000) % 10, c % 100
    if lastm: dm[(lastm[0], tag)].append(t - lastm[1])
    lastm = (tag, t)
print("MMA thread intervals:")
for k in sorted(dm): print("  ", k, sum(dm[k]) / len(dm[k]), len(dm[k]))
dp = collections.defaultdict(list); lastp = None
for c, t in P:
    tag = c % 100
    if lastp: dp[(lastp[0], tag)].append(t - lastp[1])
    lastp = (tag, t)
print("producer intervals:")
for k in sorted(dp): print("  ", k, sum(dp[k]) / len(dp[k]), len(dp[k]))
mm = len(M) // 2
print("MMA events (mid-run):"); show(M, mm, mm + 30)

try:
    rows = [tuple(map(int, l.split())) for l in open(out + ".cta")]
    t00 = min(r[1] for r in rows); dur = sorted((r[2] - r[1], r[0], r[3], r[1] - t00) for r in rows)
    print("per-CTA wall ns: min %d median %d max %d; kernel span %d" % (dur[0][0], dur[len(dur)//2][0], dur[-1][0], max(r[2] for r in rows) - t00))
    print("slowest:", dur[-8:]); print("fastest:", dur[:8])
except OSError:
    pass
