#!/usr/bin/env python
"""bench.py — labelled collocation points/sec of the DPI Monte Carlo label generator on B200.

One "step" = one pass of the hot path (terminal estimator + integral estimator with u_k value/tangent evaluation
+ MC reduction) over one batch of synthetic collocation points.  Workload at N=1: BASELINE.json configs[1]
(`scripts/burgers/base_100d_T1.0_w1.0_0.yaml`): Cha nx=100, MLP [101,128,128,128,128,1] ELU, M_T=M_I=4096,
4096 points per step per GPU (one Picard iteration's DATA_SIZE).  Points shard across ranks with no data-path
collective (weak scaling: every rank labels `--points` points per step; Philox counters are keyed by the global
point index).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--precision f16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value` is device-resident throughput; `e2e` goes through the host-buffer C-ABI
call (`dpi_label_gen_host`: pinned host tx -> H2D -> kernels -> D2H y -> sync); `roofline` is the integral
(MLP-over-paths) kernel timed by CUDA events on its own stream inside the same run; `cpu_baseline` is the CPU
restatement of the reference (oracle/dpi_oracle.py, mode="autograd" = what the reference executes, fp64) timed on
this box's host cores on a bounded sample.  `--impl reference` times only that CPU arm.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "labelled collocation points/sec"
UNIT = "points/s"

WORKLOADS = {
    # name: (nx, hidden, M, default points per GPU per step, BASELINE config it restates)
    "burgers100d": dict(nx=100, hidden=[128] * 4, M=4096, points=4096, cha_k=5.0,
                        config="scripts/burgers/base_100d_T1.0_w1.0_0.yaml"),
    "burgers1000d": dict(nx=1000, hidden=[128] * 4, M=256, points=16384, cha_k=5.0,
                         config="synthetic Burgers 1000d DPI scaling sweep"),
}


def flops_per_point(nx: int, hidden, M_I: int) -> float:
    """F_point = (M_I + 1) * F_row; F_row = 2 * (forward MACs + tangent MACs); the first-layer
    tangent along a fixed direction is a constant vector (0 flops)."""
    dims = [1 + nx] + list(hidden) + [1]
    fwd = sum(dims[i] * dims[i + 1] for i in range(len(dims) - 1))
    jvp = sum(dims[i] * dims[i + 1] for i in range(1, len(dims) - 1))
    return float((M_I + 1) * 2 * (fwd + jvp))


def make_weights(nx: int, hidden, seed: int = 0):
    """torch.nn.Linear default init under a fixed seed (a trained net is not needed to measure throughput)."""
    import torch

    torch.manual_seed(seed)
    dims = [1 + nx] + list(hidden) + [1]
    lins = [torch.nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)]
    return [(l.weight.detach().float().clone(), l.bias.detach().float().clone()) for l in lins]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), burst=float(d["bf16_tflops"]),
                    sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, burst=1590.0, sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1])); pw.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the restatement of the reference's torch path (oracle), all host threads, fp64
# ------------------------------------------------------------------------------------------------
def cpu_arm(wl, weights, n_sample: int, steps: int, warmup: int):
    import torch
    from oracle import dpi_oracle as orc

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    eq = orc.Cha(wl["nx"], 1.0, wl["cha_k"], 1.0)
    net = orc.MLP([W.double() for W, _ in weights], [b.double() for _, b in weights])
    gen = torch.Generator().manual_seed(0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        orc.sample_with_gradients(eq, net, n_sample, wl["M"], wl["M"], None, mode="autograd", generator=gen)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tot = sum(times)
    return dict(value=n_sample * len(times) / tot, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                sample=f"{n_sample} points/step x {len(times)} steps (+{warmup} warm-up), fp64, M_T=M_I={wl['M']}, "
                       f"oracle/dpi_oracle.py mode=autograd (forward + reverse-mode like picard/data.py:1226-1325)",
                ms_per_step=1e3 * tot / len(times))


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    weights = make_weights(wl["nx"], wl["hidden"])
    n_sample = args.cpu_points
    cb = cpu_arm(wl, weights, n_sample, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "restates": wl["config"], "nx": wl["nx"],
                   "mlp": [1 + wl["nx"]] + wl["hidden"] + [1], "M_T": wl["M"], "M_I": wl["M"],
                   "points_per_step": n_sample, "note": "CPU arm: bounded sample of the same workload"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
def run_b200(args, wl):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from deeppicarditeration_b200.engine import LabelEngine

    nx, M, n = wl["nx"], wl["M"], args.points or wl["points"]
    weights = make_weights(nx, wl["hidden"])
    eng = LabelEngine(f"cuda:{local}")
    eng.set_problem("Cha", nx, 1.0, 1.0, cha_k=wl["cha_k"])
    eng.set_network("mlp", main=weights)
    prec = args.precision
    base = rank * n  # contiguous shard of the global point index range
    tx = eng.sample_tx(n, seed=0, point_index_base=base)
    tx_host = tx.cpu().pin_memory()
    y_host = torch.empty_like(tx_host).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=tx.device)  # > 126 MB L2

    def step_dev(i):
        return eng.label_gen(tx, M, M, precision=prec, seed=1 + i, point_index_base=base)

    def step_host(i):
        eng.label_gen_host(n, M, M, tx_host=tx_host, y_host=y_host, precision=prec, seed=1 + i,
                           point_index_base=base)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K, W, stage=None):
        for i in range(W):
            fn(i)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        stage_ms = []
        barrier()
        l0 = eng.launch_count()
        for i in range(K):
            flush.zero_()  # L2 flush between timed iterations (outside the per-step event pair)
            ev[i][0].record()
            fn(W + i)
            ev[i][1].record()
            if stage is not None:
                stage_ms.append(eng.stage_ms(stage))
        barrier()
        launches = eng.launch_count() - l0
        tot = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([tot], device=tx.device, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, stage_ms

    eng.set_stage_timing(True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    tot_ms, launches, stage_ms = timed(step_dev, args.steps, args.warmup, stage="integral")
    clk = clocks.stop() if rank == 0 else None
    eng.set_stage_timing(False)
    e2e_ms, _, _ = timed(step_host, args.steps, max(3, args.warmup))

    y = step_dev(0)
    finite = bool(torch.isfinite(y).all())
    if dist is not None:
        dist.barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = load_peaks()
    ms_step = tot_ms / args.steps
    value = world * n / (ms_step * 1e-3)
    e2e_value = world * n / (e2e_ms / args.steps * 1e-3)
    F = flops_per_point(nx, wl["hidden"], M)
    k_ms = statistics.mean(stage_ms)
    achieved = n * F / (k_ms * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            rec = json.load(f).get(f"{args.workload}:{'f16' if prec in ('f16', 'bf16') else prec}")
            # ncu dram__bytes_read.sum + dram__bytes_write.sum of the integral kernel, valid for the same launch size only
            if isinstance(rec, dict) and rec.get("launch_points") == n:
                traffic = rec.get("dram_bytes_per_launch")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16" if prec in ("f16", "bf16") else "f32", "data": "synthetic",
        "config": {"workload": args.workload, "restates": wl["config"], "equation": "Cha", "nx": nx,
                   "mlp": [1 + nx] + wl["hidden"] + [1], "M_T": M, "M_I": M, "points_per_gpu_per_step": n,
                   "global_points_per_step": world * n, "parallelism": f"points sharded x{world}, no collective",
                   "rng": "on-device Philox4x32-10 keyed by global point index",
                   "l2": "256 MiB buffer written between timed iterations",
                   "weights": "nn.Linear default init, torch.manual_seed(0)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": tx_host.numel() * 4,
                "d2h_bytes_per_step": y_host.numel() * 4, "api": "dpi_label_gen_host (pinned host buffers)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["sustained"], "traffic": traffic,
                     "kernel": "integral estimator (u_k value+tangent over paths)", "kernel_ms": k_ms,
                     "kernel_share_of_step": k_ms / ms_step, "flops_per_point": F,
                     "peak_source": f"{peaks['source']} bf16 dense, sustained (kernel timed inside a long step)",
                     "peak_burst": peaks["burst"], "frac_burst": achieved / peaks["burst"]},
        "clocks": clk, "labels_finite": finite,
    }
    if world == 1 and not args.no_cpu:
        cb = cpu_arm(wl, weights, args.cpu_points, args.cpu_steps, 1)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="burgers100d", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=os.environ.get("DPI_BENCH_PRECISION", "f16"), choices=["fp32", "f16", "bf16"],
                    help="f16: tcgen05 path (fp16 operands, fp32 accumulation; labels within 1e-2 of the reference); "
                         "fp32: CUDA-core path (within 1e-4)")
    ap.add_argument("--points", type=int, default=0, help="points per GPU per step (default: workload's)")
    ap.add_argument("--cpu-points", type=int, default=32, help="points per CPU-arm step (bounded sample)")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    wl = WORKLOADS[args.workload]

    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        if args.impl == "reference":
            return run_reference(args, wl)
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
               os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if args.impl == "reference":
        return run_reference(args, wl)
    return run_b200(args, wl)


if __name__ == "__main__":
    sys.exit(main())
