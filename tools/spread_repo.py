"""GPU side of the acceptance check (north star, second criterion): this repository's Deep Picard Iteration on the
config oracle/measure_spread.py runs the reference on, same seeds count, fp32 and f16 labels.  Writes JSON next to
the reference's; `python tools/spread_repo.py table` prints the side-by-side table from the two files."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TAG = os.environ.get("SPREAD_TAG", "")  # e.g. "_ds4096": a second config next to the default one
REF = os.path.join(ROOT, "profiles", f"r02_spread_reference{TAG}.json")
OUT = os.path.join(ROOT, "profiles", f"r02_spread_repo{TAG}.json")


def table():
    ref, rep = json.load(open(REF)), json.load(open(OUT))
    rows = ["| run | final rRMSE (min / mean / max over seeds) | final rRMSEg (min / mean / max) | best-iteration rRMSE (mean) |", "|---|---|---|---|"]

    def stats(runs):
        import numpy as np
        f = np.array([[r[-1]["rRMSE"], r[-1]["rRMSEg"]] for r in runs.values()])
        best = np.mean([min(m["rRMSE"] for m in r) for r in runs.values()])
        return (f"{f[:,0].min():.3e} / {f[:,0].mean():.3e} / {f[:,0].max():.3e}",
                f"{f[:,1].min():.3e} / {f[:,1].mean():.3e} / {f[:,1].max():.3e}", f"{best:.3e}")
    rows.append("| reference (CPU fp64, %d seeds) | %s | %s | %s |" % ((len(ref["runs"]),) + stats(ref["runs"])))
    for prec, runs in rep["runs"].items():
        rows.append("| this repo, %s labels (B200, %d seeds) | %s | %s | %s |" % ((prec, len(runs)) + stats(runs)))
    print("\n".join(rows))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "table":
        table()
        sys.exit(0)
    import torch
    from deeppicarditeration_b200 import equations as eqs
    from deeppicarditeration_b200.picard import run_picard
    cfg = dict(nx=100, k=5.0, alpha=1.0, T=1.0, neurons=[128] * 4, M=1024, data_size=1024, batch_size=128, n_epochs=16,
               fixed_weight=1.0, loss_beta=0.0, eval_points=10000)
    if os.path.exists(REF):
        cfg = json.load(open(REF))["config"]
    n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    seeds = [int(s) for s in sys.argv[2:]] or [0, 1, 2, 3, 4]
    out = dict(config=cfg, iterations=n_iter, impl="deeppicarditeration_b200.run_picard (label kernels + regression step, fp32 weights)", runs={})
    for prec in ("fp32", "f16"):
        out["runs"][prec] = {}
        for s in seeds:
            eq = eqs.Cha(nx=cfg["nx"], alpha=cfg["alpha"], k=cfg["k"], T=cfg["T"])
            r = run_picard(eq, cfg["neurons"], n_iter, data_size=cfg["data_size"], n_mc=cfg["M"], batch_size=cfg["batch_size"],
                           n_epochs=cfg["n_epochs"], fixed_weight=cfg["fixed_weight"], loss_beta=cfg["loss_beta"],
                           precision=prec, eval_points=cfg["eval_points"], eval_freq=1, seed=s, log=None)
            out["runs"][prec][str(s)] = [dict(iteration=m["iteration"], loss=m["loss"], rRMSE=m["rRMSE"], rRMSEg=m["rRMSEg"]) for m in r["history"]]
            print(prec, "seed", s, "rRMSE per iteration:", " ".join(f"{m['rRMSE']:.3f}" for m in r["history"]), flush=True)
    json.dump(out, open(OUT, "w"), indent=1)
