"""`picard train <yaml> [KEY VALUE ...]` (picard/main.py:12-23) on the B200 stages.

    python -m deeppicarditeration_b200.main train base_100d_T1.0_w1.0_0.yaml PICARD.N 10 DATA.kwargs.precision f16
    torchrun --nproc-per-node 8 -m deeppicarditeration_b200.main train <yaml> ...      # points sharded, DP regression

Run it from the directory of the YAML, like the reference: `BASE:` files, the experiment directory (`NAME`) and the
equations' `.pt` problem instances are all relative to the CWD.  `picard evaluate` is broken
upstream and is not mirrored.
"""
from __future__ import annotations

import argparse
import os
import sys


def train(configfile: str, opts, precision=None, seed: int = 0):
    import torch

    from .config import load_cfg
    from .runner import PicardRunner

    if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1 and not torch.distributed.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl")
    cfg = load_cfg(configfile, [o.lstrip("-") if i % 2 == 0 else o for i, o in enumerate(opts)])  # main.py:20
    cfg.freeze()
    runner = PicardRunner(cfg, precision=precision, seed=seed)
    try:
        return runner.run()
    finally:
        runner.close()
        if torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="picard", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    tr = sub.add_parser("train", help="solve the PDE of a config by Deep Picard Iteration")
    tr.add_argument("configfile")
    tr.add_argument("--precision", default=None, choices=["fp32", "f16"],
                    help="label arithmetic (default: DATA.kwargs.precision or f16)")
    tr.add_argument("--seed", type=int, default=0)
    tr.add_argument("opts", nargs=argparse.REMAINDER, help="KEY VALUE overrides")
    a = ap.parse_args(argv)
    train(a.configfile, a.opts, precision=a.precision, seed=a.seed)
    return 0


if __name__ == "__main__":
    sys.exit(main())
