"""Recipe for `oracle/_ref/`: the UNMODIFIED reference package, staged so that it can travel to the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference is pure Python (nothing to compile), so the "build" is a copy of
`/root/reference/picard/*.py` -- byte for byte, checked by SHA-256 in `MANIFEST.json` -- into `oracle/_ref/picard/`.
`oracle/_ref/` is git-ignored (the reference's sources never enter this repository's history) but not
gpurun-ignored, so `bench.py --impl reference` and `bench.py`'s `cpu_baseline` leg can time the reference's OWN
`OnlineDataGenerator.sample_with_gradients` (picard/data.py:211-223) on the GPU box's host cores, behind the import
shim `oracle/ref_shim.py` (stub modules for lightning / yacs / h5py / tensorboardX, which are not installed).
Nothing in the product path reads this directory; when it is absent the CPU legs fall back to the port
(`oracle/dpi_oracle.py`) and say so (`cpu_baseline.kind = "port"`).

    python oracle/build_ref.py            # in the build container, where /root/reference is mounted
`__graft_entry__.build()` runs it when the mount is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("DPI_REFERENCE_SRC", "/root/reference")
REF_OUT = os.path.join(HERE, "_ref")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build_ref(verbose: bool = False) -> bool:
    """Stage the reference package; returns False (and leaves `_ref` alone) when the mount is absent."""
    src = os.path.join(REF_SRC, "picard")
    if not os.path.isdir(src):
        return False
    dst = os.path.join(REF_OUT, "picard")
    os.makedirs(dst, exist_ok=True)
    manifest = {}
    for name in sorted(os.listdir(src)):
        if not name.endswith(".py"):
            continue
        s, d = os.path.join(src, name), os.path.join(dst, name)
        if not os.path.exists(d) or _sha(d) != _sha(s):
            shutil.copyfile(s, d)
        manifest[f"picard/{name}"] = _sha(d)
    for stale in set(os.listdir(dst)) - {k.split("/", 1)[1] for k in manifest} - {"__pycache__"}:
        os.remove(os.path.join(dst, stale))
    with open(os.path.join(REF_OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} reference files staged from {src}")
    return True


def verify_ref() -> bool:
    """True when `_ref` holds exactly the files its manifest lists (used by the CPU test)."""
    mf = os.path.join(REF_OUT, "MANIFEST.json")
    if not os.path.exists(mf):
        return False
    with open(mf) as f:
        files = json.load(f)["files"]
    return all(os.path.exists(os.path.join(REF_OUT, k)) and _sha(os.path.join(REF_OUT, k)) == v for k, v in files.items())


if __name__ == "__main__":
    ok = build_ref(verbose=True)
    sys.exit(0 if ok else 1)
