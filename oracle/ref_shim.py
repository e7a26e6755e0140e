"""Import shim for the UNMODIFIED reference (`/root/reference/picard`).

TEST INFRASTRUCTURE ONLY.  Nothing in the product path imports this file.  It is used in
the build container (where `/root/reference` is mounted) by `oracle/make_golden.py` to
generate the committed fixtures under `tests/golden/`, by the `not gpu` test that
re-checks the CPU restatement (`oracle/dpi_oracle.py`) against the live reference when the
mount is present, and by `oracle/ref_arm.py` (the CPU legs of `bench.py`).  `/root/reference`
does not exist on the GPU box: there the shim resolves to `oracle/_ref/`, the byte-identical
staged copy made by `oracle/build_ref.py` (git-ignored, shipped by gpurun).

Why a shim: `import picard` pulls lightning / yacs / h5py / tensorboardX / matplotlib
(`picard/__init__.py:1-6`, `picard/data.py:8,12-13,24`, `picard/solution.py:6,18-19`,
`picard/data_saver.py:4`), none of which are installed here.  The hot-path modules
(`picard.equations`, `picard.data`, `picard.solution`, `picard.utils`) only need the
*names*; we register inert stand-ins and an empty `picard` package object whose
`__path__` points at the reference so `picard/__init__.py` is skipped.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")  # oracle/build_ref.py (travels to the GPU box)


def _default_root() -> str:
    """The mount in the build container; on the GPU box (no mount) the byte-identical staged copy."""
    if os.path.isdir("/root/reference/picard"):
        return "/root/reference"
    return _STAGED


REFERENCE_ROOT = os.environ.get("DPI_REFERENCE_ROOT") or _default_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "picard"))


class _AttrDict(dict):
    """Minimal stand-in for yacs.config.CfgNode (attribute access on a dict)."""

    def __init__(self, init=None, new_allowed=False, **kw):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = _AttrDict(v) if isinstance(v, dict) and not isinstance(v, _AttrDict) else v
        for k, v in kw.items():
            self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install() -> None:
    """Register stub modules and the bare `picard` package. Idempotent."""
    if "picard" in sys.modules and getattr(sys.modules["picard"], "_dpi_shim", False):
        return
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    import torch

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

        def print(self, *a, **k):
            pass

    class LightningDataModule:
        def __init__(self, *a, **k):
            pass

    class Callback:
        pass

    class Trainer:
        def __init__(self, *a, **k):
            pass

    if "lightning" not in sys.modules:
        pl_pytorch = _stub(
            "lightning.pytorch",
            LightningModule=LightningModule,
            LightningDataModule=LightningDataModule,
            Callback=Callback,
            Trainer=Trainer,
        )
        _stub("lightning.pytorch.utilities")
        _stub("lightning.pytorch.utilities.types", EVAL_DATALOADERS=object)
        _stub("lightning.pytorch.callbacks", Callback=Callback, ProgressBar=Callback, RichProgressBar=Callback)
        _stub("lightning.pytorch.loggers", WandbLogger=object, TensorBoardLogger=object)
        _stub("lightning.fabric")
        _stub("lightning.fabric.utilities")
        _stub("lightning.fabric.utilities.warnings", PossibleUserWarning=UserWarning)
        _stub(
            "lightning",
            LightningModule=LightningModule,
            LightningDataModule=LightningDataModule,
            Callback=Callback,
            Trainer=Trainer,
            pytorch=pl_pytorch,
        )
    if "yacs" not in sys.modules:
        _stub("yacs")
        _stub("yacs.config", CfgNode=_AttrDict)
    if "h5py" not in sys.modules:
        _stub("h5py", File=object)
    if "tensorboardX" not in sys.modules:
        _stub("tensorboardX", SummaryWriter=object)
    pkg = types.ModuleType("picard")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "picard")]
    pkg._dpi_shim = True
    sys.modules["picard"] = pkg


def cfgnode(d=None, **kw):
    return _AttrDict(d or {}, **kw)


@contextlib.contextmanager
def chdir(path: str):
    """The reference loads its `.pt` problem instances CWD-relative (equations.py:410-411,530-532)."""
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)
