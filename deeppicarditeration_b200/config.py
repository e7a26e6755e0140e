"""Configuration loader with the reference's key tree and merge semantics (picard/config.py),
written without yacs (not installed here): the four BASELINE YAMLs load unchanged.

Semantics kept:
  * the defaults tree below (picard/config.py:9-116); nodes listed in OPEN_NODES accept arbitrary keys, unknown keys
    anywhere else are errors;
  * `BASE: file.yaml` (relative to the CWD) is followed recursively; merge order deepest base -> ... -> top file -> CLI
    pairs; `NAME` = "_".join(all base NAMEs + the top NAME); the BASE key itself cannot be overridden from the CLI
    (picard/config.py:174-177,229-257);
  * CLI pairs `KEY VALUE` with leading dashes stripped from the key (picard/main.py:20), values parsed as Python
    literals when possible; the YAML scalar `None` (a string to PyYAML) is decoded to Python None, as yacs does;
  * `apply_cfg`: DATA.FLOAT selects torch's default dtype; DATA.DEVICE becomes "cuda" unless it is "cpu" or CUDA is
    unavailable (picard/config.py:194-200).
"""
from __future__ import annotations

import ast
import copy
import os
from typing import Any, Dict, Iterable, List, Optional

import yaml

DEFAULTS: Dict[str, Any] = {
    "BASE": None, "FORCE": False, "NAME": "exp",
    "EQUATION": {"cls": "AllenCahnEquation", "kwargs": {}},
    "METHOD": {"cls": "Picard", "num_v_samples": 16, "K": 20, "dt": 0.005, "num_sub_iter": 100},
    "PICARD": {"N": 1, "FORMULA": None},
    "TRAIN": {
        "BATCH_SIZE": 2048, "N_EPOCHS": 1, "SUPERVISE_GRADIENT": None, "SUPERVISE_HESSIAN": None, "NUM_HESS_SAMPLES": -1,
        "LOSS": {"beta": 0.0, "SCALER": {"cls": None, "kwargs": {}}, "FN": {"cls": None, "kwargs": {}},
                 "use_aux_loss": False, "weight_aux_loss": 0.1},
        "OPTIMIZER": {"cls": "Adam", "kwargs": {}, "SCHEDULER": {"cls": None, "kwargs": {}, "config": {}}},
    },
    "NETWORK": {"cls": None, "TYPE": "Value", "NEURONS": [10, 10], "ACTIVATIONS": ["Tanh", "Tanh"], "BOUND": None,
                "RELOAD": False, "USE_T_EMBEDDING": False, "PISGRADNET": False, "PRETRAIN_PATH": None, "kwargs": {}},
    "DATA": {
        "kwargs": {}, "SAVE": False, "ONLINE": True, "TRAIN_FILE": "", "N_WORKERS": 1, "DATA_SIZE": 2048 * 5000,
        "NEW_SAMPLING": False, "N_BUFFER": None, "RESERVED_MEMORY": None,
        "MEMORY": {"RESERVED": None, "REDUCE_FACTOR": 1.0, "REUSE": 9999999},
        "PREFETCH_FACTOR": None, "DEVICE": None, "FLOAT": "float", "EXACT": False, "SHUFFLE": None, "PRELOAD": False,
        "PRELOAD_N_WORKERS": None, "HESSIAN_APPROXIMATION": {"method": None, "kwargs": {}}, "SAMPLE_BOUND": None,
        "ESTIMATE_TERMINAL": "OU_ByGx", "ESTIMATE_INTEGRAL": "OU_Simple", "ESTIMATE_DELTA_T": 0.0,
    },
    "LOGGING": {"LOGGER": "wandb", "kwargs": {"project": "picard", "offline": False}, "TENSORBOARD_DIR": "tensorboard"},
    "EVAL": {"L2_N_POINTS": 10_000, "FREQ": None, "BATCH_SIZE": None, "TEST_GRAD": False, "TEST_HESSIAN": False},
}
# nodes created with new_allowed=True in the reference
OPEN_NODES = {"EQUATION.kwargs", "TRAIN.LOSS.SCALER.kwargs", "TRAIN.LOSS.FN.kwargs", "TRAIN.OPTIMIZER.kwargs",
              "TRAIN.OPTIMIZER.SCHEDULER.kwargs", "TRAIN.OPTIMIZER.SCHEDULER.config", "NETWORK.kwargs", "DATA.kwargs",
              "LOGGING.kwargs", "DATA.HESSIAN_APPROXIMATION.kwargs"}


class Config(dict):
    """Attribute-style view of the nested config (read like the reference reads its CfgNode: cfg.DATA.kwargs.x)."""

    __frozen = False

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        if k.startswith("_Config__"):
            return dict.__setattr__(self, k, v)
        if self.__frozen:
            raise AttributeError(f"config is frozen; cannot set {k}")
        self[k] = v

    def freeze(self):
        dict.__setattr__(self, "_Config__frozen", True)
        for v in self.values():
            if isinstance(v, Config):
                v.freeze()
        return self

    def to_dict(self) -> dict:
        return {k: (v.to_dict() if isinstance(v, Config) else copy.deepcopy(v)) for k, v in self.items()}

    def dump(self) -> str:
        return yaml.safe_dump(self.to_dict(), sort_keys=False)


def _wrap(d: dict) -> Config:
    return Config({k: (_wrap(v) if isinstance(v, dict) else copy.deepcopy(v)) for k, v in d.items()})


def _decode(v):
    """YAML hands `None` over as the string "None"; yacs literal-evals strings where that works."""
    if isinstance(v, str):
        if v == "None":
            return None
        return v
    if isinstance(v, list):
        return [_decode(x) for x in v]
    return v


def _merge(dst: Config, src: dict, path: str = "", open_node: bool = False) -> None:
    """Merge `src` into `dst`; keys must exist unless the node (or an ancestor) is one of OPEN_NODES."""
    open_node = open_node or path in OPEN_NODES
    for k, v in src.items():
        here = f"{path}.{k}" if path else k
        if k not in dst and not open_node:
            raise KeyError(f"Non-existent config key: {here}")
        if isinstance(v, dict):
            if not isinstance(dst.get(k), Config):
                if k in dst and not open_node and dst[k] is not None:
                    raise KeyError(f"config key {here} is a scalar, got a section")
                dst[k] = Config()
            _merge(dst[k], v, here, open_node)
        else:
            if isinstance(dst.get(k), Config) and not open_node:
                raise KeyError(f"config key {here} is a section, got a scalar")
            dst[k] = _decode(v)


def _read_chain(path: str) -> List[dict]:
    """[deepest base, ..., top file] (picard/config.py:229-254); BASE paths are relative to the CWD."""
    chain, seen = [], set()
    while path is not None:
        real = os.path.abspath(path)
        if real in seen:
            raise ValueError(f"BASE cycle through {path}")
        seen.add(real)
        with open(path) as f:
            doc = yaml.safe_load(f) or {}
        chain.append(doc)
        path = _decode(doc.get("BASE"))
    return chain[::-1]


def _parse_cli(pairs: Iterable[str]) -> List[tuple]:
    pairs = list(pairs)
    if len(pairs) % 2:
        raise ValueError(f"override list must be KEY VALUE pairs, got {pairs}")
    out = []
    for k, v in zip(pairs[0::2], pairs[1::2]):
        k = k.lstrip("-")
        if k == "BASE":
            raise ValueError("BASE cannot be overridden from the command line")
        try:
            val = ast.literal_eval(v)
        except (ValueError, SyntaxError):
            val = v
        out.append((k, _decode(val)))
    return out


def get_default_cfg() -> Config:
    return _wrap(DEFAULTS)


def load_cfg(config_file: Optional[str] = None, opts: Iterable[str] = ()) -> Config:
    """picard/config.py:242-266."""
    cfg = get_default_cfg()
    names = []
    if config_file is not None:
        for doc in _read_chain(config_file):
            doc = dict(doc)
            doc.pop("BASE", None)
            if "NAME" in doc:
                names.append(str(doc["NAME"]))
            _merge(cfg, doc)
        if names:
            cfg["NAME"] = "_".join(names)
    cfg["BASE"] = None
    for key, val in _parse_cli(opts):
        node, parts, path = cfg, key.split("."), ""
        for p in parts[:-1]:
            if p not in node or not isinstance(node[p], Config):
                raise KeyError(f"Non-existent config key: {key}")
            node = node[p]
            path = f"{path}.{p}" if path else p
        if parts[-1] not in node and path not in OPEN_NODES:
            raise KeyError(f"Non-existent config key: {key}")
        node[parts[-1]] = val
    if cfg.DATA.RESERVED_MEMORY is not None:  # compatibility_check, picard/config.py:119-125
        if cfg.DATA.MEMORY.RESERVED is not None:
            raise ValueError("Both RESERVED_MEMORY and MEMORY.RESERVED are set.")
        cfg.DATA.MEMORY.RESERVED = cfg.DATA.RESERVED_MEMORY
    apply_cfg(cfg)
    return cfg


FLOAT_NAMES = {"float": "float32", "float32": "float32", "f32": "float32", "single": "float32", "32": "float32",
               "double": "float64", "float64": "float64", "f64": "float64", "64": "float64"}


def get_standard_float_dtype(name):
    """picard/config.py:132-144."""
    import torch

    key = str(name).lower()
    if key not in FLOAT_NAMES:
        raise ValueError(f"Unknown DATA.FLOAT {name}")
    return getattr(torch, FLOAT_NAMES[key])


def apply_cfg(cfg: Config) -> None:
    """picard/config.py:194-200: default dtype from DATA.FLOAT; DATA.DEVICE -> "cuda" when available and not "cpu"."""
    import torch

    torch.set_default_dtype(get_standard_float_dtype(cfg.DATA.FLOAT))
    if cfg.DATA.DEVICE != "cpu" and torch.cuda.is_available():
        cfg.DATA.DEVICE = "cuda"
    elif cfg.DATA.DEVICE is None:
        cfg.DATA.DEVICE = "cpu"
