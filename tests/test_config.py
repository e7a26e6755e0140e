"""CPU tests of the config loader (picard/config.py semantics without yacs)."""
import os

import pytest
import torch

from deeppicarditeration_b200 import config as C


@pytest.fixture(autouse=True)
def _restore_dtype():
    yield
    torch.set_default_dtype(torch.float32)


def _write(tmp_path, name, text):
    (tmp_path / name).write_text(text)


def test_base_chain_name_join_overrides_and_none(tmp_path, monkeypatch):
    _write(tmp_path, "a.yaml", "NAME: A\nPICARD: {N: 5}\nDATA: {FLOAT: double, PREFETCH_FACTOR: None, kwargs: {n_estimate_terminal: 8}}\n")
    _write(tmp_path, "b.yaml", "BASE: a.yaml\nNAME: B\nTRAIN: {N_EPOCHS: 3}\nNETWORK: {BOUND: None}\n")
    _write(tmp_path, "c.yaml", "BASE: b.yaml\nNAME: C\nPICARD: {N: 7}\nEQUATION: {cls: Cha, kwargs: {nx: 10, k: 5.0}}\n")
    monkeypatch.chdir(tmp_path)  # BASE paths are CWD-relative (config.py:229-254)
    cfg = C.load_cfg("c.yaml", ["--TRAIN.BATCH_SIZE", "64", "DATA.kwargs.n_estimate_integral", "16", "NETWORK.NEURONS", "[8, 8]",
                                "DATA.DEVICE", "cpu"])
    assert cfg.NAME == "A_B_C" and cfg.PICARD.N == 7 and cfg.TRAIN.N_EPOCHS == 3 and cfg.TRAIN.BATCH_SIZE == 64
    assert cfg.DATA.PREFETCH_FACTOR is None and cfg.NETWORK.BOUND is None and cfg.BASE is None
    assert dict(cfg.DATA.kwargs) == {"n_estimate_terminal": 8, "n_estimate_integral": 16}
    assert cfg.NETWORK.NEURONS == [8, 8] and cfg.EQUATION.kwargs.nx == 10 and cfg.DATA.DEVICE == "cpu"
    assert torch.get_default_dtype() == torch.float64  # apply_cfg (config.py:194-200)
    assert cfg.TRAIN.LOSS.beta == 0.0 and cfg.EVAL.L2_N_POINTS == 10000  # defaults survive
    cfg.freeze()
    with pytest.raises(AttributeError):
        cfg.PICARD.N = 1
    assert "PICARD" in cfg.dump()


def test_unknown_keys_and_base_override_are_rejected(tmp_path, monkeypatch):
    _write(tmp_path, "bad.yaml", "PICARD: {M: 3}\n")
    _write(tmp_path, "ok.yaml", "NAME: x\n")
    monkeypatch.chdir(tmp_path)
    with pytest.raises(KeyError):
        C.load_cfg("bad.yaml")
    with pytest.raises(KeyError):
        C.load_cfg("ok.yaml", ["PICARD.M", "3"])
    with pytest.raises(ValueError):
        C.load_cfg("ok.yaml", ["BASE", "other.yaml"])
    with pytest.raises(ValueError):
        C.load_cfg("ok.yaml", ["PICARD.N"])
    cfg = C.load_cfg("ok.yaml", ["EQUATION.kwargs.anything", "1.5"])  # open node accepts new keys
    assert cfg.EQUATION.kwargs.anything == 1.5


@pytest.mark.skipif(not os.path.isdir("/root/reference/scripts"), reason="reference not mounted (GPU box)")
def test_every_reference_yaml_loads(monkeypatch):
    import glob
    files = glob.glob("/root/reference/scripts/*/*.yaml") + glob.glob("/root/reference/scripts/*/*/*.yaml")
    assert len(files) >= 13
    for f in files:
        monkeypatch.chdir(os.path.dirname(f))
        cfg = C.load_cfg(os.path.basename(f), ["DATA.DEVICE", "cpu"])
        assert cfg.EQUATION.cls in ("Cha", "OUProcessEquation", "GBMEquationComplexExact")
    monkeypatch.chdir("/root/reference/scripts/fully_nonlinear/case_1")
    cfg = C.load_cfg("base_100d_T1.0_w0.0_nov_0.yaml", ["DATA.DEVICE", "cpu"])
    assert cfg.DATA.HESSIAN_APPROXIMATION.method == "SDGD" and cfg.DATA.HESSIAN_APPROXIMATION.kwargs.v == 100
    assert cfg.DATA.kwargs.n_estimate_integral == 1024 and cfg.DATA.DATA_SIZE == 2048 and cfg.EVAL.TEST_HESSIAN is True
