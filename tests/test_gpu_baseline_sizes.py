"""GPU parity AT THE BASELINE SIZES against the pinned CPU oracle (not against a sibling kernel).

Every BASELINE.json config at its own Monte Carlo width -- Burgers-100d M_T = M_I = 4096 (32 tiles = 2 work units per
point on the tensor-core path), GBM/SDGD M = 1024 with v = 100, HJB/PISGradNet 4 x 512 at M = 512, Burgers-1000d at
M = 256 -- in deterministic mode: the oracle (`mode="closed"`, fp64) and the CUDA path consume the same fp32-rounded
draws.  Tolerances are the north-star ones (rel-L2 <= 1e-4 for the fp32 paths, <= 1e-2 for the 16-bit / TF32
tensor-core paths), asserted separately on the value column, the gradient columns and the integral estimator alone
(label minus the oracle's terminal part), so a defect that hides behind the dominant terminal term is still seen.

The second half closes the loop for the on-device Philox mode: `dpi_export_draws` materialises the draws a Philox-mode
call consumes, the deterministic mode replays them bit for bit, and the oracle labels on those exported draws must
match the Philox-mode labels to the same tolerances.
"""
import pytest
import torch

from oracle import dpi_oracle as orc
from tests.golden_util import Golden

pytestmark = pytest.mark.gpu

TOL32, TOL16 = 1e-4, 1e-2


def _check(y, y_ref, yT_ref, tol, tol_integral=None, what=""):
    y, y_ref = y.double().cpu(), y_ref.double()
    assert torch.isfinite(y).all(), what
    errs = {
        "all": orc.rel_l2(y, y_ref),
        "value": orc.rel_l2(y[:, :1], y_ref[:, :1]),
        "grad": orc.rel_l2(y[:, 1:], y_ref[:, 1:]),
    }
    yI_ref = y_ref - yT_ref
    if float(yI_ref.norm()) > 0:
        errs["integral_value"] = orc.rel_l2((y - yT_ref)[:, :1], yI_ref[:, :1])
        errs["integral_grad"] = orc.rel_l2((y - yT_ref)[:, 1:], yI_ref[:, 1:])
    print(f"{what}: " + "  ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    for k, v in errs.items():
        lim = (tol_integral if tol_integral is not None else tol) if k.startswith("integral") else tol
        assert v < lim, (what, k, v, lim)
    return errs


def _oracle(eq, net, n, M_T, M_I, v, seed):
    d = orc.draw(eq, n, M_T, M_I, v, torch.Generator().manual_seed(seed), round_fp32=True)
    tx = orc.sample_tx(eq, d).float().double()  # the kernels see fp32 points
    yT = orc.estimate_terminal_with_gradients(eq, tx, d.zT, M_T)
    yI = orc.estimate_integral_with_gradients(eq, net, tx, d.u_s, d.zS, d.idx, M_I, mode="closed")
    return d, tx, yT, (yT + yI).detach()


# ---------------------------------------------------------------------------------------------------------------
# deterministic mode, oracle draws
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("gain", [1.0, 2.0])
def test_burgers_100d_m4096_matches_oracle(gain):
    """BASELINE configs 1/2: Cha nx=100, MLP [101,128,128,128,128,1], M_T = M_I = 4096 (data.py:471-527,899-926)."""
    from tests.gpu_util import draws_dict, engine_for
    eq = orc.Cha(100, 1.0, 5.0, 1.0)
    net = orc.make_mlp(101, [128] * 4, seed=21, gain=gain)
    n, M = 3, 4096
    d, tx, yT, y_ref = _oracle(eq, net, n, M, M, None, seed=100 + int(gain))
    e = engine_for(eq, net)
    y32 = e.label_gen(tx, M, M, precision="fp32", draws=draws_dict(d), check=True)
    _check(y32, y_ref, yT, TOL32, what=f"burgers100d fp32 gain={gain}")
    y16 = e.label_gen(tx, M, M, precision="f16", draws=draws_dict(d), check=True)
    _check(y16, y_ref, yT, TOL16, what=f"burgers100d f16 gain={gain}")
    e.close()


def test_gbm_100d_m1024_v100_matches_oracle():
    """BASELINE config 4: GBMEquationComplexExact nx=100, MLP [101,64,64,64,1], M = 1024, SDGD v = 100
    (data.py:497-502,1273-1303; equations.py:457-466) with the shipped problem instance (w, v of the golden fixture)."""
    from tests.gpu_util import draws_dict, engine_for
    g = Golden("gbm_100d_mlp64_sdgd")
    net = orc.make_mlp(101, [64] * 3, seed=22, gain=1.5)
    n, M, v = 3, 1024, 100
    d, tx, yT, y_ref = _oracle(g.eq, net, n, M, M, v, seed=102)
    e = engine_for(g.eq, net, v)
    y32 = e.label_gen(tx, M, M, precision="fp32", draws=draws_dict(d), check=True)
    _check(y32, y_ref, yT, TOL32, what="gbm100d fp32")
    y16 = e.label_gen(tx, M, M, precision="f16", draws=draws_dict(d), check=True)
    _check(y16, y_ref, yT, TOL16, what="gbm100d f16")
    e.close()


def test_hjb_100d_4x512_m512_matches_oracle():
    """BASELINE config 3 at the shipped width: OUProcessEquation nx=100 (GMM K=5), PISGradNet hidden [512]*4
    (solution.py:256-289; equations.py:657-666; utils.py:852-915), M = 512: strict fp32 and the TF32 tensor-core GEMMs."""
    from tests.gpu_util import draws_dict, engine_for
    g = Golden("hjb_100d_pisgrad")
    net = orc.make_pisgrad(g.nx, [512] * 4, T=g.T, seed=23, phase_scale=0.3)
    n, M = 3, 512
    d, tx, yT, y_ref = _oracle(g.eq, net, n, M, M, None, seed=103)
    e = engine_for(g.eq, net)
    y32 = e.label_gen(tx, M, M, precision="fp32", draws=draws_dict(d), check=True)
    _check(y32, y_ref, yT, TOL32, tol_integral=5 * TOL32, what="hjb100d fp32")
    y19 = e.label_gen(tx, M, M, precision="tf32", draws=draws_dict(d), check=True)
    _check(y19, y_ref, yT, TOL16, what="hjb100d tf32")
    y16 = e.label_gen(tx, M, M, precision="f16", draws=draws_dict(d), check=True)  # fused tcgen05 kernel (dpi_hjb_tc.cu)
    _check(y16, y_ref, yT, TOL16, what="hjb100d f16 fused")
    assert not torch.equal(y16, y19)
    e.close()


def test_burgers_1000d_m256_matches_oracle():
    """BASELINE config 5 at its smallest sweep width: Cha nx=1000, MLP [1001,128,128,128,128,1], M = 256."""
    from tests.gpu_util import draws_dict, engine_for
    eq = orc.Cha(1000, 1.0, 5.0, 1.0)
    net = orc.make_mlp(1001, [128] * 4, seed=24, gain=1.2)
    n, M = 3, 256
    d, tx, yT, y_ref = _oracle(eq, net, n, M, M, None, seed=104)
    e = engine_for(eq, net)
    y32 = e.label_gen(tx, M, M, precision="fp32", draws=draws_dict(d), check=True)
    _check(y32, y_ref, yT, TOL32, what="burgers1000d fp32")
    y16 = e.label_gen(tx, M, M, precision="f16", draws=draws_dict(d), check=True)
    _check(y16, y_ref, yT, TOL16, what="burgers1000d f16")
    e.close()


# ---------------------------------------------------------------------------------------------------------------
# Philox mode: exported draws -> deterministic replay (bitwise) and the oracle (tolerance)
# ---------------------------------------------------------------------------------------------------------------
def _philox_roundtrip(eq, net, v, n, M_T, M_I, precisions, what, tol_integral32=None):
    from tests.gpu_util import engine_for
    e = engine_for(eq, net, v)
    seed, base = 77, 1000
    tx = e.sample_tx(n, seed=seed, point_index_base=base)
    dr = e.export_draws(n, M_T, M_I, seed=seed, point_index_base=base)
    cpu = lambda t: None if t is None else t.cpu()
    d = orc.Draws(None, None, None, cpu(dr["zT"]).double(), cpu(dr["u_s"]).double(), cpu(dr["zS"]).double(),
                  None if dr["idx"] is None else cpu(dr["idx"]).long())
    txd = tx.double().cpu()
    yT = orc.estimate_terminal_with_gradients(eq, txd, d.zT, M_T)
    y_ref = (yT + orc.estimate_integral_with_gradients(eq, net, txd, d.u_s, d.zS, d.idx, M_I, mode="closed")).detach()
    for prec, tol in precisions:
        y_ph = e.label_gen(tx, M_T, M_I, precision=prec, seed=seed, point_index_base=base, check=True)
        y_det = e.label_gen(tx, M_T, M_I, precision=prec, draws=dr, check=True)
        # the deterministic mode fed the exported draws IS the Philox mode: same kernels, same summation order
        assert torch.equal(y_ph, y_det), (what, prec, float((y_ph - y_det).abs().max()))
        _check(y_ph, y_ref, yT, tol, tol_integral=tol_integral32 if tol == TOL32 else None, what=f"{what} philox {prec}")
    e.close()


def test_philox_mode_burgers_100d_m4096_replays_and_matches_oracle():
    eq = orc.Cha(100, 1.0, 5.0, 1.0)
    _philox_roundtrip(eq, orc.make_mlp(101, [128] * 4, seed=31, gain=1.5), 0, 3, 4096, 4096,
                      [("fp32", TOL32), ("f16", TOL16)], "burgers100d")


def test_philox_mode_gbm_100d_m1024_replays_and_matches_oracle():
    g = Golden("gbm_100d_mlp64_sdgd")
    _philox_roundtrip(g.eq, orc.make_mlp(101, [64] * 3, seed=32, gain=1.5), 100, 3, 1024, 1024,
                      [("fp32", TOL32), ("f16", TOL16)], "gbm100d")


def test_philox_mode_hjb_100d_4x512_replays_and_matches_oracle():
    g = Golden("hjb_100d_pisgrad")
    _philox_roundtrip(g.eq, orc.make_pisgrad(g.nx, [512] * 4, T=g.T, seed=33, phase_scale=0.3), 0, 3, 512, 512,
                      [("fp32", TOL32), ("tf32", TOL16), ("f16", TOL16)], "hjb100d", tol_integral32=5 * TOL32)


def test_philox_mode_burgers_1000d_m256_replays_and_matches_oracle():
    eq = orc.Cha(1000, 1.0, 5.0, 1.0)
    _philox_roundtrip(eq, orc.make_mlp(1001, [128] * 4, seed=34, gain=1.2), 0, 3, 256, 256,
                      [("fp32", TOL32), ("f16", TOL16)], "burgers1000d")


@pytest.mark.parametrize("hidden,M_I,n", [([512] * 4, 300, 5), ([256] * 3, 129, 4), ([128] * 2, 64, 3)])
def test_hjb_fused_kernel_shapes_ragged_tiles_and_shards(hidden, M_I, n):
    """The fused PISGradNet kernel on other eligible shapes (hidden width a multiple of 128, 2-4 hidden layers), ragged
    128-row tiles, deterministic mode against the oracle, and bitwise shard invariance in Philox mode."""
    from tests.gpu_util import draws_dict, engine_for
    g = Golden("hjb_100d_pisgrad")
    net = orc.make_pisgrad(g.nx, hidden, T=g.T, seed=51, phase_scale=0.3)
    M_T = 32
    d, tx, yT, y_ref = _oracle(g.eq, net, n, M_T, M_I, None, seed=105)
    e = engine_for(g.eq, net)
    y16 = e.label_gen(tx, M_T, M_I, precision="f16", draws=draws_dict(d), check=True)
    _check(y16, y_ref, yT, TOL16, what=f"hjb fused {hidden} M_I={M_I}")
    y32 = e.label_gen(tx, M_T, M_I, precision="fp32", draws=draws_dict(d), check=True)
    assert not torch.equal(y16, y32)  # the fused kernel really ran
    txp = e.sample_tx(7, seed=2)
    yp = e.label_gen(txp, M_T, M_I, precision="f16", seed=9)
    yq = torch.cat([e.label_gen(txp[:3], M_T, M_I, precision="f16", seed=9),
                    e.label_gen(txp[3:], M_T, M_I, precision="f16", seed=9, point_index_base=3)])
    assert torch.equal(yp, yq) and torch.isfinite(yp).all()
    e.close()
