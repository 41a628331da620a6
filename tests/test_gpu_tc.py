"""The stated reduced-precision mode (chest_set_precision(SPLIT_BF16), kernels_tc.cuh): the estimated-CSI interference
cancellation on tcgen05 tensor cores with split-BF16 operands and FP32 accumulation in TMEM, against the FP64 DMMA path
of the same library on the same seeded draws (which is itself bit-identical in its decisions to the CPU oracle).
Tolerance (north_star): 1e-4 relative on the estimated quantities; hard decisions may differ only where an equalised
symbol sits within that distance of a decision boundary."""
import numpy as np
import pytest

from tests.helpers import context_from_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _states(ctx, B, snrs, reps):
    out = {}
    for name in ("aux", "cod", "ofdm"):
        for r in reps:
            for s in snrs:
                out[(name, r, s, "xD")] = ctx.get_state("xD_est", name, r, s)
                out[(name, r, s, "hdiag")] = ctx.get_state("hdiag", name, r, s)
                out[(name, r, s, "hP")] = ctx.get_state("hP", name, r, s)
    return out


@pytest.mark.parametrize("B", [40, 300])
def test_split_bf16_mode_matches_fp64_mode(ds_default, B):
    """B = 40: two full 16-column units and a ragged one per (scheme, SNR) -- one 128-column work item with 88 unused
    columns; B = 300: three work items per (scheme, SNR), 18 row tiles each, several items per persistent CTA.
    One iteration isolates the arithmetic (no decision feedback): the estimated channel (pilot estimates, h = diag(D_est))
    agrees to 1e-4 -- measured ~3e-6 / 1e-5 -- and the data-symbol estimates x = y_ic / h element-wise to 1e-4 of their own
    magnitude except where a faded h amplifies the difference (max-norm bound 1e-2).  Four iterations: decision feedback makes
    the trajectories differ wherever one symbol flips, so the check is the fraction of hard decisions that differ."""
    S = ds_default
    ctx = context_from_oracle(S, max_batch=B)
    seed, first = 5, 1000
    reps = sorted({0, 15, 16, B // 2, B - 1})
    snrs = (0, ctx.n_snr - 1)
    ref1 = ctx.run_batch(B, 1, None, seed=seed, first_rep=first)
    st_ref = _states(ctx, B, snrs, reps)
    ref4 = ctx.run_batch(B, 4, None, seed=seed, first_rep=first)
    ctx.set_precision("split_bf16")
    got1 = ctx.run_batch(B, 1, None, seed=seed, first_rep=first)
    st_got = _states(ctx, B, snrs, reps)
    got4 = ctx.run_batch(B, 4, None, seed=seed, first_rep=first)
    mode, flops, img_bytes = ctx.precision_info()
    assert mode == "split_bf16" and flops > 0 and img_bytes > 0
    worst = {"hP": 0.0, "hdiag": 0.0, "xD": 0.0}
    for key, a in st_ref.items():
        b = st_got[key]
        rel = np.max(np.abs(a - b)) / np.max(np.abs(a))
        worst[key[3]] = max(worst[key[3]], rel)
        if key[3] == "xD":
            own = np.abs(a - b) / np.maximum(np.abs(a), 1e-3 * np.max(np.abs(a)))
            assert np.quantile(own, 0.9) < TOL and rel < 1e-2, (key, rel, np.quantile(own, 0.9))
        else:
            assert rel < TOL, (key, rel)
    n_bits = ctx.bit_counts()
    for ref, got, n_it in ((ref1, got1, 1), (ref4, got4, 4)):
        # the perfect-CSI twin and the one-tap stage never touch the tensor-core kernel
        assert np.array_equal(got[:, :, :, :, 1, :], ref[:, :, :, :, 1, :])
        assert np.array_equal(got[:, :, 0], ref[:, :, 0])
        diff = np.abs(got.astype(np.int64) - ref.astype(np.int64))
        frac = diff[:, :, 1:, :, 0, 0].sum() / (B * ctx.n_snr * n_it * n_bits[:, 0].sum())
        print("split-BF16 vs FP64, %d iteration(s): worst deviations %s, counter cells that differ %d of %d, bit-decision fraction %.2e"
              % (n_it, {k: "%.1e" % v for k, v in worst.items()}, int((diff > 0).sum()), diff.size, frac))
        assert frac < (2e-5 if n_it == 1 else 5e-4)
    # back to FP64: bit-identical to the first run
    ctx.set_precision("fp64")
    again = ctx.run_batch(B, 4, None, seed=seed, first_rep=first)
    assert np.array_equal(again, ref4)
    ctx.close()


def test_split_bf16_mode_needs_the_factored_pass(ds_default):
    import chest_b200
    ctx = context_from_oracle(ds_default, max_batch=16)
    ctx.set_perfect_csi_mode("dense")
    ctx.set_precision("split_bf16")
    with pytest.raises(chest_b200.ChestError, match="factored"):
        ctx.run_batch(16, 1, None, seed=1, first_rep=0)
    ctx.close()
