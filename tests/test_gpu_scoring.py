"""GPU parity: segmented stable radix sort, cut-point prefix sums, metric prologue / NLL / AUCE histogram
and the ``ause`` / ``auce`` drop-ins (C ABI) vs the oracle (oracle/metrics.py, pinned to the reference's
own ause.py / auce.py by tests/test_oracle_golden.py) and vs the committed golden vectors.

Bit-exact: sort permutations, sorted keys, AUCE interval counts.  Float: 1e-5 relative.
"""
import os

import numpy as np
import pytest
import torch

from oracle import metrics as om
from uncertainty_nerf_gs_b200 import synthetic

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-5


def _tricky_keys(n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, generator=g)
    x = torch.clamp(x, min=0.2)                      # a big tie group
    x[::7] = torch.round(x[::7] * 50) / 50           # many small tie groups
    x[1::101] = 0.0
    x[2::101] = -0.0
    x[3::211] = float("nan")
    x[4::211] = float("inf")
    x[5::211] = -float("inf")
    x[6::97] = -x[6::97]
    x[7::303] = 1e-42                                # subnormals
    return x


@pytest.mark.parametrize("n", [1, 2, 31, 4096, 4097, 10007, 640000])
def test_sort_matches_torch_stable(built_library, n):
    from uncertainty_nerf_gs_b200 import ops

    x = _tricky_keys(n, seed=n)
    ref_vals, ref_idx = torch.sort(x, stable=True)
    vals, perm = ops.segmented_sort(x.cuda(), [n])
    assert torch.equal(perm.cpu().long(), ref_idx)
    assert torch.equal(torch.nan_to_num(vals.cpu(), nan=123.0), torch.nan_to_num(ref_vals, nan=123.0))
    keys_only, none = ops.segmented_sort(x.cuda(), [n], want_perm=False)
    assert none is None
    assert torch.equal(torch.nan_to_num(keys_only.cpu(), nan=123.0), torch.nan_to_num(ref_vals, nan=123.0))


def test_sort_ragged_segments(built_library):
    from uncertainty_nerf_gs_b200 import ops

    lens = [5000, 0, 1, 4096, 12345, 7]
    x = _tricky_keys(sum(lens), seed=1)
    vals, perm = ops.segmented_sort(x.cuda(), lens)
    lo = 0
    for ln in lens:
        rv, ri = torch.sort(x[lo:lo + ln], stable=True)
        assert torch.equal(perm[lo:lo + ln].cpu().long(), ri)
        assert torch.equal(torch.nan_to_num(vals[lo:lo + ln].cpu(), nan=1.0), torch.nan_to_num(rv, nan=1.0))
        lo += ln


def test_sort_full_view_properties(built_library):
    """1297x840 keys x 3 segments: sortedness, permutation validity, stability, idempotence."""
    from uncertainty_nerf_gs_b200 import ops

    n = 1297 * 840
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.clamp(0.1 * torch.rand(3 * n, generator=g, device="cuda"), min=0.03) ** 2
    vals, perm = ops.segmented_sort(x, [n, n, n])
    for s in range(3):
        v, p, src = vals[s * n:(s + 1) * n], perm[s * n:(s + 1) * n].long(), x[s * n:(s + 1) * n]
        assert bool((v[1:] >= v[:-1]).all())
        assert torch.equal(src[p], v)
        assert torch.equal(torch.sort(p).values, torch.arange(n, device="cuda"))
        ties = v[1:] == v[:-1]
        assert bool((p[1:][ties] > p[:-1][ties]).all()), "ties must keep ascending original index"
    again, perm2 = ops.segmented_sort(vals, [n, n, n])
    assert torch.equal(again, vals)
    assert torch.equal(perm2.long(), torch.arange(n, device="cuda").repeat(3))


def test_cut_prefix_sums(built_library):
    from uncertainty_nerf_gs_b200 import ops
    from uncertainty_nerf_gs_b200.metrics import ause_cut_counts

    lens = [10007, 4096, 100]
    g = torch.Generator().manual_seed(2)
    a, b = torch.rand(sum(lens), generator=g), torch.rand(sum(lens), generator=g)
    perm = torch.cat([torch.randperm(ln, generator=g) for ln in lens]).int()
    cuts = np.stack([ause_cut_counts(ln) for ln in lens])
    out = ops.cut_prefix_sums([a.cuda(), b.cuda()], perm.cuda(), lens, cuts).cpu()
    plain = ops.cut_prefix_sums([a.cuda()], None, lens, cuts).cpu()
    lo = 0
    for s, ln in enumerate(lens):
        pa = a[lo:lo + ln][perm[lo:lo + ln].long()].double().cumsum(0)
        pb = b[lo:lo + ln][perm[lo:lo + ln].long()].double().cumsum(0)
        pp = a[lo:lo + ln].double().cumsum(0)
        for c, cut in enumerate(cuts[s]):
            ra = pa[cut - 1] if cut > 0 else torch.tensor(0.0, dtype=torch.float64)
            rb = pb[cut - 1] if cut > 0 else torch.tensor(0.0, dtype=torch.float64)
            rp = pp[cut - 1] if cut > 0 else torch.tensor(0.0, dtype=torch.float64)
            torch.testing.assert_close(out[s, 0, c], ra, rtol=1e-12, atol=1e-12)
            torch.testing.assert_close(out[s, 1, c], rb, rtol=1e-12, atol=1e-12)
            torch.testing.assert_close(plain[s, 0, c], rp, rtol=1e-12, atol=1e-12)
        lo += ln


@pytest.mark.parametrize("err_type", ["mae", "mse", "rmse"])
@pytest.mark.parametrize("n", [10007, 99, 50])
def test_ause_dropin_vs_oracle(built_library, err_type, n):
    from uncertainty_nerf_gs_b200.metrics import ause

    g = torch.Generator().manual_seed(n)
    unc = torch.clamp(0.1 * torch.rand(n, generator=g), min=0.03) ** 2   # tie group at the floor
    err = torch.rand(n, generator=g) ** 2
    r0, e0, v0, a0 = om.ause(unc, err, err_type)
    r1, e1, v1, a1 = ause(unc.cuda(), err.cuda(), err_type)
    assert np.array_equal(r0, r1)
    np.testing.assert_allclose(e1, e0, rtol=RTOL, atol=0, equal_nan=True)
    np.testing.assert_allclose(v1, v0, rtol=RTOL, atol=0, equal_nan=True)
    np.testing.assert_allclose(a1, a0, rtol=RTOL, atol=1e-9, equal_nan=True)
    assert type(a1) is type(a0) and v1.dtype == v0.dtype


def test_ause_golden_from_reference(built_library):
    """Golden vectors produced by the reference's own ause() (stable sort forced), tests/golden/make_golden.py."""
    from uncertainty_nerf_gs_b200.metrics import ause

    z = np.load(os.path.join(GOLDEN, "ause_golden.npz"))
    unc, err = torch.from_numpy(z["unc"]).cuda(), torch.from_numpy(z["err"]).cuda()
    for err_type in ("mae", "mse", "rmse"):
        _, e, v, a = ause(unc, err, err_type)
        np.testing.assert_allclose(e, z[f"{err_type}_err"], rtol=RTOL)
        np.testing.assert_allclose(v, z[f"{err_type}_err_by_var"], rtol=RTOL)
        np.testing.assert_allclose(a, z[f"{err_type}_ause"], rtol=RTOL, atol=1e-9)


def _auce_inputs(n, seed):
    g = torch.Generator().manual_seed(seed)
    m = torch.rand(n, 3, generator=g)
    s = torch.clamp(0.1 * torch.rand(n, 1, generator=g), min=0.03).repeat(1, 3)
    t = torch.clamp(m + s * torch.randn(n, 3, generator=g), 0, 1)
    t[::17] = m[::17]                       # exact hits (|t - m| = 0)
    s[5::29] = 0.0                          # sigma = 0: predicate reduces to t == m
    t[7::31] = float("nan")
    return m, s, t


def test_auce_counts_bit_exact(built_library):
    from uncertainty_nerf_gs_b200.metrics import auce

    m, s, t = _auce_inputs(20011, seed=4)
    ref = om.auce(m.numpy(), s.numpy(), t.numpy())
    out = auce(m.numpy(), s.numpy(), t.numpy())
    assert list(out.keys()) == list(ref.keys())
    n = float(m.numel())
    assert np.array_equal(np.rint(out["coverage_values"] * n), np.rint(ref["coverage_values"] * n))
    assert np.array_equal(out["coverage_values"], ref["coverage_values"])
    for k in ("avg_length_values", "coverage_error_values", "abs_coverage_error_values",
              "neg_coverage_error_values"):
        np.testing.assert_allclose(out[k], ref[k], rtol=RTOL, atol=1e-12)
    for k in ("auc_abs_error_values", "auc_length_values", "auc_neg_error_values"):
        np.testing.assert_allclose(out[k], ref[k], rtol=RTOL, atol=1e-12)


def test_auce_golden_from_reference(built_library):
    from uncertainty_nerf_gs_b200.metrics import auce

    z = np.load(os.path.join(GOLDEN, "auce_golden.npz"))
    out = auce(z["mean"], z["sigma"], z["target"])
    assert np.array_equal(out["coverage_values"], z["coverage_values"])
    np.testing.assert_allclose(out["avg_length_values"], z["avg_length_values"], rtol=RTOL)
    for k in ("auc_abs_error_values", "auc_length_values", "auc_neg_error_values"):
        np.testing.assert_allclose(out[k], z[k], rtol=RTOL, atol=1e-12)


@pytest.mark.parametrize("hw,batch", [((40, 50), 3), ((800, 800), 1)])
def test_score_rgb_batch_vs_oracle(built_library, hw, batch):
    """The fused scorer against get_unc_metrics_rgb restated by the oracle, image by image."""
    from uncertainty_nerf_gs_b200.metrics import score_rgb_batch

    imgs = [synthetic.scoring_image(*hw, seed=i) for i in range(batch)]
    pred = torch.stack([i[0] for i in imgs]).cuda()
    std = torch.stack([i[1] for i in imgs]).cuda()
    gt = torch.stack([i[2] for i in imgs]).cuda()
    outs = score_rgb_batch(pred, gt, std, min_rgb_std_for_nll=3e-2)
    for (p, s, g), out in zip(imgs, outs):
        ref = om.unc_metrics_rgb(p, g, s, min_rgb_std_for_nll=3e-2)
        assert np.array_equal(out["coverage_values"], ref["coverage_values"])
        for k in ("err_mae", "err_mse", "err_rmse", "err_var_mae", "err_var_mse", "err_var_rmse",
                  "avg_length_values"):
            np.testing.assert_allclose(out[k], ref[k], rtol=RTOL, atol=0)
        for k in ("ause_mae", "ause_mse", "ause_rmse", "nll_rgb", "avg_var", "auc_abs_error_values",
                  "auc_length_values", "auc_neg_error_values"):
            np.testing.assert_allclose(out[k], ref[k], rtol=RTOL, atol=1e-9)
        np.testing.assert_allclose(out["mse_mean"], float(ref["mse"].mean()), rtol=RTOL)


def test_prologue_vectors_match_torch(built_library):
    from uncertainty_nerf_gs_b200 import ops
    from uncertainty_nerf_gs_b200.metrics import _z_table

    p, s, g = synthetic.scoring_image(33, 47, seed=9)
    ref = om.rgb_metric_prologue(p, g, s)
    out = ops.score_prologue(p.reshape(-1, 3).cuda(), g.reshape(-1, 3).cuda(), s.reshape(-1).cuda(), [33 * 47],
                             _z_table("cuda"), nll_min_std=3e-2)
    torch.testing.assert_close(out["squared_error"].cpu(), ref["squared_error"], rtol=RTOL, atol=1e-9)
    torch.testing.assert_close(out["absolute_error"].cpu(), ref["absolute_error"], rtol=RTOL, atol=1e-9)
    assert torch.equal(out["var"].cpu(), ref["var"])
    nll = om.negative_gaussian_loglikelihood(p.reshape(-1, 3), g.reshape(-1, 3), s, eps=3e-2)
    torch.testing.assert_close(out["sums"][0, 3].cpu(), nll.double().sum(), rtol=1e-6, atol=0)


def test_auce_counts_exact_at_interval_boundaries(built_library):
    """Adversarial inputs for the three-tier coverage count (csrc/score_prologue.cu): targets placed on, and one
    float32 ulp either side of, the float64 interval ends m -+ z_k sigma for every threshold k; tiny and huge
    sigma relative to |m| (forces the general path); zero / NaN / negative sigma.  Counts must equal numpy's."""
    from scipy.stats import norm

    from uncertainty_nerf_gs_b200.metrics import auce

    rng = np.random.default_rng(7)
    alphas = np.arange(start=0.01, stop=1.0, step=0.01)
    z = norm.ppf(1 - alphas / 2)
    rows = []
    for k in range(len(z)):
        for rep in range(24):
            m = np.float32(rng.uniform(-3, 3) if rep % 3 else rng.uniform(-1e3, 1e3))
            scale = [1.0, 1e-3, 1e-7, 30.0][rep % 4]
            s = np.float32(abs(rng.normal()) * scale + 1e-9)
            for sign in (-1.0, 1.0):
                edge = np.float64(m) + sign * (np.float64(z[k]) * np.float64(s))   # numpy >= 2: float64 arithmetic
                t0 = np.float32(edge)
                for t in (t0, np.nextafter(t0, np.float32(np.inf)), np.nextafter(t0, np.float32(-np.inf))):
                    rows.append((m, s, t))
    rows += [(np.float32(1.0), np.float32(0.0), np.float32(1.0)), (np.float32(1.0), np.float32(0.0), np.float32(1.5)),
             (np.float32(0.5), np.float32(-0.2), np.float32(0.5)), (np.float32(0.5), np.float32(np.nan), np.float32(0.5)),
             (np.float32(np.nan), np.float32(0.1), np.float32(0.5)), (np.float32(0.5), np.float32(0.1), np.float32(np.nan)),
             (np.float32(0.0), np.float32(1e-38), np.float32(1e-39)), (np.float32(1e30), np.float32(1e25), np.float32(1.00001e30)),
             (np.float32(0.3), np.float32(np.inf), np.float32(0.9))]
    arr = np.array(rows, dtype=np.float32)
    m, s, t = arr[:, 0:1].copy(), arr[:, 1:2].copy(), arr[:, 2:3].copy()
    with np.errstate(invalid="ignore", over="ignore"):
        ref = om.auce(m, s, t)
    out = auce(m, s, t)
    n = float(m.size)
    assert np.array_equal(np.rint(np.asarray(out["coverage_values"]) * n), np.rint(np.asarray(ref["coverage_values"]) * n))


def test_auce_counts_with_zero_sigma_pixels(built_library):
    """Empty rays render with variance 0: sigma == 0 makes every interval the point m (count 0 unless t == m).  The
    prologue answers that from the ratio table (an infinite ratio) instead of the exact float64 search, which a warp
    would otherwise enter for every such pixel; counts must still equal numpy's, also where t == m."""
    from uncertainty_nerf_gs_b200.metrics import auce

    rng = np.random.default_rng(11)
    n = 200_000
    m = rng.random((n, 1)).astype(np.float32)
    s = (0.05 + 0.1 * rng.random((n, 1))).astype(np.float32)
    t = np.clip(m + s * rng.standard_normal((n, 1)).astype(np.float32), 0, 1).astype(np.float32)
    zero = rng.random(n) < 0.05
    s[zero] = 0.0
    same = zero & (rng.random(n) < 0.3)
    t[same] = m[same]                                   # sigma == 0 and t == m: inside every interval
    s[rng.random(n) < 0.001] = -0.1                      # a few negative sigmas: never inside
    with np.errstate(invalid="ignore", over="ignore"):
        ref = om.auce(m, s, t)
    out = auce(m, s, t)
    assert np.array_equal(np.rint(np.asarray(out["coverage_values"]) * n), np.rint(np.asarray(ref["coverage_values"]) * n))
