"""CPU: host-side logic of the product package (no kernels): cut counts, z table, the numpy tails of
ause / auce fed with exact intermediate results, argument validation, binning, and the failure mode when
no CUDA device / library is present (the product path must fail loudly, never fall back)."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics as om
from oracle import splat as osp
from uncertainty_nerf_gs_b200 import binning, metrics, ops, synthetic


def test_cut_counts_follow_reference_float64_truncation():
    n = 640000
    cuts = metrics.ause_cut_counts(n)
    assert cuts.tolist() == om.ause_cut_counts(n)
    naive = np.array([n * (100 - i) // 100 for i in range(100)])
    assert (cuts != naive).sum() > 0          # SURVEY hard part 2: integer arithmetic would be wrong
    assert cuts[31] == 441599
    assert cuts[0] == n and (np.diff(cuts) <= 0).all()


def test_z_table_matches_reference_expression():
    z = metrics.z_values_host()
    assert z.shape == (99,) and (np.diff(z) < 0).all()
    assert np.array_equal(z, om.auce_z_values())
    assert z[0] == pytest.approx(2.5758293035489004) and z[-1] == pytest.approx(0.012533469508069276)


@pytest.mark.parametrize("err_type", ["mae", "mse", "rmse"])
@pytest.mark.parametrize("n", [10007, 57])
def test_ause_host_tail_reproduces_oracle_bit_for_bit(err_type, n):
    """Feed the numpy tail with float64 prefix sums computed on CPU: everything the device returns is a
    float64 sum, so this pins the host half of the drop-in exactly."""
    g = torch.Generator().manual_seed(n)
    unc = torch.clamp(0.1 * torch.rand(n, generator=g), min=0.03) ** 2
    err = torch.rand(n, generator=g)
    cuts = metrics.ause_cut_counts(n)
    es = torch.sort(err, stable=True).values
    eb = err[torch.sort(unc, stable=True).indices]

    def sums(v):
        c = torch.cat([torch.zeros(1, dtype=torch.float64), v.double().cumsum(0)])
        return c[cuts].numpy()

    r, e, v, a = metrics._ause_tail(metrics._prefix_means(sums(es), cuts, err_type),
                                    metrics._prefix_means(sums(eb), cuts, err_type))
    r0, e0, v0, a0 = om.ause(unc, err, err_type)
    assert np.array_equal(r, r0)
    np.testing.assert_allclose(e, e0, rtol=3e-7, equal_nan=True)   # torch's fp32 mean vs fp64 sum / n
    np.testing.assert_allclose(v, v0, rtol=3e-7, equal_nan=True)
    np.testing.assert_allclose(a, a0, rtol=1e-5, atol=1e-9, equal_nan=True)
    assert e.dtype == np.asarray(e0).dtype and v.dtype == v0.dtype and type(a) is type(a0)


def test_auce_host_tail_from_exact_histogram():
    g = torch.Generator().manual_seed(0)
    m = torch.rand(5000, 3, generator=g).numpy()
    s = np.clip(0.1 * torch.rand(5000, 3, generator=g).numpy(), 0.03, None).astype(np.float32)
    t = (m + s * torch.randn(5000, 3, generator=g).numpy()).astype(np.float32)
    ref = om.auce(m, s, t)
    z = metrics.z_values_host()
    inside = np.stack([(t >= m - zk * s) & (t <= m + zk * s) for zk in z])     # float64 promotion
    counts = inside.sum(axis=0).ravel()                                        # leading thresholds satisfied
    assert (np.diff(inside.astype(np.int8), axis=0) <= 0).all()                # monotone in k
    hist = np.bincount(counts, minlength=100)
    out = metrics._auce_from_hist(hist, float(s.astype(np.float64).sum()), float(m.size), z)
    assert list(out.keys()) == list(ref.keys())
    assert np.array_equal(out["coverage_values"], ref["coverage_values"])
    np.testing.assert_allclose(out["avg_length_values"], ref["avg_length_values"], rtol=1e-6)
    for k in ("auc_abs_error_values", "auc_length_values", "auc_neg_error_values"):
        np.testing.assert_allclose(out[k], ref[k], rtol=1e-6, atol=1e-12)


def test_ops_refuse_cpu_tensors_and_bad_shapes():
    z = torch.zeros(4, 48)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        ops.composite_rays(z, z, z, z, torch.zeros(4, 48, 3))
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        ops.reduce_members([torch.zeros(4, 3)], "std")
    with pytest.raises(ValueError):
        metrics.ause(torch.zeros(4), torch.zeros(5))
    with pytest.raises(ValueError):
        metrics.ause(torch.zeros(4), torch.zeros(4), "l1")


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from uncertainty_nerf_gs_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "libub200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()


def test_binning_ranges_and_order():
    sc = synthetic.splat_scene(300, 50, 70, seed=1)
    ids, bins = osp.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], 50, 70)
    tx, ty = binning.tile_grid(50, 70)
    assert bins.shape == (tx * ty, 2) and ids.dtype == torch.int32
    assert int(bins[0, 0]) == 0 and int(bins[-1, 1]) == ids.numel()
    assert bool((bins[1:, 0] == bins[:-1, 1]).all())
    # a Gaussian is listed in a tile iff its +-radius box overlaps the tile rectangle rule of gsplat
    g0 = int(ids[0])
    assert sc["radii"][g0] > 0


def test_synthetic_shapes():
    inp = synthetic.ray_samples(100, 48, seed=0)
    assert inp["rgb"].shape == (100, 48, 3) and inp["beta"].shape == (100, 48, 1)
    assert bool((inp["ends"] > inp["starts"]).all())
    p, s, g = synthetic.scoring_image(8, 9)
    assert p.shape == (8, 9, 3) and s.shape == (8, 9, 1) and float(s.min()) == pytest.approx(0.03)


def test_batched_host_tails_equal_per_image_numpy():
    """The [B, 100] tails must give, row by row, what the reference's 1-D numpy expressions give
    (np.trapz pairwise sums, float32-vs-float64 division of the oracle curve, Python max with NaNs)."""
    rng = np.random.default_rng(5)
    ratios = np.linspace(0, 1, 100, endpoint=False)
    o = (rng.random((6, 100)) * 3).astype(np.float32)
    v = (rng.random((6, 100)) * 3).astype(np.float32)
    o[1, 7] = np.nan
    v[2, 0] = np.nan
    v[3] = o[3] * 0.5                                  # oracle maximum wins -> float32 division
    oo, vv, aa = metrics._ause_tail_batch(o, v)
    for i in range(6):
        by = np.zeros(100)
        by[:] = v[i]                                   # ause.py:27-34 keeps this curve in a float64 array
        a, b = metrics._py_max(o[i]), metrics._py_max(by)
        mx = b if b > a else a
        want_o, want_v = np.array(o[i] / mx), np.array(by / mx)
        assert oo[i].dtype == want_o.dtype
        assert np.array_equal(oo[i], want_o, equal_nan=True) and np.array_equal(vv[i], want_v, equal_nan=True)
        assert np.array_equal(aa[i], np.trapz(want_v - want_o, ratios), equal_nan=True)
    hist = rng.integers(0, 500, (4, 100))
    n = hist.sum(1).astype(np.float64)
    ss = rng.random(4) * 1e3
    z = metrics.z_values_host()
    rows = metrics._auce_from_hist_batch(hist, ss, n, z)
    alphas = metrics._alphas()
    for i in range(4):
        cov = np.cumsum(hist[i][::-1])[::-1][1:] / n[i]
        err = cov - (1.0 - np.array(alphas))
        assert np.array_equal(rows[i]["coverage_values"], cov)
        assert rows[i]["auc_abs_error_values"] == np.trapz(y=np.abs(err), x=alphas)
        assert rows[i]["auc_length_values"] == np.trapz(y=list(2.0 * z * (ss[i] / n[i])), x=alphas)


def test_numa_binding_helper_is_harmless_without_nvml():
    from uncertainty_nerf_gs_b200 import pipeline

    before = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    assert pipeline.bind_host_thread_to_gpu(0) in (True, False)       # no GPU here: False, and nothing changes
    if before is not None and not torch.cuda.is_available():
        assert os.sched_getaffinity(0) == before


def test_ause_path_switch_and_select_refuses_cpu_tensors(monkeypatch):
    """The scorer takes the AUSE sums from the select kernels unless UB_AUSE_SORT=1 or a segment exceeds the
    2^24 keys the select path supports; like every op, the select wrapper has no CPU route."""
    monkeypatch.delenv("UB_AUSE_SORT", raising=False)
    assert metrics._use_select(1089480)
    assert not metrics._use_select((1 << 24) + 1)
    monkeypatch.setenv("UB_AUSE_SORT", "1")
    assert not metrics._use_select(1089480)
    monkeypatch.setenv("UB_AUSE_SORT", "0")
    assert metrics._use_select(640000)
    x = torch.rand(100)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        ops.cut_select_sums([(x, x, None)], [100], metrics.ause_cut_counts(100)[None, :])


def _random_packed(rng, b, n, c, nzp, kind):
    """A packed score buffer as the device would fill it: slice sums (non-increasing in the cut index, like prefix sums
    of non-negative errors at decreasing cut counts), prologue sums, an interval histogram that sums to n * c."""
    from uncertainty_nerf_gs_b200 import metrics as M

    packed = np.zeros(b * (400 + 5 + nzp))
    sums_v, psums, hist_v = M._packed_views(packed, b, nzp)
    cuts = M.ause_cut_counts(n)
    per_el = rng.random((b, 4, 1)) * 10.0 ** rng.integers(-6, 3, (b, 4, 1))
    sums_v[:] = cuts[None, None, :] * per_el * (1.0 + 0.3 * rng.random((b, 4, 100)))
    if kind == "oracle_wins":
        sums_v[:, 2:] *= 50.0                      # the float32 oracle maximum beats the by-uncertainty one
    if kind == "nan":
        sums_v[0, 0, 3] = np.nan                   # a NaN inside a curve: Python's max() semantics
        sums_v[-1, 2, 0] = np.nan                  # a leading NaN sticks
    if kind == "zero":
        sums_v[0] = 0.0                            # 0 / 0 curves
    psums[:] = rng.random((b, 5)) * n
    h = rng.multinomial(n * c, np.full(nzp, 1.0 / nzp), size=b).astype(np.int64)
    hist_v[:] = h.view(np.float64)
    return packed, cuts


@pytest.mark.parametrize("kind", ["plain", "oracle_wins", "nan", "zero"])
@pytest.mark.parametrize("b,n", [(1, 640000), (5, 1089480), (3, 37), (2, 100)])
def test_native_score_tail_equals_numpy_tail(kind, b, n):
    """ub_score_tail_host (csrc/score_tail.cu, a HOST function of the C ABI) against the numpy statement of the same
    tail, bit for bit: dtypes of the curves included (the oracle curve is float32 unless the by-uncertainty maximum
    wins), np.trapz's pairwise row sums included, NaN placement included."""
    from uncertainty_nerf_gs_b200 import metrics as M

    rng = np.random.default_rng(hash((kind, b, n)) % (1 << 32))
    c = 3
    nzp = len(M.z_values_host()) + 1
    packed, cuts = _random_packed(rng, b, n, c, nzp, kind)
    with np.errstate(all="ignore"):
        want = M._numpy_tail(packed, b, n, c, cuts)
    got = M._native_tail(packed, b, n, c, cuts)
    assert len(got) == len(want) == b
    for dg, dw in zip(got, want):
        assert list(dg.keys()) == list(dw.keys())
        for k in dw:
            g, w = np.asarray(dg[k]), np.asarray(dw[k])
            assert g.dtype == w.dtype, (k, g.dtype, w.dtype)
            assert g.shape == w.shape, k
            assert np.array_equal(np.atleast_1d(g).view(np.uint8), np.atleast_1d(w).view(np.uint8)) or \
                np.array_equal(g, w, equal_nan=True), k       # same bits (NaN payloads aside)


def test_native_score_tail_ragged_depth_views():
    """The depth modality: one channel, a different pixel count (hence different slice lengths) per view, a view with
    no valid pixel at all.  Native tail == the per-view numpy statement, bit for bit."""
    from uncertainty_nerf_gs_b200 import metrics as M

    rng = np.random.default_rng(5)
    lens = [40000, 0, 1237, 99, 7]
    b = len(lens)
    nzp = len(M.z_values_host()) + 1
    packed = np.zeros(b * (400 + 5 + nzp))
    sums_v, psums, hist_v = M._packed_views(packed, b, nzp)
    cuts = np.stack([M.ause_cut_counts(n) for n in lens])
    sums_v[:] = cuts[:, None, :] * (rng.random((b, 4, 1)) + 0.1) * (1.0 + 0.3 * rng.random((b, 4, 100)))
    psums[:] = rng.random((b, 5)) * np.asarray(lens)[:, None]
    h = np.stack([rng.multinomial(n, np.full(nzp, 1.0 / nzp)) for n in lens]).astype(np.int64)
    hist_v[:] = h.view(np.float64)
    with np.errstate(all="ignore"):
        want = M._numpy_depth_tail(packed, b, lens, cuts)
    got = M._native_tail(packed, b, np.asarray(lens, dtype=np.int64), 1, cuts, nll_key="nll_depth")
    for dg, dw in zip(got, want):
        assert list(dg.keys()) == list(dw.keys())
        for k in dw:
            g, w = np.asarray(dg[k]), np.asarray(dw[k])
            assert g.dtype == w.dtype and g.shape == w.shape, k
            assert np.array_equal(np.atleast_1d(g).view(np.uint8), np.atleast_1d(w).view(np.uint8)) or \
                np.array_equal(g, w, equal_nan=True), k
