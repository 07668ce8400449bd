"""GPU parity of the sort-free AUSE cut sums (``ub_cut_select_sums``, csrc/select_cuts.cu).

Checker 1: float64 torch on the CPU -- ``torch.sort(stable=True)`` + ``cumsum`` of the payload at the cut
counts, i.e. metrics/ause.py:10-34 before the division.  Checker 2: the sort path of this library
(``ub_segmented_sort`` + ``ub_cut_prefix_sums``), itself bit-exact against ``torch.sort(stable=True)``
(tests/test_gpu_scoring.py).  The select path must sum exactly the same *sets* of elements, so the float64
sums agree to summation order (1e-12 relative to the sum of magnitudes), far inside the 1e-5 contract.
"""
import numpy as np
import pytest
import torch

from uncertainty_nerf_gs_b200 import metrics as M

pytestmark = pytest.mark.gpu

TOL = 1e-12


def _reference(keys, pays, lens, cuts):
    """[nseg, len(pays), ncuts] float64 from a stable CPU sort."""
    out = np.zeros((len(lens), len(pays), cuts.shape[1]))
    lo = 0
    for s, n in enumerate(lens):
        k = keys[lo:lo + n]
        _, idx = torch.sort(k, stable=True)
        for v, pay in enumerate(pays):
            cs = torch.cat([torch.zeros(1, dtype=torch.float64), torch.cumsum(pay[lo:lo + n][idx].double(), 0)])
            out[s, v] = cs[torch.from_numpy(np.minimum(cuts[s], n))].numpy()
        lo += n
    return out


def _scale(pays, lens):
    """sum of |payload| per segment and payload: the yardstick of the summation-order tolerance."""
    out = np.zeros((len(lens), len(pays), 1))
    lo = 0
    for s, n in enumerate(lens):
        for v, pay in enumerate(pays):
            out[s, v, 0] = float(torch.nan_to_num(pay[lo:lo + n].double().abs(), nan=0.0, posinf=0.0).sum())
        lo += n
    return out + 1e-300


def _check(keys, pay0, pay1, lens, cuts, check_sort_path=True):
    from uncertainty_nerf_gs_b200 import ops

    kd, p0d, p1d = keys.cuda(), pay0.cuda(), pay1.cuda()
    got = ops.cut_select_sums([(kd, p0d, p1d), (p0d, p0d, None), (p1d, p1d, None)], lens, cuts).cpu().numpy()
    want = np.concatenate([_reference(keys, [pay0, pay1], lens, cuts), _reference(pay0, [pay0], lens, cuts),
                           _reference(pay1, [pay1], lens, cuts)], axis=1)
    scale = np.concatenate([_scale([pay0, pay1], lens), _scale([pay0], lens), _scale([pay1], lens)], axis=1)
    assert got.shape == want.shape
    both_nan = (np.isnan(got) & np.isnan(want)) | (got == want)          # NaN / inf poison the same cuts
    with np.errstate(invalid="ignore"):
        err = np.where(both_nan, 0.0, np.abs(got - want) / scale)
    assert not np.isnan(err).any(), "NaN pattern differs from the stable-sort reference"
    assert err.max() <= TOL, f"max scaled deviation {err.max():.3e}"
    if check_sort_path:
        total = sum(lens)
        both = torch.cat([kd, p0d, p1d])
        sorted_all, perm_all = ops.segmented_sort(both, list(lens) * 3, want_perm=True, want_keys=True)
        pv = perm_all[:total]
        viasort = ops.cut_prefix_sums([p0d, p1d, sorted_all[total:2 * total], sorted_all[2 * total:]],
                                      [pv, pv, None, None], lens, cuts).cpu().numpy()
        both_nan = (np.isnan(got) & np.isnan(viasort)) | (got == viasort)
        with np.errstate(invalid="ignore"):
            err2 = np.where(both_nan, 0.0, np.abs(got - viasort) / scale)
        assert not np.isnan(err2).any()
        assert err2.max() <= TOL


def _ause_cuts(lens):
    return np.stack([M.ause_cut_counts(n) for n in lens])


def _image(n, seed, floor=0.03):
    g = torch.Generator().manual_seed(seed)
    pred = torch.rand(n, 3, generator=g)
    std = torch.clamp(0.1 * torch.rand(n, generator=g), min=floor)     # SURVEY config 1: a big tie group
    gt = torch.clamp(pred + std[:, None] * torch.randn(n, 3, generator=g), 0, 1)
    d = pred - gt
    return std * std, d.abs().sum(-1), (d * d).sum(-1)


@pytest.mark.parametrize("n", [640000, 1089480])
def test_select_image_with_tie_floor(built_library, n):
    var, ae, se = _image(n, seed=n)
    _check(var, ae, se, [n], _ause_cuts([n]))


def test_select_smooth_batch(built_library):
    lens = [200000, 123457, 300001]
    parts = [_image(n, seed=i, floor=0.0) for i, n in enumerate(lens)]
    var, ae, se = (torch.cat([p[j] for p in parts]) for j in range(3))
    _check(var, ae, se, lens, _ause_cuts(lens))


@pytest.mark.parametrize("n", [1, 2, 5, 99, 100, 101, 257, 2048, 2049, 4097, 10007])
def test_select_small_segments(built_library, n):
    var, ae, se = _image(n, seed=100 + n)
    _check(var, ae, se, [n], _ause_cuts([n]))


def test_select_ragged_with_empty_segments(built_library):
    lens = [5000, 0, 1, 4096, 12345, 7, 0, 70001]
    parts = [_image(max(n, 1), seed=7 + i) for i, n in enumerate(lens)]
    var, ae, se = (torch.cat([p[j][:n] for p, n in zip(parts, lens)]) for j in range(3))
    _check(var, ae, se, lens, _ause_cuts(lens))


def test_select_all_keys_equal(built_library):
    n = 300000
    g = torch.Generator().manual_seed(3)
    var = torch.full((n,), 0.0009)
    ae, se = torch.rand(n, generator=g), torch.rand(n, generator=g) ** 2
    _check(var, ae, se, [n], _ause_cuts([n]))


def test_select_few_distinct_keys(built_library):
    """Quantised uncertainties: every cut lands inside a tie group; groups adjacent in float space share
    a bin (a cell that is large and *not* one tie group) -- the fallback of sorting the cell."""
    n = 400000
    g = torch.Generator().manual_seed(4)
    levels = torch.tensor([0.001, 0.0010000001, 0.00100000021, 0.5, 0.50000006, 7.0])
    var = levels[torch.randint(0, len(levels), (n,), generator=g)]
    ae, se = torch.rand(n, generator=g), torch.rand(n, generator=g)
    _check(var, ae, se, [n], _ause_cuts([n]))
    var8 = torch.round(torch.rand(n, generator=g) * 255) / 255             # 256 levels
    _check(var8, ae, se, [n], _ause_cuts([n]))


def test_select_special_values(built_library):
    n = 150000
    g = torch.Generator().manual_seed(5)
    var = torch.rand(n, generator=g)
    var[1::101] = 0.0
    var[2::101] = -0.0
    var[3::211] = float("nan")
    var[4::211] = float("inf")
    var[5::211] = -float("inf")
    var[6::97] = -var[6::97]
    var[7::303] = 1e-42
    ae, se = torch.rand(n, generator=g), torch.rand(n, generator=g)
    # the error keys carry specials too (NaN errors poison exactly the cuts that reach them)
    ae[11::5003] = float("nan")
    se[13::7001] = float("inf")
    _check(var, ae, se, [n], _ause_cuts([n]))


def test_select_arbitrary_cuts(built_library):
    lens = [50000, 33333]
    parts = [_image(n, seed=20 + i) for i, n in enumerate(lens)]
    var, ae, se = (torch.cat([p[j] for p in parts]) for j in range(3))
    rng = np.random.default_rng(0)
    cuts = np.stack([rng.integers(0, n + 1, size=37) for n in lens]).astype(np.int64)
    cuts[0, :4] = [0, lens[0], 1, lens[0] - 1]
    cuts[1, :4] = [7, 7, 7, 0]                                            # duplicates
    _check(var, ae, se, lens, cuts)


def test_scorer_select_equals_sort(built_library, monkeypatch):
    """score_rgb_batch through the select path and through the sort path: identical dictionaries up to the
    float32 rounding of a float64 sum whose order changed (curves 1e-6 relative, in practice bit-equal)."""
    b, h, w = 3, 120, 200
    g = torch.Generator().manual_seed(9)
    pred = torch.rand(b, h, w, 3, generator=g)
    std = torch.clamp(0.1 * torch.rand(b, h, w, 1, generator=g), min=0.03)
    gt = torch.clamp(pred + std * torch.randn(b, h, w, 3, generator=g), 0, 1)
    monkeypatch.setenv("UB_AUSE_SORT", "0")
    a = M.score_rgb_batch(pred.cuda(), gt.cuda(), std.cuda())
    monkeypatch.setenv("UB_AUSE_SORT", "1")
    c = M.score_rgb_batch(pred.cuda(), gt.cuda(), std.cuda())
    for da, dc in zip(a, c):
        assert da.keys() == dc.keys()
        for k in da:
            np.testing.assert_allclose(np.asarray(da[k], dtype=np.float64), np.asarray(dc[k], dtype=np.float64),
                                       rtol=1e-6, atol=1e-12, err_msg=k)


@pytest.mark.parametrize("seed", range(12))
def test_select_randomized_mixtures(built_library, seed):
    """Random mixtures of what breaks selection schemes: tie groups of every size (some a few ulps apart, some
    next to a continuous bulk), heavy tails over many octaves, negative keys, ragged segment lengths, random cuts."""
    rng = np.random.default_rng(1000 + seed)
    g = torch.Generator().manual_seed(1000 + seed)
    nseg = int(rng.integers(1, 5))
    lens = [int(rng.integers(1, 120000)) for _ in range(nseg)]
    keys = []
    for n in lens:
        kind = seed % 4
        x = torch.rand(n, generator=g)
        if kind == 0:      # log-uniform over 12 octaves with a floor tie group and its near neighbours
            x = torch.exp2(-12.0 * x)
            x = torch.clamp(x, min=float(rng.choice([1e-3, 0.05, 0.3])))
        elif kind == 1:    # a few tie values a few ulps apart inside a continuous bulk
            base = np.float32(0.37)
            ties = torch.tensor([base, np.nextafter(base, np.float32(1)), np.nextafter(base, np.float32(0)), 0.5, 0.0],
                                dtype=torch.float32)
            pick = torch.randint(0, 12, (n,), generator=g)
            x = torch.where(pick < len(ties), ties[pick.clamp(max=len(ties) - 1)], x)
        elif kind == 2:    # signed, heavy-tailed
            x = torch.randn(n, generator=g) ** 3 * 10.0 ** float(rng.integers(-6, 6))
        else:              # coarse quantisation: every key is a tie
            x = torch.round(x * float(rng.choice([3, 17, 1000]))) / 7.0
        keys.append(x)
    var = torch.cat(keys)
    total = sum(lens)
    ae = torch.rand(total, generator=g) * 10.0 ** float(rng.integers(-3, 3))
    se = torch.exp2(-20.0 * torch.rand(total, generator=g))
    ncut = int(rng.integers(1, 129))
    cuts = np.stack([rng.integers(0, n + 1, size=ncut) for n in lens]).astype(np.int64)
    _check(var, ae, se, lens, cuts, check_sort_path=seed % 3 == 0)


def test_select_background_of_exact_and_tiny_errors(built_library):
    """A render over a flat background: half the pixels have an error of exactly zero, a quarter an error 2^-40 below
    the largest one (outside the fixed-point window: the float64 path of sel_classify, whole warps of it), some are
    negative.  Exactness must not depend on which path a value takes."""
    n = 300000
    g = torch.Generator().manual_seed(11)
    var = torch.clamp(0.1 * torch.rand(n, generator=g), min=0.03) ** 2
    ae = torch.rand(n, generator=g)
    kind = torch.randint(0, 4, (n,), generator=g)
    ae = torch.where(kind <= 1, torch.zeros(()), ae)                    # exact zeros (and -0.0 below)
    ae = torch.where(kind == 2, ae * 2.0 ** -40, ae)                    # far below the window of the maximum
    ae[::7] = -ae[::7]                                                  # signed payloads / keys, -0.0 included
    se = ae * ae
    lens = [n]
    _check(var, ae, se, lens, _ause_cuts(lens))


@pytest.mark.parametrize("signed", [False, True])
def test_select_payloads_that_are_no_family_keys(built_library, signed):
    """A single family whose payloads are not the key array of any family: the fixed-point scale of the class sums then
    comes from a scan of each block's own payloads (different scales in different blocks) and so does the knowledge
    that a payload is negative.  Payload magnitudes vary by 10 orders across the segment, so the blocks do get
    different scales."""
    from uncertainty_nerf_gs_b200 import ops

    g = torch.Generator().manual_seed(21 + int(signed))
    lens = [150000, 70001]
    n = sum(lens)
    keys = torch.rand(n, generator=g)
    ramp = torch.exp(torch.linspace(-12.0, 11.0, n))                   # block to block: very different magnitudes
    p0 = torch.rand(n, generator=g) * ramp
    p1 = torch.rand(n, generator=g) / ramp
    if signed:
        p0 = p0 * torch.where(torch.rand(n, generator=g) < 0.5, -1.0, 1.0)
        p1[::3] = -p1[::3]
    cuts = _ause_cuts(lens)
    got = ops.cut_select_sums([(keys.cuda(), p0.cuda(), p1.cuda())], lens, cuts).cpu().numpy()
    want = _reference(keys, [p0, p1], lens, cuts)
    scale = _scale([p0, p1], lens)
    assert got.shape == want.shape
    err = np.abs(got - want) / scale
    assert err.max() <= TOL, f"max scaled deviation {err.max():.3e}"
    one = ops.cut_select_sums([(keys.cuda(), p0.cuda(), None)], lens, cuts).cpu().numpy()      # one payload array
    assert np.abs(one[:, 0] - want[:, 0]).max() / scale[:, 0].max() <= TOL
