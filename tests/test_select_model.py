"""Invariants the select kernels (csrc/select_cuts.cu) rely on, restated in numpy and checked on the CPU.

These are not the kernels (the GPU parity suite, tests/test_gpu_select.py, checks those against a stable sort);
they pin the integer arithmetic of the binning so that a change of a constant that breaks a bound fails here,
without a GPU: (1) the fine-bin allocation never exceeds 8192 bins and caps a bin at ~n/2048 keys unless keys
tie, (2) the bin index is monotone in the order-preserving key and gives every heavy tie value a bin of its
own, (3) the exponent read off a coarse bin is the exponent of its keys, (4) the fixed-point limbs add up
exactly.
"""
import numpy as np
import pytest

COARSE, LOW_BITS, FINE = 4096, 20, 8192


def order_key(x: np.ndarray) -> np.ndarray:
    """ub_common.cuh:sort_key_from_float."""
    x = np.where(x == 0, np.float32(0), x).astype(np.float32)
    b = x.view(np.uint32)
    k = np.where(b & 0x80000000, ~b, b | 0x80000000).astype(np.uint32)
    return np.where(np.isnan(x), np.uint32(0xFFFFFFFE), k)


def allocate(keys_u: np.ndarray):
    """sel_alloc: (first fine bin, shift) per coarse bin."""
    n = len(keys_u)
    target = max(1, (n + 2047) // 2048)
    cnt = np.bincount(keys_u >> LOW_BITS, minlength=COARSE)
    lg = np.zeros(COARSE, dtype=np.int64)
    big = cnt > target
    q = (cnt[big] + target - 1) // target
    lg[big] = np.minimum(np.ceil(np.log2(q)).astype(np.int64), LOW_BITS)
    nsub = np.where(cnt > 0, 1 << lg, 0)
    base = np.concatenate([[0], np.cumsum(nsub)[:-1]])
    return base, LOW_BITS - lg, int(nsub.sum()), target


def bins_of(keys_u, base, shift, heavy=()):
    c = keys_u >> LOW_BITS
    b = base[c] + ((keys_u & ((1 << LOW_BITS) - 1)) >> shift[c])
    for h in heavy:
        b = b + (h < keys_u) + (h <= keys_u)
    return b


def _samples(rng, kind, n):
    if kind == "uniform":
        return rng.random(n, dtype=np.float32)
    if kind == "loguniform":
        return np.exp2(-30 * rng.random(n)).astype(np.float32)
    if kind == "signed":
        return (rng.standard_normal(n) ** 3).astype(np.float32)
    if kind == "floor":
        return np.maximum(rng.random(n, dtype=np.float32) * 0.01, np.float32(0.0009))
    if kind == "specials":
        x = rng.standard_normal(n).astype(np.float32)
        x[::7] = np.nan
        x[1::11] = np.inf
        x[2::13] = -np.inf
        x[3::5] = 0.0
        x[4::17] = -0.0
        return x
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["uniform", "loguniform", "signed", "floor", "specials"])
@pytest.mark.parametrize("n", [1, 100, 2047, 2049, 50000, 640000])
def test_fine_bin_allocation_bounds(kind, n):
    rng = np.random.default_rng(n)
    u = order_key(_samples(rng, kind, n))
    base, shift, nfine, target = allocate(u)
    assert nfine <= FINE
    b = bins_of(u, base, shift)
    assert b.min() >= 0 and b.max() < nfine
    # monotone in the key
    o = np.argsort(u, kind="stable")
    assert (np.diff(b[o]) >= 0).all()
    # a bin that holds more than `target` keys cannot be split further (single key value) or sits in a coarse
    # bin that already has the maximum resolution its count pays for (<= 2 * target keys per bin on average)
    cnt = np.bincount(b, minlength=nfine)
    for fb in np.nonzero(cnt > 2 * target)[0][:50]:
        ks = u[b == fb]
        c = ks[0] >> LOW_BITS
        coarse_cnt = int((u >> LOW_BITS == c).sum())
        nsub = 1 << (LOW_BITS - shift[c])
        assert nsub >= coarse_cnt / target / 2 or len(np.unique(ks)) == 1


def test_heavy_values_own_a_bin_and_order_is_kept():
    rng = np.random.default_rng(0)
    x = np.maximum(rng.random(200000, dtype=np.float32) * 0.01, np.float32(0.0009))   # 9 % tie at the floor
    u = order_key(x)
    base, shift, nfine, _ = allocate(u)
    heavy = [int(order_key(np.array([0.0009], dtype=np.float32))[0])]
    b = bins_of(u, base, shift, heavy)
    assert b.max() < nfine + 2 * len(heavy)
    o = np.argsort(u, kind="stable")
    assert (np.diff(b[o]) >= 0).all()
    hb = np.unique(b[u == heavy[0]])
    assert len(hb) == 1 and (u[b == hb[0]] == heavy[0]).all()          # the tie value, and nothing else


def test_exponent_of_a_coarse_bin():
    rng = np.random.default_rng(1)
    x = np.concatenate([_samples(rng, "signed", 5000), _samples(rng, "loguniform", 5000), -_samples(rng, "loguniform", 5000)])
    x = x[np.isfinite(x)]
    u = order_key(x)
    c = (u >> LOW_BITS).astype(np.int64)
    ex = (np.where(c >= 2048, c, ~c) >> 3) & 0xFF                       # sel_alloc
    assert (ex == ((np.abs(x).view(np.uint32) >> 23) & 0xFF)).all()


def test_fixed_point_limbs_are_exact():
    """sel_window / sel_limbs: q = mantissa << (exponent + 33 - emax) for values within 2^33 of the largest magnitude,
    three 19-bit limbs (each negated for a negative value) added as integers into one of 8 columns; a column word sees
    at most 4096 adds (16384 keys per block, 8 columns, a lane's own keys plus queued ones), so it cannot overflow, and
    limbs . (1, 2^19, 2^38) . 2^(emax - 150 - 33) is the exact sum."""
    import fractions

    WINDOW, LIMB, COLS = 33, 19, 8
    rng = np.random.default_rng(2)
    v = (rng.standard_normal(16384) * np.exp2(rng.integers(-30, 1, 16384))).astype(np.float32)
    v = v[v != 0]
    bits = v.view(np.uint32).astype(np.int64)
    be = (bits >> 23) & 0xFF
    emax = int(be.max())
    ok = (be != 0) & (be + WINDOW >= emax)
    assert ok.mean() > 0.97
    q = (((bits & 0x7FFFFF) | 0x800000).astype(object) << (be + WINDOW - emax).clip(0).astype(object))
    sign = np.where(bits >> 31, -1, 1)
    mask = (1 << LIMB) - 1
    col = np.arange(len(v)) % COLS
    words = np.zeros((3, COLS), dtype=object)
    for qi, sg, c, good in zip(q, sign, col, ok):
        if not good:
            continue
        qi = int(qi)
        assert qi < 1 << (24 + WINDOW)
        for j in range(3):
            words[j, c] += int(sg) * ((qi >> (LIMB * j)) & mask)
    assert all(abs(int(w)) < 2 ** 31 for w in words.reshape(-1))
    assert 2 * (16384 // COLS) * mask < 2 ** 31                              # the kernel's static_assert
    total = sum(int(words[j].sum()) << (LIMB * j) for j in range(3))
    exact = sum((fractions.Fraction(float(x)) for x in v[ok]), fractions.Fraction(0))
    assert fractions.Fraction(total) * fractions.Fraction(2) ** (emax - 150 - WINDOW) == exact
