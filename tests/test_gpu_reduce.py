"""GPU parity: fused K-member mean / spread (C ABI) vs the torch-CPU oracle restating
mcdropout_models.py:121-126 and ensemble_pipeline.py:159-190."""
import pytest
import torch

from oracle import reduce as orc
from uncertainty_nerf_gs_b200 import synthetic

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _cuda_list(outs):
    return [{k: v.cuda() for k, v in o.items()} for o in outs]


def _compare(out, ref):
    assert list(out.keys()) == list(ref.keys())
    for k in ref:
        atol = 1e-7 if ("std" in k or "var" in k) else 1e-6
        torch.testing.assert_close(out[k].cpu(), ref[k], rtol=RTOL, atol=atol, msg=lambda m: f"{k}: {m}")


@pytest.mark.parametrize("k_passes,hw", [(10, (37, 53)), (2, (16, 16)), (5, (101, 7)), (3, (1, 1))])
def test_mcdropout_reduce(built_library, k_passes, hw):
    from uncertainty_nerf_gs_b200.models.outputs import mcdropout_reduce

    outs = synthetic.member_renders(k_passes, *hw, seed=k_passes)
    _compare(mcdropout_reduce(_cuda_list(outs)), orc.mcdropout_reduce(outs))


@pytest.mark.parametrize("with_pred_std", [False, True])
def test_ensemble_reduce_both_branches(built_library, with_pred_std):
    from uncertainty_nerf_gs_b200.models.outputs import ensemble_reduce

    outs = synthetic.member_renders(5, 45, 61, seed=7, with_pred_std=with_pred_std)
    ref = orc.ensemble_reduce(outs)
    out = ensemble_reduce(_cuda_list(outs))
    _compare(out, ref)
    if with_pred_std:
        # the reference's order quirk: combined rgb_var / rgb_std are overwritten by plain member means
        member_mean = torch.stack([o["rgb_var"] for o in outs]).mean(0)
        torch.testing.assert_close(out["rgb_var"].cpu(), member_mean, rtol=RTOL, atol=1e-8)
        assert "rgb_var_epi" in out and "depth_var_alea" in out


def test_identical_members_give_zero_spread(built_library):
    from uncertainty_nerf_gs_b200 import ops

    x = torch.rand(1000, 3, device="cuda")
    mean, std = ops.reduce_members([x, x.clone(), x.clone()], "std")
    assert torch.equal(mean, x)
    assert float(std.abs().max()) == 0.0


def test_unaligned_and_odd_sizes(built_library):
    from uncertainty_nerf_gs_b200 import ops

    base = [torch.rand(4 * 1003 + 1, device="cuda") for _ in range(4)]
    members = [b[1:].reshape(-1, 1)[:1003 * 3].reshape(1003, 3) for b in base]  # 4-byte aligned only
    mean, std = ops.reduce_members(members, "std")
    st = torch.stack([m.cpu() for m in members])
    torch.testing.assert_close(mean.cpu(), st.mean(0), rtol=RTOL, atol=1e-6)
    torch.testing.assert_close(std.cpu(), st.std(0).mean(-1)[..., None], rtol=RTOL, atol=1e-7)


def test_single_member_std_is_nan_like_torch(built_library):
    from uncertainty_nerf_gs_b200 import ops

    x = torch.rand(64, 3, device="cuda")
    mean, std = ops.reduce_members([x], "std")
    assert torch.equal(mean, x)
    assert bool(torch.isnan(std).all())  # torch.std over one sample (unbiased) is NaN


def test_full_view_linearity(built_library):
    """BASELINE size (1297x840, K=5): mean is linear -- mean(a*x + c) == a*mean(x) + c for power-of-two a,
    std scales by |a| and ignores c (exact in binary floating point)."""
    from uncertainty_nerf_gs_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(0)
    ms = [torch.rand(840 * 1297, 3, generator=g, device="cuda") for _ in range(5)]
    mean, std = ops.reduce_members(ms, "std")
    mean2, std2 = ops.reduce_members([4.0 * m for m in ms], "std")
    assert torch.equal(mean2, 4.0 * mean)
    assert torch.equal(std2, 4.0 * std)
    ref = torch.stack(ms[:5]).double()
    torch.testing.assert_close(mean.double(), ref.mean(0), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(std.double(), ref.std(0).mean(-1, keepdim=True), rtol=1e-5, atol=1e-7)


def test_ensemble_branch_a_when_the_combined_entries_survive(built_library):
    """If a member dict lists rgb_var / rgb_std *before* rgb, the reference's loop overwrites nothing and the
    epistemic + aleatoric sum survives: the skipped-dead-work shortcut must not change that."""
    from uncertainty_nerf_gs_b200.models.outputs import ensemble_reduce

    outs = synthetic.member_renders(4, 23, 31, seed=11, with_pred_std=True)
    order = ["rgb_var", "rgb_std", "rgb", "accumulation", "depth", "expected_depth", "depth_var", "depth_std"]
    outs = [{k: o[k] for k in order if k in o} | {k: v for k, v in o.items() if k not in order} for o in outs]
    ref = orc.ensemble_reduce(outs)
    out = ensemble_reduce(_cuda_list(outs))
    _compare(out, ref)
    torch.testing.assert_close(out["rgb_var"].cpu(), (ref["rgb_var_epi"] + ref["rgb_var_alea"]), rtol=RTOL, atol=1e-8)
