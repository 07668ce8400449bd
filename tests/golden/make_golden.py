"""Generate the committed golden vectors.  Run in the dev container, where ``/root/reference`` is mounted:

    python tests/golden/make_golden.py

* ``ause_golden.npz`` / ``auce_golden.npz`` -- outputs of the REFERENCE's own
  ``nerfuncertainty/metrics/ause.py`` and ``auce.py`` (imported by file path, ``oracle/ref_loader.py``)
  on small seeded inputs.  ``ause`` runs with ``torch.sort`` forced to ``stable=True`` (the ranking
  contract); ``ause_nostable`` entries record the literal default call on tie-free inputs, where the two
  coincide.
* ``composite_golden.npz`` / ``reduce_golden.npz`` / ``laplace_golden.npz`` / ``splat_golden.npz`` -- outputs of the
  oracle restatements (the reference cannot be imported for these: nerfstudio / gsplat are absent), kept
  to detect drift of the oracle across torch versions.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import compositing as oc, laplace as ol, reduce as orc, ref_loader, splat as osp  # noqa: E402
from uncertainty_nerf_gs_b200 import synthetic  # noqa: E402


def main():
    loaded = ref_loader.load_reference_metrics()
    if loaded is None:
        raise SystemExit("/root/reference is not mounted: cannot regenerate reference-derived goldens")
    ref_ause, ref_auce = loaded
    torch.manual_seed(0)

    # ---- AUSE: ties at a variance floor, N = 10007 ----
    g = torch.Generator().manual_seed(1234)
    n = 10007
    unc = (torch.clamp(0.1 * torch.rand(n, generator=g), min=0.03) ** 2)
    err = torch.rand(n, generator=g) ** 2
    out = {"unc": unc.numpy(), "err": err.numpy()}
    with ref_loader.stable_torch_sort():
        for et in ("mae", "mse", "rmse"):
            ratio, e, v, a = ref_ause(unc, err, et)
            out[f"{et}_err"] = np.asarray(e, dtype=np.float64)
            out[f"{et}_err_by_var"] = np.asarray(v, dtype=np.float64)
            out[f"{et}_ause"] = np.float64(a)
    out["ratio"] = ratio
    # tie-free input: the reference's literal (unstable) call is reproducible
    unc_nt = (torch.randperm(n, generator=g).float() + 1.0) / n
    assert unc_nt.unique().numel() == n
    out["unc_notie"] = unc_nt.numpy()
    for et in ("mae", "rmse"):
        _, e, v, a = ref_ause(unc_nt, err, et)
        out[f"notie_{et}_err"] = np.asarray(e, dtype=np.float64)
        out[f"notie_{et}_err_by_var"] = np.asarray(v, dtype=np.float64)
        out[f"notie_{et}_ause"] = np.float64(a)
    np.savez_compressed(os.path.join(HERE, "ause_golden.npz"), **out)

    # ---- AUCE: 4096 x 3 float32 with exact hits, sigma = 0, NaN targets ----
    g = torch.Generator().manual_seed(99)
    m = torch.rand(4096, 3, generator=g)
    s = torch.clamp(0.1 * torch.rand(4096, 1, generator=g), min=0.03).repeat(1, 3)
    t = torch.clamp(m + s * torch.randn(4096, 3, generator=g), 0, 1)
    t[::17] = m[::17]
    s[5::29] = 0.0
    t[7::31] = float("nan")
    d = ref_auce(m.numpy(), s.numpy(), t.numpy())
    np.savez_compressed(os.path.join(HERE, "auce_golden.npz"), mean=m.numpy(), sigma=s.numpy(), target=t.numpy(),
                        **{k: np.asarray(v) for k, v in d.items()})

    # ---- oracle-derived goldens ----
    inp = synthetic.ray_samples(64, 48, seed=7)
    ref = oc.active_nerfacto_outputs(**inp)
    np.savez_compressed(os.path.join(HERE, "composite_golden.npz"),
                        **{f"in_{k}": v.numpy() for k, v in inp.items()},
                        **{f"out_{k}": v.numpy() for k, v in ref.items() if k != "density"},
                        out_weights=oc.get_weights(inp["density"], inp["deltas"]).numpy())

    outs = synthetic.member_renders(5, 9, 11, seed=3)
    red = orc.ensemble_reduce(outs)
    np.savez_compressed(os.path.join(HERE, "reduce_golden.npz"), **{f"out_{k}": v.numpy() for k, v in red.items()})

    lap = synthetic.laplace_head(257, 64, 3, 100, seed=5)
    theta = ol.posterior_samples(lap["mu_q"], lap["ggn"], lap["eps_draws"])
    mu, mu2, s2 = ol.sample_laplace(lap["x"], theta, 3, torch.sigmoid)
    np.savez_compressed(os.path.join(HERE, "laplace_golden.npz"), mean=mu.numpy(), mean2=mu2.numpy(),
                        sigma2=s2.numpy())

    sc = synthetic.splat_scene(400, 40, 56, seed=2, mean_scale_px=4.0)
    ids, bins = osp.bin_gaussians(sc["xys"], sc["depths"], sc["radii"], 40, 56)
    so = osp.active_splatfacto_outputs(sc["xys"], sc["depths"], sc["conics"], sc["opacities"], sc["rgbs"],
                                       sc["betas"], ids, bins, 40, 56, torch.tensor([0.1, 0.2, 0.3]))
    np.savez_compressed(os.path.join(HERE, "splat_golden.npz"), ids=ids.numpy(), bins=bins.numpy(),
                        **{f"out_{k}": v.numpy() for k, v in so.items()})
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
