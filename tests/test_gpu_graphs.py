"""GPU: CUDA-graph replay of the step (``pipeline.GraphedViews`` / ``GraphedScore``) returns exactly what the
kernel-by-kernel path returns -- same kernels, same buffers' worth of arithmetic, only the enqueue differs."""
import numpy as np
import pytest
import torch

from uncertainty_nerf_gs_b200 import synthetic

pytestmark = pytest.mark.gpu


def _same(a, b):
    assert list(a.keys()) == list(b.keys())
    for k in b:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k]), equal_nan=True), k


def test_graphed_views_equal_eager_evaluation(built_library):
    from uncertainty_nerf_gs_b200 import pipeline

    h, w, S, M = 24, 40, 48, 3
    sets = [[synthetic.ray_samples(h * w, S, seed=100 * v + i, device="cuda") for i in range(M)] for v in range(3)]
    gts = [synthetic.scoring_image(h, w, seed=v, device="cuda")[2] for v in range(7)]
    order = [(v % 3, v) for v in range(7)]                   # member set v % 3 with ground truth v
    want = [pipeline.evaluate_view(sets[s], gts[g], h, w, rays_per_chunk=256) for s, g in order]
    gv = pipeline.GraphedViews(h, w, rays_per_chunk=256)
    timers = []
    pend = [gv.launch(sets[s], gts[g], timers=timers) for s, g in order[:2]]
    got = [p.finish() for p in pend]
    prev = None
    for s, g in order[2:]:                                   # streamed: launch i+1 before finishing i
        cur = gv.launch(sets[s], gts[g], timers=timers)
        if prev is not None:
            got.append(prev.finish())
        prev = cur
    got.append(prev.finish())
    assert len(got) == len(want)
    for a, b in zip(got, want):
        _same(a, b)
    torch.cuda.synchronize()
    assert len(timers) == 7 and all(t0.elapsed_time(t1) > 0 and cnt == M for t0, t1, cnt in timers)
    # results handed out earlier are not views of the recycled pinned buffers
    _same(got[0], want[0])


def test_graphed_score_equals_eager(built_library):
    from uncertainty_nerf_gs_b200 import metrics, pipeline

    p, s, g = synthetic.scoring_image(120, 200, seed=3, device="cuda")
    want = metrics.score_rgb_batch(p, g, s)[0]
    gs = pipeline.GraphedScore(p, g, s)
    first = gs.launch().finish()[0]
    _same(first, want)
    g2 = torch.clamp(g + 0.01, 0, 1)
    g.copy_(g2)                                              # same buffers, new contents: the graph re-reads them
    second = gs.launch().finish()[0]
    _same(second, metrics.score_rgb_batch(p, g, s)[0])
    _same(first, want)


def test_host_view_evaluator_with_and_without_the_deltas_stream(built_library):
    """The end-to-end entry (pinned host ray samples in, metric dict out) equals the device-resident evaluation of the
    same view; with ``derive_deltas`` the deltas are neither copied nor read and -- for ray samples whose deltas are
    ``ends - starts``, nerfstudio's invariant -- nothing changes, bit for bit."""
    from uncertainty_nerf_gs_b200 import pipeline

    h, w, S, M = 24, 40, 48, 3
    members = [synthetic.ray_samples(h * w, S, seed=11 + i) for i in range(M)]
    for m in members:
        m["deltas"] = m["ends"] - m["starts"]
    gt = synthetic.scoring_image(h, w, seed=3)[2]
    want = pipeline.evaluate_view([{k: v.cuda() for k, v in m.items()} for m in members], gt.cuda(), h, w, rays_per_chunk=256)
    host = [{k: v.pin_memory() for k, v in m.items()} for m in members]
    full = pipeline.HostViewEvaluator(M, h * w, S, h, w, "cuda:0")
    lean = pipeline.HostViewEvaluator(M, h * w, S, h, w, "cuda:0", derive_deltas=True)
    assert full.h2d_bytes - lean.h2d_bytes == M * h * w * S * 4
    for ev in (full, lean, lean):
        got = ev(host, gt.pin_memory(), 256)
        for k in want:
            assert np.array_equal(np.asarray(got[k]), np.asarray(want[k]), equal_nan=True), k
