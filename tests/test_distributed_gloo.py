"""CPU, world_size 2 (gloo): the N > 1 path of the pipeline -- block view sharding, the single all_gather of
fixed-size per-view records, and the ordered host-side aggregation -- gives exactly the single-process
result (the cross-rank step is an ordered concat, so equality must be bit-for-bit)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from uncertainty_nerf_gs_b200 import pipeline

NUM_VIEWS = 7  # not a multiple of the world size: the last rank gets a short block


def _fake_view_record(view_id: int) -> np.ndarray:
    """A per-view metric dict with the reference's keys and deterministic pseudo-random contents."""
    rng = np.random.default_rng(1000 + view_id)
    d = {k: rng.random(100) for k in pipeline.CURVE_KEYS_100}
    d.update({k: rng.random(99) for k in pipeline.CURVE_KEYS_99})
    d.update({k: float(rng.random()) for k in pipeline.SCALAR_KEYS})
    depth = None
    if True:
        depth = {k: rng.random(100) for k in pipeline.CURVE_KEYS_100}
        depth.update({k: rng.random(99) for k in pipeline.CURVE_KEYS_99})
        depth.update({k: float(rng.random()) for k in pipeline.DEPTH_SCALAR_KEYS})
    extra = {k: float(rng.random()) for k in pipeline.IMAGE_KEYS + pipeline.TIMING_KEYS}
    return pipeline.pack_record(view_id, d, depth, extra)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = pipeline.shard_views(NUM_VIEWS, rank, world)
        local = np.stack([_fake_view_record(v) for v in mine]) if len(mine) else np.zeros((0, pipeline.RECORD_LEN))
        # blocks differ in length (7 views over 2 ranks; 3 views over 2 ranks leaves rank 1 with one, see below):
        # gather_records pads and strips by itself, with or without the row count given up front
        gathered = pipeline.gather_records(local, device=None)
        per = (NUM_VIEWS + world - 1) // world
        again = pipeline.gather_records(local, device=None, rows_per_rank=per)
        assert np.array_equal(gathered, again)
        few = pipeline.gather_records(local[:1] if rank == 0 else local[:0], device=None)      # an empty block
        assert few.shape == (1, pipeline.RECORD_LEN) and few[0, -1] == 0
        agg = pipeline.aggregate_records(gathered)
        np.save(os.path.join(out_dir, f"ids_{rank}.npy"), gathered[:, -1])
        np.save(os.path.join(out_dir, f"curve_{rank}.npy"), agg["err_var_rmse"])
        np.save(os.path.join(out_dir, f"scal_{rank}.npy"), np.array([agg[k] for k in pipeline.ALL_SCALAR_KEYS]))
        np.save(os.path.join(out_dir, f"dcurve_{rank}.npy"), agg["depth_coverage_values"])
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    single = np.stack([_fake_view_record(v) for v in range(NUM_VIEWS)])
    ref = pipeline.aggregate_records(single)
    for rank in range(world):
        ids = np.load(tmp_path / f"ids_{rank}.npy")
        assert ids.tolist() == list(range(NUM_VIEWS))           # ordered by view id on every rank
        assert np.array_equal(np.load(tmp_path / f"curve_{rank}.npy"), ref["err_var_rmse"])
        assert np.array_equal(np.load(tmp_path / f"scal_{rank}.npy"),
                              np.array([ref[k] for k in pipeline.ALL_SCALAR_KEYS]))
        assert np.array_equal(np.load(tmp_path / f"dcurve_{rank}.npy"), ref["depth_coverage_values"])


def test_shard_views_is_a_partition():
    for n, w in [(64, 8), (7, 2), (3, 4), (1, 1), (0, 2)]:
        seen = [v for r in range(w) for v in pipeline.shard_views(n, r, w)]
        assert seen == list(range(n))


def agg_keys(recs):
    return [k for k, v in pipeline.aggregate_records(recs).items() if isinstance(v, float)]


def test_record_roundtrip_and_reference_aggregation_semantics():
    recs = np.stack([_fake_view_record(v) for v in range(5)])
    vid, d = pipeline.unpack_record(recs[3])
    assert vid == 3 and d["err_mae"].shape == (100,) and d["coverage_values"].shape == (99,)
    assert d["depth_err_rmse"].shape == (100,) and "psnr" in d and "depth_nll" in d
    assert [k for k in agg_keys(recs)] == list(pipeline.ALL_SCALAR_KEYS)
    # a record without the depth / image-metric groups keeps them out of the aggregate
    rgb_only = np.stack([pipeline.pack_record(v, {**{k: np.zeros(100) for k in pipeline.CURVE_KEYS_100},
                                                  **{k: np.zeros(99) for k in pipeline.CURVE_KEYS_99},
                                                  **{k: 1.0 for k in pipeline.SCALAR_KEYS}}) for v in range(2)])
    only = pipeline.aggregate_records(rgb_only)
    assert "depth_nll" not in only and "psnr" not in only and "depth_err_mae" not in only and only["rgb_nll"] == 1.0
    agg = pipeline.aggregate_records(recs)
    # eval_uncertainty.py:1070-1077: float32 mean of python floats; :920-946: float64 curve sums / n
    vals = [pipeline.unpack_record(r)[1]["rgb_nll"] for r in recs]
    assert agg["rgb_nll"] == float(torch.mean(torch.tensor(vals)))
    total = np.zeros(100)
    for r in recs:
        total += pipeline.unpack_record(r)[1]["err_mse"]
    assert np.array_equal(agg["err_mse"], total / 5)
