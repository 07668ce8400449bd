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
    return pipeline.pack_record(view_id, d)


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
        per = (NUM_VIEWS + world - 1) // world
        local = np.stack([_fake_view_record(v) for v in mine]) if len(mine) else np.zeros((0, pipeline.RECORD_LEN))
        # all_gather_into_tensor needs equal shapes: pad short blocks with records marked view id = -1
        pad = np.zeros((per - local.shape[0], pipeline.RECORD_LEN))
        pad[:, -1] = -1
        gathered = pipeline.gather_records(np.concatenate([local, pad]), device=None)
        gathered = gathered[gathered[:, -1] >= 0]
        agg = pipeline.aggregate_records(gathered)
        np.save(os.path.join(out_dir, f"ids_{rank}.npy"), gathered[:, -1])
        np.save(os.path.join(out_dir, f"curve_{rank}.npy"), agg["err_var_rmse"])
        np.save(os.path.join(out_dir, f"scal_{rank}.npy"), np.array([agg[k] for k in pipeline.SCALAR_KEYS]))
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
                              np.array([ref[k] for k in pipeline.SCALAR_KEYS]))


def test_shard_views_is_a_partition():
    for n, w in [(64, 8), (7, 2), (3, 4), (1, 1), (0, 2)]:
        seen = [v for r in range(w) for v in pipeline.shard_views(n, r, w)]
        assert seen == list(range(n))


def test_record_roundtrip_and_reference_aggregation_semantics():
    recs = np.stack([_fake_view_record(v) for v in range(5)])
    vid, d = pipeline.unpack_record(recs[3])
    assert vid == 3 and d["err_mae"].shape == (100,) and d["coverage_values"].shape == (99,)
    agg = pipeline.aggregate_records(recs)
    # eval_uncertainty.py:1070-1077: float32 mean of python floats; :920-946: float64 curve sums / n
    vals = [pipeline.unpack_record(r)[1]["rgb_nll"] for r in recs]
    assert agg["rgb_nll"] == float(torch.mean(torch.tensor(vals)))
    total = np.zeros(100)
    for r in recs:
        total += pipeline.unpack_record(r)[1]["err_mse"]
    assert np.array_equal(agg["err_mse"], total / 5)
