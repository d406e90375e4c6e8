"""Multi-GPU host logic on CPU: world_size 2, gloo. The data path has no collective (scenes are independent);
what is distributed is the scene -> rank assignment and the max-over-ranks of the measured time."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audiblelight_b200 import sharding, workload


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.shard_scene_indices(4, rank, world)
        specs = [workload.c3_scene_spec(i) for i in mine]
        local_seconds = sum(s.duration for s in specs)
        local_bytes = sum(workload.algorithmic_bytes(s) for s in specs)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        t = sharding.max_over_ranks(10.0 + rank)       # the slower rank defines the step time
        total_seconds = sharding.sum_over_ranks(local_seconds)
        total_bytes = sharding.sum_over_ranks(float(local_bytes))
        out[rank] = dict(mine=mine, gathered=gathered, t=t, total_seconds=total_seconds, total_bytes=total_bytes)
    finally:
        dist.destroy_process_group()


def test_scene_sharding_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert sorted(out.keys()) == [0, 1]
    all_scenes = sorted(out[0]["mine"] + out[1]["mine"])
    assert all_scenes == list(range(8))                       # disjoint cover of the job
    assert out[0]["gathered"] == out[1]["gathered"] == [out[0]["mine"], out[1]["mine"]]
    assert out[0]["t"] == out[1]["t"] == 11.0                # max over ranks
    assert out[0]["total_seconds"] == out[1]["total_seconds"] == 8 * 60.0
    expect_bytes = float(sum(workload.algorithmic_bytes(workload.c3_scene_spec(i)) for i in range(8)))
    assert out[0]["total_bytes"] == expect_bytes
    for r in range(world):
        assert all(sharding.owner_of_scene(i, world) == r for i in out[r]["mine"])


def test_shard_indices_validation():
    assert sharding.shard_scene_indices(3, 1, 4) == [1, 5, 9]
    with pytest.raises(ValueError):
        sharding.shard_scene_indices(3, 4, 4)
    assert sharding.max_over_ranks(3.5) == 3.5  # not initialised -> identity


def test_workload_specs_are_deterministic_and_shaped():
    a, b = workload.c3_scene_spec(7), workload.c3_scene_spec(7)
    assert [(e.n_audio, e.n_irs, e.snr, e.start) for e in a.events] == [(e.n_audio, e.n_irs, e.snr, e.start) for e in b.events]
    assert len(a.events) == 9 and sum(e.n_irs > 1 for e in a.events) == 3
    for e in a.events:
        assert 2 * 24000 <= e.n_audio <= 10 * 24000
        if e.n_irs > 1:
            assert e.n_irs == int(round(10.0 * e.n_audio / 24000)) + 1
    assert abs(workload.algorithmic_bytes(workload.c1_scene_spec()) - 9.024e6) < 1e3   # C1 = 9.0 MB
