"""Multi-process (gloo, world size 2) check of the data-parallel path on the CPU: images are sharded by batch, each rank runs
the encoder algorithm on its own shard with no collective, and the gathered embeddings equal the single-process result.  The
per-rank compute is the ORACLE here (no GPU in this test tier); bench.py drives the CUDA encoder with the same sharding."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import iuvl_b200 as ib
from iuvl_b200.sharding import all_gather_embeddings, shard_batch, shard_range


def test_shard_range_partitions_any_batch():
    for total in (0, 1, 2, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import sam_vit_oracle as orc
        torch.set_num_threads(2)
        cfg = ib.PRESETS["tiny64"]
        sd = ib.make_state_dict(cfg, 77, rel_std=0.1)
        x = ib.make_images(total, cfg, 5)                  # every rank builds the same global batch, then keeps its shard
        mine = shard_batch(x, rank, world)
        local = orc.encoder_forward_cfg(sd, mine, cfg) if mine.shape[0] else {
            k: torch.zeros((0,) + s) for k, s in (("res2", (128, 256, 256)), ("res3", (256, 128, 128)),
                                                  ("res4", (512, 64, 64)), ("res5", (1024, 32, 32)))}
        full = all_gather_embeddings(local, total)
        if rank == 0:
            torch.save({k: v for k, v in full.items()}, path)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_forward_matches_single_process(tmp_path):
    from oracle import sam_vit_oracle as orc
    total, world = 3, 2                                    # ragged: rank 0 gets 2 images, rank 1 gets 1
    path = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(world, _free_port(), total, path), nprocs=world, join=True)
    got = torch.load(path)
    cfg = ib.PRESETS["tiny64"]
    sd = ib.make_state_dict(cfg, 77, rel_std=0.1)
    ref = orc.encoder_forward_cfg(sd, ib.make_images(total, cfg, 5), cfg)
    for k in ref:
        assert got[k].shape == ref[k].shape
        assert ib.rel_l2(got[k], ref[k]) < 1e-6
