"""world_size-2 gloo test of the host-side sharding used by the multi-GPU LBA
path: landmark shards -> per-rank partial reduced systems (CPU
oracle stands in for the device kernels) -> one all-reduce(sum) == the full
system; frames shard with no exchange at all."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from orb_slam3_b200 import scenes


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    g, _ = scenes.lba_graph(8, 300, seed=2)
    sub, lm, ed = scenes.shard_graph(g, rank, world)
    # every landmark / edge lands on exactly one rank
    cnt = torch.zeros(len(g["mp_pos"]), dtype=torch.int32)
    cnt[torch.from_numpy(lm)] = 1
    dist.all_reduce(cnt)
    assert int(cnt.min()) == 1 and int(cnt.max()) == 1
    ecnt = torch.tensor([len(ed)])
    dist.all_reduce(ecnt)
    assert int(ecnt) == len(g["e_kf"])
    S, bs, chi = O.lba_reduced_system(scenes.lba_view(sub), 2.5)
    buf = torch.from_numpy(np.concatenate([S.ravel(), bs, [chi]]))
    dist.all_reduce(buf)  # the single exchange of a trial: [S | b_s | chi2]
    if rank == 0:
        Sf, bf, chif = O.lba_reduced_system(scenes.lba_view(g), 2.5)
        full = np.concatenate([Sf.ravel(), bf, [chif]])
        ret["err"] = float(np.abs(buf.numpy() - full).max() / np.abs(full).max())
    dist.destroy_process_group()


def test_landmark_shards_allreduce_to_full_system(oracle):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret["err"] < 1e-12


def test_frame_sharding_needs_no_exchange():
    """Streams map to ranks by s % world: a partition with no shared state."""
    streams = np.arange(16)
    for world in (1, 2, 4, 8):
        owned = [streams[streams % world == r] for r in range(world)]
        assert sorted(np.concatenate(owned)) == list(streams)
