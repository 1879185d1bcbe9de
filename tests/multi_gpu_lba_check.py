"""torchrun helper (not collected by pytest): landmark-sharded LBA over NCCL must
reproduce the single-GPU solve.  Usage: torchrun --nproc-per-node N tests/multi_gpu_lba_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam3_b200 import scenes  # noqa: E402
from orb_slam3_b200.optimizer import LocalBundleAdjustment  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g, _ = scenes.lba_graph(20, 3000, seed=2)
    single = LocalBundleAdjustment(device=local)(scenes.lba_view(g))
    lba = LocalBundleAdjustment(device=local)
    uid = [LocalBundleAdjustment.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    lba.init_comm(rank, world, uid[0])
    sub, lm, ed = scenes.shard_graph(g, rank, world)
    r = lba(scenes.lba_view(sub))
    assert r["iterations"] == single["iterations"] and r["stats"]["trials"] == single["stats"]["trials"], \
        (r["iterations"], single["iterations"])
    step = np.abs(single["mp_pos"] - g["mp_pos"]).max()
    assert np.abs(r["kf_pose"] - single["kf_pose"]).max() < 1e-7
    assert np.abs(r["mp_pos"] - single["mp_pos"][lm]).max() < 1e-6 * max(step, 1.0)
    assert np.allclose(r["chi2"], single["chi2"][ed], rtol=1e-6, atol=1e-6)
    assert abs(r["stats"]["chi2_final"] - single["stats"]["chi2_final"]) < 1e-8 * single["stats"]["chi2_final"]
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_LBA_OK world=%d trials=%d" % (world, r["stats"]["trials"]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
