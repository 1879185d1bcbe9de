#!/usr/bin/env python3
"""bench.py -- BASELINE.json metric on synthetic data.

python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the front-end hot path over one batch of synthetic
1280x720 frames (8 levels, 2000 features): ORB extract (+ projection match once
the matcher stage is enabled).  One process per GPU (torchrun for N>1), frames
are independent so ranks share nothing: weak scaling, no data-path collective.

  value : frames/s with the batch already resident in HBM (CUDA events on the
          launching stream, max over ranks)
  e2e   : frames/s through the host-buffer C ABI (pinned host frames -> H2D ->
          kernels -> D2H keypoints+descriptors), same batches
  roofline / cpu_baseline : see DESIGN.md "Measurement"
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, NFEAT, NLEVELS = 720, 1280, 2000, 8
METRIC = "frames/sec ORB extract+match 1280x720x8lvl"


def level_pixels():
    inv = [np.float32(1.0)]
    sc = np.float32(1.0)
    for _ in range(1, NLEVELS):
        sc = np.float32(np.float64(sc) * np.float64(np.float32(1.2)))
        inv.append(np.float32(1.0) / sc)
    return [int(np.rint(np.float32(W) * s)) * int(np.rint(np.float32(H) * s)) for s in inv]


def make_frames(n, seed0):
    from orb_slam3_b200.synth import synth_frame, shifted_frame
    rng = np.random.default_rng(seed0)
    frames = [synth_frame(H, W, seed0)]
    for t in range(1, n):
        dx, dy = rng.integers(-8, 9, size=2)
        frames.append(shifted_frame(frames[-1], int(dx), int(dy), seed0 * 1000 + t))
    return np.stack(frames)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_reference_fps(frames, threads, seconds_budget=20.0):
    """The oracle (CPU restatement of the reference path) on the host cores:
    one extractor instance per std::thread, frames are independent (the
    reference runs one thread per extractor, Frame.cc:122-125)."""
    from oracle import oracle as O
    O.build()
    fps1, _, dt1 = O.extract_throughput(frames, NFEAT, 1, 2)
    iters = int(min(64, max(2, seconds_budget * fps1)))
    return O.extract_throughput(frames, NFEAT, threads, iters)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    frames = make_frames(8, 1)
    K, Wm = args.steps, args.warmup
    # one step = a bounded sample: `threads` x per_thread frames
    vals = []
    for i in range(Wm + K):
        fps, done, dt = cpu_reference_fps(frames, threads, seconds_budget=1.0)
        if i >= Wm:
            vals.append((fps, done, dt))
    fps = sum(v[1] for v in vals) / sum(v[2] for v in vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": K, "warmup": Wm, "ms_per_step": 1e3 * sum(v[2] for v in vals) / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": "configs[1]: 1280x720 8-level 2000-feature ORB extract, synthetic stream",
                   "frames_per_step": int(vals[0][1]), "stages": "extract"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d frames/step over %d threads (oracle C++ port, -O3 x86-64-v3)"
                                   % (vals[0][1], threads)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=64, help="frames per step per GPU")
    ap.add_argument("--pool", type=int, default=4, help="distinct batches rotated through (L2 defeat)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from orb_slam3_b200.extractor import ORBextractor

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: orb_slam3_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B, K, Wm, POOL = args.batch, args.steps, max(args.warmup, 3), args.pool

    # ---- synthetic streams: POOL distinct batches of B frames (this rank's cameras)
    uniq = make_frames(min(B, 16), 1000 * rank + 1)
    host_pool = []
    for p in range(POOL):
        t = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
        for b in range(B):
            src = uniq[(b + p) % len(uniq)]
            t[b] = torch.from_numpy(np.roll(src, (p * 7 + b // len(uniq), b // len(uniq) * 3), (0, 1)).copy())
        host_pool.append(t)
    dev_pool = [t.cuda() for t in host_pool]
    ext = ORBextractor(NFEAT, 1.2, NLEVELS, 20, 7, device=local_rank)
    ext._lib.orb_extract_batch_device  # noqa: B018  (fail early if the ABI is missing)
    tstream = torch.cuda.Stream()  # non-default stream so the engine launches where the events are recorded
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    def step_device(i):
        d = dev_pool[i % POOL]
        ext.extract_batch_device(d.data_ptr(), B, H, W, W, H * W, stream=stream)

    cap = ext.cap
    from orb_slam3_b200._lib import KP_DTYPE, check, ptr
    import ctypes as C
    out_k = torch.empty((B, cap * 28), dtype=torch.uint8).pin_memory()
    out_d = torch.empty((B, cap * 32), dtype=torch.uint8).pin_memory()
    out_n = np.zeros(B, np.int32)
    out_m = np.zeros(B, np.int32)

    def step_host(i):
        t = host_pool[i % POOL]
        arr = (C.c_void_p * B)(*[t.data_ptr() + b * H * W for b in range(B)])
        check(ext._lib.orb_extract_batch(ext._h, B, arr, H, W, W, None, C.c_void_p(out_k.data_ptr()),
                                         C.c_void_p(out_d.data_ptr()), cap, ptr(out_n), ptr(out_m)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        ext.synchronize()
        barrier()
        ms = e0.elapsed_time(e1)
        return ms

    # ---- device-resident metric
    for i in range(Wm):
        step_device(i)
    launches0 = ext.kernel_launches()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev = timed(step_device, K)
    launches = ext.kernel_launches() - launches0
    # ---- end to end through the host-buffer ABI (wall clock brackets the D2H sync)
    for i in range(Wm):
        step_host(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        step_host(i)
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    sampler.stop_flag = True
    sampler.join(timeout=2)
    nkp = int(out_n.sum())

    # ---- per-kernel times for the roofline (events around each launch; separate pass)
    ext.set_profiling(True)
    ext.stage_times(reset=True)
    for i in range(K):
        step_device(i)
    ext.synchronize()
    st = ext.stage_times(reset=True)
    ext.set_profiling(False)

    tens = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tens, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = tens.tolist()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"
        P = sum(level_pixels())
        # algorithmic bytes per frame and per kernel (DESIGN.md "Kernels")
        alg = {
            "pyramid": W * H + (P - W * H),          # read level 0, write levels 1..7
            "fast": P + 0,                           # one read of the pyramid
            "blur": 2 * P,                           # read + write blurred pyramid
            "octree": 38000 * 8,                     # candidate records
            "describe": NFEAT * (28 + 32) + NFEAT * (31 * 31 + 512),
        }
        kern = {k: v for k, v in st.items() if k in alg and v[1] > 0}
        dom = max(kern, key=lambda k: kern[k][0])
        dom_ms_per_launch = kern[dom][0] / (kern[dom][1] if dom != "pyramid" else kern[dom][1] / (NLEVELS - 1))
        achieved = alg[dom] * B / (dom_ms_per_launch * 1e-3) / 1e9
        total_ms = sum(v[0] for v in kern.values())
        fps_dev = world * B * K / (ms_dev * 1e-3)
        fps_e2e = world * B * K / (ms_e2e * 1e-3)
        # bounded CPU sample on rank 0 at N=1
        cpu = None
        if world == 1:
            threads = os.cpu_count() or 1
            fps_cpu, done, dt = cpu_reference_fps(uniq, threads, seconds_budget=10.0)
            fps_1, done1, dt1 = cpu_reference_fps(uniq, 1, seconds_budget=4.0)
            cpu = {"value": fps_cpu, "unit": "frames/s", "cores": threads, "kind": "port",
                   "sample": "%d frames over %d threads in %.1fs; single thread: %.1f frames/s"
                             % (done, threads, dt, fps_1)}
        line = {
            "metric": METRIC, "value": fps_dev, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[1]: 1280x720 8-level 2000-feature ORB extract, synthetic stream",
                       "frames_per_step_per_gpu": B, "stages": "extract",
                       "l2": "inputs+pyramids %.0f MB per rotation > 126 MB L2 (%d batches rotated)"
                             % (POOL * B * (W * H + 2 * P) / 1e6, POOL),
                       "keypoints_last_step": nkp},
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": B * H * W,
                    "d2h_bytes_per_step": int(B * (NFEAT + 4 * NLEVELS) * 60 + 8 * B)},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src,
                         "share_of_step": kern[dom][0] / total_ms,
                         "stage_ms_per_step": {k: v[0] / K for k, v in st.items()}},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
