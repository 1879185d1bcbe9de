#!/usr/bin/env python3
"""bench.py -- BASELINE.json metric on synthetic data.

python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the front-end hot path over one batch (128 per GPU) of synthetic
1280x720 frames (BASELINE.json configs[1]: 8 levels, 2000 features): per frame
ORB extract -> SearchByProjection against the previous frame (th 15, rotation
check) -> SearchByProjection against ~3000 local map points (th 3).  One
process per GPU (torchrun for N>1); frames are independent so ranks share
nothing on this path: weak scaling, no data-path collective.

  value : frames/s with the batch already resident in HBM (CUDA events on the
          launching stream, max over ranks)
  e2e   : frames/s through the host-buffer C ABI: pinned host frames -> H2D ->
          kernels -> D2H keypoints + descriptors, then host views -> match ->
          D2H assignments
  lba   : LocalBA LM iterations/s (optimize(10) incl. rejected trials) on the
          synthetic config-4 graph (N=1) / config-5 graph sharded by landmark
          with one NCCL all-reduce per trial (N>1)
  roofline / cpu_baseline : see DESIGN.md "Measurement"
"""
import argparse
import ctypes as C
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
N_LOCAL_EXTRA = 1000          # random map points on top of one per keypoint (~3000 total)
TH_LAST, TH_LOCAL = 15.0, 3.0
METRIC = "frames/sec ORB extract+match 1280x720x8lvl"
WORKLOAD = ("configs[1]: 1280x720 8-level 2000-feature ORB extract + SearchByProjection "
            "(last frame th=15 + ~3000 local map points th=3), synthetic stream")


def level_pixels():
    inv = [np.float32(1.0)]
    sc = np.float32(1.0)
    for _ in range(1, NLEVELS):
        sc = np.float32(np.float64(sc) * np.float64(np.float32(1.2)))
        inv.append(np.float32(1.0) / sc)
    return [int(np.rint(np.float32(W) * s)) * int(np.rint(np.float32(H) * s)) for s in inv]


def make_frames(n, seed0):
    """A synthetic stream: frame t+1 = frame t shifted by (dx,dy) in [-8,8]^2."""
    from orb_slam3_b200.synth import synth_frame, shifted_frame
    rng = np.random.default_rng(seed0)
    frames, shifts = [synth_frame(H, W, seed0)], [(0, 0)]
    for t in range(1, n):
        dx, dy = (int(v) for v in rng.integers(-8, 9, size=2))
        frames.append(shifted_frame(frames[-1], dx, dy, seed0 * 1000 + t))
        shifts.append((dx, dy))
    return np.stack(frames), shifts


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region.  NVML in-process (pynvml) so that no
    nvidia-smi process has to be spawned next to the measurement; nvidia-smi is the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.sm, self.mx, self.reasons = [], [], set()
        self.samples = 0
        self.stop_flag = False
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
        except Exception:
            self.nv = None

    def _sample_nvml(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        table = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        for name, bit in table.items():
            if r & bit:
                self.reasons.add(name)
        self.samples += 1

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
        if not out:
            return
        r = [c.strip() for c in out.split(",")]
        try:
            self.sm.append(float(r[1]))
            self.mx.append(float(r[2]))
        except ValueError:
            pass
        for nm, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
            if v.lower().startswith("active"):
                self.reasons.add(nm)
        self.samples += 1

    def run(self):
        while not self.stop_flag:
            try:
                if self.nv:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.01 if self.nv else 0.2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": self.samples,
                "source": "nvml" if self.nv else "nvidia-smi"}


# --------------------------------------------------------------------------- CPU arm
def cpu_extract_fps(frames, threads, seconds_budget):
    from oracle import oracle as O
    fps1, _, _ = O.extract_throughput(frames, NFEAT, 1, 2)
    iters = int(min(64, max(2, seconds_budget * fps1)))
    return O.extract_throughput(frames, NFEAT, threads, iters)


_BEST_THREADS = {}


def best_cpu_threads():
    """Thread count that gives the reference arm its best throughput on this host:
    shared boxes often expose more logical CPUs than the process can really use."""
    if "n" in _BEST_THREADS:
        return _BEST_THREADS["n"]
    from oracle import oracle as O
    ncpu = os.cpu_count() or 1
    frames, _ = make_frames(4, 1)
    cands = sorted({c for c in (4, 8, 16, 32, 64, ncpu) if c <= ncpu})
    best, best_fps = ncpu, 0.0
    for c in cands:
        fps, _, _ = O.extract_throughput(frames, NFEAT, c, 3)
        if fps > best_fps:
            best, best_fps = c, fps
    _BEST_THREADS["n"] = best
    return best


def cpu_match_seconds_per_frame(frames, shifts, n_pairs=3):
    """Oracle matchers (one thread, like the Tracking thread) on a few frame pairs."""
    from oracle import oracle as O
    from orb_slam3_b200 import scenes
    ex = O.OracleExtractor(NFEAT)
    feats = [ex.extract(f)[:2] for f in frames[:n_pairs + 1]]
    t_m = 0.0
    for t in range(1, n_pairs + 1):
        (ka, da), (kb, db) = feats[t - 1], feats[t]
        cur, last, Tcw = scenes.last_frame_scene(ka, da, kb, db, W, H, shifts[t], seed=2 * t)
        F, mps = scenes.local_map_scene(kb, db, W, H, N_LOCAL_EXTRA, seed=7 * t)
        t1 = time.perf_counter()
        O.match_project_last(cur, last, Tcw, TH_LAST)
        O.match_project_local(F, mps, TH_LOCAL, 0.8)
        t_m += time.perf_counter() - t1
    return t_m / n_pairs


def cpu_lba(K, L, seed=0):
    from oracle import oracle as O
    from orb_slam3_b200 import scenes
    g, _ = scenes.lba_graph(K, L, seed=seed)
    r = O.lba_solve(scenes.lba_view(g))
    st = r["stats"]
    return {"config": "%d KF x %d landmarks, %d edges" % (K, L, len(g["e_kf"])), "iterations": st["iterations"],
            "trials": st["trials"], "seconds": st["ms_total"] / 1e3,
            "value": st["trials"] / (st["ms_total"] / 1e3), "unit": "LM iterations/s", "threads": 1}


def run_reference(args):
    """The reference's own CPU implementation of the path: the oracle port (the
    reference cannot be compiled here), all host threads for the per-frame work
    (one extractor instance per thread), single thread for LBA like g2o."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    threads = best_cpu_threads()
    frames, shifts = make_frames(8, 1)
    K, Wm = args.steps, args.warmup
    t_match = cpu_match_seconds_per_frame(frames, shifts)
    vals = []
    for i in range(Wm + K):
        fps, done, dt = cpu_extract_fps(frames, threads, seconds_budget=1.0)
        if i >= Wm:
            vals.append((done, dt))
    done = sum(v[0] for v in vals)
    secs = sum(v[1] for v in vals)
    # matching runs on the same threads: add its per-frame cost to each thread's frame time
    t_ext_frame_thread = secs * threads / max(done, 1)
    fps = threads / (t_ext_frame_thread + t_match)
    lba = cpu_lba(50, 20000)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": K, "warmup": Wm, "ms_per_step": 1e3 * secs / max(K, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": int(vals[0][0]) if vals else 0},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d frames/step over %d std::threads (best of 4..nproc; oracle C++ port, -O3 x86-64-v3); "
                                   "extract %.2f ms + match %.2f ms per frame per thread"
                                   % (vals[0][0] if vals else 0, threads, 1e3 * t_ext_frame_thread, 1e3 * t_match)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "lba": lba,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- GPU arm
class Workload:
    """POOL distinct batches of B frames with everything the matchers need."""

    def __init__(self, B, POOL, rank, device, stream, share=None):
        import torch
        from orb_slam3_b200 import scenes
        from orb_slam3_b200.extractor import ORBextractor
        from orb_slam3_b200.matcher import ORBmatcher
        self.torch, self.B, self.POOL = torch, B, POOL
        self.ext = ORBextractor(NFEAT, 1.2, NLEVELS, 20, 7, device=device)
        self.m_last = ORBmatcher(0.9, True, device=device)
        self.m_local = ORBmatcher(0.8, True, device=device)
        self.stream = stream
        for m in (self.m_last, self.m_local):
            m.set_stream(stream)
        cap = self.ext.cap
        self.cap = cap
        if share is not None:
            # extra e2e worker: same inputs, own engine handles and output buffers
            self.uniq, self.host_pool, self.dev_pool, self.meta = share.uniq, share.host_pool, None, share.meta
            self.sf, self.sf2 = share.sf, share.sf2
            self._alloc_outputs(torch, B, cap)
            return
        uniq, shifts = make_frames(min(B, 16) + 1, 1000 * rank + 1)
        self.uniq = uniq
        nu = len(uniq) - 1
        self.host_pool, self.dev_pool, self.meta = [], [], []
        self.sf = scenes.scale_factors()
        self.sf2 = (self.sf * self.sf).astype(np.float32)
        # one untimed extraction of the distinct frames (GPU) to build the match inputs
        res = self.ext.extract_batch(list(uniq))
        scene_cache = {}
        for p in range(POOL):
            t = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
            idx = [1 + (b + p) % nu for b in range(B)]        # frame idx[b]; its predecessor is idx[b]-1
            for b in range(B):
                t[b] = torch.from_numpy(uniq[idx[b]])
            self.host_pool.append(t)
            self.dev_pool.append(t.cuda())
            entry = {"n": [], "cur": [], "last": [], "T": [], "F": [], "mps": [], "d_last": [], "d_mps": [],
                     "d_taken": []}
            for b in range(B):
                i = idx[b]
                if i not in scene_cache:  # only len(uniq)-1 distinct frames exist
                    _, kb, db = res[i]
                    _, ka, da = res[i - 1]
                    cur, last, Tcw = scenes.last_frame_scene(ka, da, kb, db, W, H, shifts[i], seed=2 * i)
                    F, mps = scenes.local_map_scene(kb, db, W, H, N_LOCAL_EXTRA, seed=7 * i)
                    scene_cache[i] = (
                        len(kb), cur, last, Tcw, F, mps,
                        {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in last._keep.items()},
                        {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in mps._keep.items()},
                        (torch.from_numpy(cur._keep[5]).cuda(), torch.from_numpy(F._keep[5]).cuda()))
                n_i, cur, last, Tcw, F, mps, dl, dm, dtk = scene_cache[i]
                entry["n"].append(n_i)
                entry["cur"].append(cur); entry["last"].append(last); entry["T"].append(Tcw)
                entry["F"].append(F); entry["mps"].append(mps)
                entry["d_last"].append(dl); entry["d_mps"].append(dm); entry["d_taken"].append(dtk)
            entry["T"] = np.stack(entry["T"])
            self.meta.append(entry)
        self.d_assign_last = torch.empty((B, cap), dtype=torch.int32, device="cuda")
        self.d_assign_local = torch.empty((B, cap), dtype=torch.int32, device="cuda")
        self._alloc_outputs(torch, B, cap)

    def _alloc_outputs(self, torch, B, cap):
        self.out_k = torch.empty((B, cap * 28), dtype=torch.uint8).pin_memory()
        self.out_d = torch.empty((B, cap * 32), dtype=torch.uint8).pin_memory()
        self.out_n = np.zeros(B, np.int32)
        self.out_m = np.zeros(B, np.int32)
        self.nmatch_last = np.zeros(B, np.int32)
        self.nmatch_local = np.zeros(B, np.int32)
        self._dev_views = {}

    # views whose frame side points at the extractor's device results (built once per pool entry)
    def _device_views(self, p):
        if p in self._dev_views:
            return self._dev_views[p]
        from orb_slam3_b200.views import orb_frame_view, orb_lastframe_view, orb_mappoint_view
        kp, ds, _, _, cap = self.ext.device_results()
        e = self.meta[p]
        curs, lasts, Fs, Ms = [], [], [], []
        for b in range(self.B):
            for which, src, lst in ((0, e["cur"][b], curs), (1, e["F"][b], Fs)):
                v = orb_frame_view()
                C.memmove(C.byref(v), C.byref(src), C.sizeof(v))
                v.n = e["n"][b]
                v.keys = kp + b * cap * 28
                v.desc = ds + b * cap * 32
                v.u_right = None
                v.kp_taken = e["d_taken"][b][which].data_ptr()
                v.scale_factors, v.level_sigma2 = self.sf.ctypes.data, self.sf2.ctypes.data
                lst.append(v)
            lv = orb_lastframe_view()
            lv.n = e["last"][b].n
            for k in ("has_mp", "has_obs", "world_pos", "desc", "octave", "angle"):
                setattr(lv, k, e["d_last"][b][k].data_ptr())
            lasts.append(lv)
            mv = orb_mappoint_view()
            mv.n = e["mps"][b].n
            for k in ("track_in_view", "is_bad", "has_obs", "proj_x", "proj_y", "proj_xr", "scale_level",
                      "view_cos", "depth", "desc"):
                setattr(mv, k, e["d_mps"][b][k].data_ptr())
            Ms.append(mv)
        a1 = [self.d_assign_last.data_ptr() + 4 * b * self.cap for b in range(self.B)]
        a2 = [self.d_assign_local.data_ptr() + 4 * b * self.cap for b in range(self.B)]
        self._dev_views[p] = (curs, lasts, Fs, Ms, a1, a2)
        return self._dev_views[p]

    def enable_side_streams(self, main_stream):
        """The two projection matchers only depend on the extraction, not on each other: give each
        its own stream so their latency-bound kernels overlap (joined back before the next step)."""
        torch = self.torch
        self.main_stream = main_stream
        self.side = (torch.cuda.Stream(), torch.cuda.Stream())
        self.m_last.set_stream(self.side[0].cuda_stream)
        self.m_local.set_stream(self.side[1].cuda_stream)

    def step_device(self, i):
        p = i % self.POOL
        d = self.dev_pool[p]
        self.ext.extract_batch_device(d.data_ptr(), self.B, H, W, W, H * W, stream=self.stream)
        curs, lasts, Fs, Ms, a1, a2 = self._device_views(p)
        side = getattr(self, "side", None)
        if side:
            for st in side:
                st.wait_stream(self.main_stream)
        r1, _ = self.m_last.project_last_batch(curs, lasts, self.meta[p]["T"], TH_LAST, on_device=True,
                                               assign_ptrs=a1)
        r2, _ = self.m_local.project_local_batch(Fs, Ms, TH_LOCAL, on_device=True, assign_ptrs=a2)
        if side:  # the next extraction overwrites the keypoints/descriptors the matchers read
            for st in side:
                self.main_stream.wait_stream(st)
        self.nmatch_last, self.nmatch_local = r1, r2  # filled when the asynchronous batches complete

    def finish_device(self):
        self.m_last.synchronize()
        self.m_local.synchronize()

    def step_host(self, i):
        from orb_slam3_b200._lib import check, ptr
        from orb_slam3_b200.views import orb_frame_view
        p = i % self.POOL
        t = self.host_pool[p]
        B, cap = self.B, self.cap
        arr = (C.c_void_p * B)(*[t.data_ptr() + b * H * W for b in range(B)])
        check(self.ext._lib.orb_extract_batch(self.ext._h, B, arr, H, W, W, None, C.c_void_p(self.out_k.data_ptr()),
                                              C.c_void_p(self.out_d.data_ptr()), cap, ptr(self.out_n),
                                              ptr(self.out_m)))
        e = self.meta[p]
        curs, Fs = [], []
        for b in range(B):
            for src, lst in ((e["cur"][b], curs), (e["F"][b], Fs)):
                v = orb_frame_view()
                C.memmove(C.byref(v), C.byref(src), C.sizeof(v))
                v.n = int(self.out_n[b])
                v.keys = self.out_k.data_ptr() + b * cap * 28      # the keypoints just downloaded
                v.desc = self.out_d.data_ptr() + b * cap * 32
                lst.append(v)
        r1, _ = self.m_last.project_last_batch(curs, e["last"], e["T"], TH_LAST)
        r2, _ = self.m_local.project_local_batch(Fs, e["mps"], TH_LOCAL)
        self.nmatch_last, self.nmatch_local = r1, r2

    def e2e_bytes(self):
        e = self.meta[0]
        h2d = self.B * H * W
        d2h = int(self.B * (NFEAT + 4 * NLEVELS) * 60 + 8 * self.B)
        for b in range(self.B):
            n = e["n"][b]
            h2d += 2 * n * 61 + e["last"][b].n * 50 + e["mps"][b].n * 59
            d2h += 2 * n * 4
        return h2d, d2h


def run_lba_gpu(rank, world, device):
    """LocalBA on the GPU: configs 4 and 5 at N=1, config 5 sharded by landmark for N>1."""
    import torch
    import torch.distributed as dist
    from orb_slam3_b200 import scenes
    from orb_slam3_b200.optimizer import LocalBundleAdjustment
    out = {}
    lba = LocalBundleAdjustment(device=device)
    configs = [("config4", 50, 20000), ("config5", 200, 80000)] if world == 1 else [("config5", 200, 80000)]
    if world > 1:
        uid = [LocalBundleAdjustment.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        lba.init_comm(rank, world, uid[0])
    for name, K, L in configs:
        g, _ = scenes.lba_graph(K, L, seed=0)
        sub = g if world == 1 else scenes.shard_graph(g, rank, world)[0]
        gv = scenes.lba_view(sub)
        best = None
        for rep in range(4):  # the first run warms up allocations
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            st = lba(gv)["stats"]
            if rep > 0 and (best is None or st["ms_total"] < best["ms_total"]):
                best = st
        ms = torch.tensor([best["ms_total"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_total = float(ms.item())
        out[name] = {
            "config": "%d KF x %d landmarks, %d edges%s" % (K, L, len(g["e_kf"]),
                                                            "" if world == 1 else ", landmark shards + ncclAllReduce"),
            "iterations": best["iterations"], "trials": best["trials"], "ms_total": ms_total,
            "value": best["trials"] / (ms_total * 1e-3), "unit": "LM iterations/s",
            "chi2_initial": best["chi2_initial"], "chi2_final": best["chi2_final"],
            "stage_ms": {k: best[k] for k in ("ms_linearize", "ms_schur", "ms_solve", "ms_update")},
            "schur_gflops_sparse": best["schur_flops"] * best["trials"] / max(best["ms_schur"], 1e-9) / 1e6,
            "n_pose_pairs": best["n_pairs"],
        }
    return out


def run_stereo_gpu(device, tstream, pairs=64, reps=5, cpu=True):
    """SURVEY.md 8(f-1), config 3 shape: `pairs` rectified 1280x720 stereo pairs per step; both eyes
    extracted by their own handle, then Frame::ComputeStereoMatches on the device-resident results."""
    import torch
    from orb_slam3_b200.extractor import ORBextractor
    from orb_slam3_b200.stereo import StereoMatcher
    from orb_slam3_b200.synth import synth_frame, stereo_right
    bf, b = 386.0, 386.0 / 700.0
    uniq = 4
    lefts = [synth_frame(H, W, 9000 + i) for i in range(uniq)]
    rights = [stereo_right(l, 9100 + i, disparities=(6, 24, 12)) for i, l in enumerate(lefts)]
    dl = torch.from_numpy(np.stack([lefts[i % uniq] for i in range(pairs)])).cuda()
    dr = torch.from_numpy(np.stack([rights[i % uniq] for i in range(pairs)])).cuda()
    el = ORBextractor(NFEAT, 1.2, NLEVELS, 20, 7, device=device)
    er = ORBextractor(NFEAT, 1.2, NLEVELS, 20, 7, device=device)
    sm = StereoMatcher(device)
    cs = tstream.cuda_stream

    def step():
        el.extract_batch_device(dl.data_ptr(), pairs, H, W, W, H * W, stream=cs)
        er.extract_batch_device(dr.data_ptr(), pairs, H, W, W, H * W, stream=cs)
        sm.compute_batch(el, er, pairs, bf, b, on_device=True, cuda_stream=cs)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ms_match = 0.0
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms_all = e0.elapsed_time(e1) / reps
    for _ in range(reps):  # the matcher alone, device time between its own events
        sm.compute_batch(el, er, pairs, bf, b, on_device=True, cuda_stream=cs)
        ms_match += sm.last_ms()
    ms_match /= reps
    kept, ur, dp = sm.compute_batch(el, er, pairs, bf, b)
    out = {"config": "%d stereo pairs 1280x720, 2000 features per eye, disparity bands 6/24/12 px" % pairs,
           "pairs_per_s": pairs / (ms_all * 1e-3), "ms_per_step": ms_all, "unit": "stereo pairs/s (2 x extract + ComputeStereoMatches)",
           "stereo_match_us_per_pair": 1e3 * ms_match / pairs, "matches_per_pair": float(kept.mean()),
           "gpu_launches_per_step": 3}
    if cpu:
        from oracle import oracle as O
        exl, exr = O.OracleExtractor(NFEAT), O.OracleExtractor(NFEAT)
        t0 = time.perf_counter()
        kl, d1, _ = exl.extract(lefts[0])
        kr, d2, _ = exr.extract(rights[0])
        t_ext = time.perf_counter() - t0
        pl = [exl.level_image(l) for l in range(NLEVELS)]
        pr = [exr.level_image(l) for l in range(NLEVELS)]
        t0 = time.perf_counter()
        for _ in range(5):
            n_ref, ur_ref, dp_ref, _ = O.stereo_match(kl, d1, kr, d2, pl, pr, bf, b)
        t_sm = (time.perf_counter() - t0) / 5
        n0 = len(kl)
        out["cpu_baseline"] = {"stereo_match_ms_per_pair": 1e3 * t_sm, "extract2_ms_per_pair": 1e3 * t_ext, "threads": 1,
                               "kind": "port", "parity_pair0": bool(np.array_equal(ur[0, :n0], ur_ref) and
                                                                    np.array_equal(dp[0, :n0], dp_ref))}
    return out


def run_pose_gpu(device, frames=128, edges=1000, reps=5, cpu=True):
    """`frames` independent Optimizer::PoseOptimization problems (one per tracked frame,
    `edges` matched MapPoints each, 80 % stereo, 10 % gross mismatches) per step, one kernel launch."""
    from orb_slam3_b200 import scenes
    from orb_slam3_b200.optimizer import PoseOptimization
    po = PoseOptimization(device)
    uniq = 16
    base = [scenes.pose_scene(edges, seed=100 + i)[0] for i in range(uniq)]
    views = [base[i % uniq] for i in range(frames)]
    po.batch(views)
    ms_dev, t0 = 0.0, time.perf_counter()
    for _ in range(reps):
        inl, pose, outs, stats = po.batch(views)
        ms_dev += po.last_ms()
    ms_wall = (time.perf_counter() - t0) * 1e3 / reps
    ms_dev /= reps
    out = {"config": "%d frames x %d edges (80%% stereo, 10%% mismatches), 4 rounds x optimize(10)" % (frames, edges),
           "value": frames / (ms_dev * 1e-3), "unit": "PoseOptimization calls/s (kernel, CUDA events)",
           "e2e_value": frames / (ms_wall * 1e-3), "e2e_unit": "calls/s through the host-buffer C ABI (H2D + kernel + D2H)",
           "ms_per_step": ms_dev, "gpu_launches_per_step": 1,
           "lm_trials_per_frame": float(stats[:, 2].mean()), "inliers_per_frame": float(inl.mean())}
    if cpu:
        from oracle import oracle as O
        t0 = time.perf_counter()
        refs = [O.pose_optimize(v) for v in base]
        t_cpu = (time.perf_counter() - t0) / uniq
        ok = all(np.array_equal(outs[i], refs[i]["outlier"]) and
                 np.abs(pose[i] - refs[i]["pose"]).max() < 1e-7 for i in range(uniq))
        out["cpu_baseline"] = {"ms_per_call": 1e3 * t_cpu, "threads": 1, "kind": "port", "parity_first_%d" % uniq: bool(ok)}
    return out


def run_frustum_gpu(device, cpu=True):
    """Frame::isInFrustum over the local map of one frame (5000 points, the
    Tracking::SearchLocalPoints shape) and over 2^20 points (streaming rate of the kernel)."""
    from orb_slam3_b200 import scenes
    from orb_slam3_b200.frustum import FrustumCuller
    fc = FrustumCuller(device)
    out = {}
    for name, n, reps in (("local_map_5000", 5000, 20), ("stream_1M", 1 << 20, 5)):
        v, _ = scenes.frustum_scene(n, seed=3)
        k, o = fc.isInFrustum(v)
        ms_dev, t0 = 0.0, time.perf_counter()
        for _ in range(reps):
            k, o = fc.isInFrustum(v, out=o)
            ms_dev += fc.last_ms()
        ms_wall = (time.perf_counter() - t0) * 1e3 / reps
        ms_dev /= reps
        e = {"points": n, "in_view": int(k), "kernel_us": 1e3 * ms_dev, "host_call_us": 1e3 * ms_wall,
             "kernel_GBps_algorithmic": n * 57 / (ms_dev * 1e-3) / 1e9, "bytes_per_point": 57}
        if cpu:
            from oracle import oracle as O
            t0 = time.perf_counter()
            for _ in range(3):
                k_ref, ref = O.is_in_frustum(v)
            e["cpu_port_us"] = (time.perf_counter() - t0) / 3 * 1e6
            e["parity"] = bool(k_ref == k and all(np.array_equal(o[f], ref[f]) for f in ref))
        out[name] = e
    return out


def run_bow_gpu(device, cpu=True):
    """Frame::ComputeBoW = DBoW2 transform(levelsup 4) of one frame's 2000 descriptors
    against a resident synthetic 10-ary depth-5 vocabulary (the ORB vocabulary is 10-ary, depth 6)."""
    from orb_slam3_b200 import scenes
    from orb_slam3_b200.bow import ORBVocabulary
    from orb_slam3_b200.extractor import ORBextractor
    from orb_slam3_b200.synth import synth_frame
    voc = scenes.synth_vocabulary(10, 5, seed=3)
    gv = ORBVocabulary(voc, device)
    ext = ORBextractor(NFEAT, 1.2, NLEVELS, 20, 7, device=device)
    _, kps, desc = ext(synth_frame(H, W, 4242))
    got = gv.transform_extracted(ext, 0, 4)
    reps, ms_dev, t0 = 20, 0.0, time.perf_counter()
    for _ in range(reps):
        got = gv.transform_extracted(ext, 0, 4)
        ms_dev += gv.last_ms()
    ms_wall = (time.perf_counter() - t0) * 1e3 / reps
    out = {"config": "%d descriptors, vocabulary %d nodes (k=10, L=5), levelsup 4" % (len(desc), voc.n_nodes),
           "kernels_us": 1e3 * ms_dev / reps, "call_us_descriptors_on_device": 1e3 * ms_wall, "gpu_launches_per_call": 2,
           "words": int(len(got["bow_ids"])), "fv_nodes": int(len(got["fv_node_ids"]))}
    if cpu:
        from oracle import oracle as O
        t0 = time.perf_counter()
        for _ in range(3):
            ref = O.bow_transform(voc, desc, 4)
        out["cpu_port_us"] = (time.perf_counter() - t0) / 3 * 1e6
        out["parity"] = bool(all(np.array_equal(got[k], ref[k]) for k in ("bow_ids", "bow_vals", "fv_node_ids", "fv_ptr", "fv_idx")))
    return out


def _guarded(fn, *a, **k):
    """The extra legs must never cost the headline line."""
    try:
        return fn(*a, **k)
    except Exception as e:  # noqa: BLE001
        return {"error": "%s: %s" % (type(e).__name__, e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=128, help="frames per step per GPU")
    ap.add_argument("--pool", type=int, default=3, help="distinct batches rotated through (L2 defeat)")
    ap.add_argument("--e2e-workers", type=int, default=4, help="host threads feeding the GPU in the e2e leg")
    ap.add_argument("--no-lba", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-stereo", action="store_true", help="skip the 8(f) legs (ComputeStereoMatches, PoseOptimization, isInFrustum, ComputeBoW)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly one JSON line: libraries that write to fd 1 (NCCL prints its version banner
    # there) are pointed at stderr, the line itself goes to a duplicate of the original stdout
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: orb_slam3_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B, K, Wm, POOL = args.batch, args.steps, max(args.warmup, 3), args.pool

    tstream = torch.cuda.Stream()  # non-default stream: the engines launch where the events are recorded
    torch.cuda.set_stream(tstream)
    wl = Workload(B, POOL, rank, local_rank, tstream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def launches_now():
        return wl.ext.kernel_launches() + wl.m_last.kernel_launches() + wl.m_local.kernel_launches()

    # ---- device-resident metric (matcher batches are asynchronous: the host prepares the next
    # submission while the GPU works; everything is ordered on one stream)
    for m in (wl.m_last, wl.m_local):
        m.set_async(True)
    wl.enable_side_streams(tstream)
    for i in range(Wm):
        wl.step_device(i)
    wl.finish_device()
    launches0 = launches_now()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        wl.step_device(i)
    e1.record()
    wl.finish_device()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    launches = launches_now() - launches0
    nm_last, nm_local = int(wl.nmatch_last.sum()), int(wl.nmatch_local.sum())
    # ---- end to end through the host-buffer ABI (wall clock brackets every copy and sync).
    # Batches come from independent camera streams, so `--e2e-workers` host threads (each with its
    # own extractor + matcher handles, i.e. its own CUDA streams) work on different batches at once:
    # one batch's PCIe copies and host staging overlap another batch's kernels.
    from concurrent.futures import ThreadPoolExecutor
    for m in (wl.m_last, wl.m_local):
        m.set_async(False)
    workers = [wl] + [Workload(B, POOL, rank, local_rank, None, share=wl) for _ in range(args.e2e_workers - 1)]
    for w in workers[1:]:
        for m in (w.m_last, w.m_local):
            m.set_stream(None)
    for m in (wl.m_last, wl.m_local):
        m.set_stream(None)
    pool_exec = ThreadPoolExecutor(len(workers))

    def run_host_steps(k):
        def work(widx):
            torch.cuda.set_device(local_rank)
            for i in range(widx, k, len(workers)):
                workers[widx].step_host(i)
        list(pool_exec.map(work, range(len(workers))))

    run_host_steps(max(Wm, len(workers)))
    barrier()
    t0 = time.perf_counter()
    run_host_steps(K)
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    for m in (wl.m_last, wl.m_local):
        m.set_stream(tstream.cuda_stream)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    nkp = int(wl.out_n.sum())

    # ---- per-kernel times for the roofline (events around each launch; separate pass)
    wl.ext.set_profiling(True)
    wl.ext.stage_times(reset=True)
    ms_match = [0.0, 0.0]
    for m in (wl.m_last, wl.m_local):
        m.set_async(False)
        m.set_stream(tstream.cuda_stream)
    wl.side = None
    for i in range(K):
        wl.step_device(i)
        ms_match[0] += wl.m_last.last_ms()
        ms_match[1] += wl.m_local.last_ms()
    wl.ext.synchronize()
    st = wl.ext.stage_times(reset=True)
    wl.ext.set_profiling(False)

    tens = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tens, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = tens.tolist()

    lba = None
    if not args.no_lba:
        lba = run_lba_gpu(rank, world, local_rank)
    stereo = None
    if rank == 0 and not args.no_stereo:
        stereo = _guarded(run_stereo_gpu, local_rank, tstream, cpu=(world == 1 and not args.no_cpu))
    pose = None
    if rank == 0 and not args.no_stereo:
        pose = _guarded(run_pose_gpu, local_rank, cpu=(world == 1 and not args.no_cpu))
    frustum = None
    if rank == 0 and not args.no_stereo:
        frustum = _guarded(run_frustum_gpu, local_rank, cpu=(world == 1 and not args.no_cpu))
    bow = None
    if rank == 0 and not args.no_stereo:
        bow = _guarded(run_bow_gpu, local_rank, cpu=(world == 1 and not args.no_cpu))
    if world > 1:
        dist.barrier()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        P = sum(level_pixels())
        n_mp = NFEAT + N_LOCAL_EXTRA
        # algorithmic bytes per frame and per kernel (DESIGN.md "Kernels")
        alg = {
            "pyramid": W * H + (P - W * H),           # read level 0, write levels 1..7
            "fast": P,                                # one read of the pyramid
            "blur": 2 * P,                            # read + write blurred pyramid
            "octree": 38000 * 8,                      # candidate records in
            "describe": NFEAT * (28 + 32) + NFEAT * (31 * 31 + 512),
        }
        kern = {k: v for k, v in st.items() if k in alg and v[1] > 0}

        def per_launch_ms(k):
            n_l = kern[k][1] if k != "pyramid" else kern[k][1] / (NLEVELS - 1)
            return kern[k][0] / n_l

        dom = max(kern, key=lambda k: kern[k][0])
        achieved = alg[dom] * B / (per_launch_ms(dom) * 1e-3) / 1e9
        total_ms = sum(v[0] for v in kern.values()) + sum(ms_match)
        traffic = None
        try:  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "fast_cells_traffic_r1.json")))
            if dom == "fast":
                traffic = tr["dram_bytes_per_launch"] * B / tr["batch"]
        except Exception:
            pass
        fps_dev = world * B * K / (ms_dev * 1e-3)
        fps_e2e = world * B * K / (ms_e2e * 1e-3)
        stage = {k: v[0] / K for k, v in st.items()}
        stage["match_last(th15)"] = ms_match[0] / K
        stage["match_local(th3)"] = ms_match[1] / K
        per_kernel = {}
        for k in kern:
            gbs = alg[k] * B / (per_launch_ms(k) * 1e-3) / 1e9
            per_kernel[k] = {"GB/s": gbs, "frac_of_hbm": gbs / hbm_peak}
        b_match = (NFEAT + n_mp) * 32 + NFEAT * 16 + n_mp * 24 + 64 * 48 * 4 + NFEAT * 4
        gbs = b_match * B / (max(ms_match[1], 1e-9) / K * 1e-3) / 1e9
        per_kernel["match_local"] = {"GB/s": gbs, "frac_of_hbm": gbs / hbm_peak}
        cpu = None
        if world == 1 and not args.no_cpu:
            from oracle import oracle as O
            O.build()
            threads = best_cpu_threads()
            fps_cpu, done, dt = cpu_extract_fps(wl.uniq[:8], threads, seconds_budget=8.0)
            fps_1, _, _ = cpu_extract_fps(wl.uniq[:8], 1, seconds_budget=3.0)
            t_match = cpu_match_seconds_per_frame(*make_frames(4, 1))
            t_ext = threads / fps_cpu
            cpu = {"value": threads / (t_ext + t_match), "unit": "frames/s", "cores": threads, "kind": "port",
                   "sample": "extract: %d frames over %d std::threads in %.1fs (%.1f frames/s; 1 thread %.1f frames/s); "
                             "match: %.2f ms/frame/thread (oracle C++ port)" % (done, threads, dt, fps_cpu, fps_1,
                                                                               1e3 * t_match)}
            if lba is not None:
                lba["cpu_baseline_config4"] = cpu_lba(50, 20000)
        h2d, d2h = wl.e2e_bytes()
        line = {
            "metric": METRIC, "value": fps_dev, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": B,
                       "l2": "inputs+pyramids %.0f MB per rotation > 126 MB L2 (%d batches rotated)"
                             % (POOL * B * (W * H + 2 * P) / 1e6, POOL),
                       "keypoints_last_step": nkp, "matches_last_step": {"last": nm_last, "local": nm_local}},
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "host_threads": args.e2e_workers},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "algorithmic_bytes_per_launch": alg[dom] * B,
                         "peak_source": peak_src,
                         "share_of_step": kern[dom][0] / total_ms, "stage_ms_per_step": stage,
                         "per_kernel": per_kernel},
            "cpu_baseline": cpu,
            "lba": lba,
            "stereo": stereo,
            "pose_optimization": pose,
            "is_in_frustum": frustum,
            "compute_bow": bow,
        }
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
