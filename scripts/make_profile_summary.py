#!/usr/bin/env python3
"""profiles/r1_summary.md from the bench lines and ncu captures copied into profiles/."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def run(*a):
    return subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py")] + list(a),
                          capture_output=True, text=True).stdout


d = json.load(open(os.path.join(P, "bench_r1_n1.json")))
r = json.load(open(os.path.join(P, "bench_r1_reference_arm.json")))
n2 = json.load(open(os.path.join(P, "bench_r1_n2.json")))
extract_raw = run("raw", os.path.join(P, "extract_r1.ncu-rep"))
lba_raw = run("raw", os.path.join(P, "lba_r1.ncu-rep"))
launch = run("launches", os.path.join(P, "launches_r1.csv"))
sec = extract_raw[extract_raw.index("### fast_cells_kernel"):]
mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = re.search(r"dram__bytes_read.sum = ([\d.]+) (\w+)", sec)
wr = re.search(r"dram__bytes_write.sum = ([\d.]+) (\w+)", sec)
traffic = float(rd.group(1)) * mul[rd.group(2)] + float(wr.group(1)) * mul[wr.group(2)]
json.dump({"kernel": "fast_cells_kernel", "batch": 64, "dram_bytes_per_launch": traffic,
           "source": "profiles/extract_r1.ncu-rep (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"},
          open(os.path.join(P, "fast_cells_traffic_r1.json"), "w"), indent=1)
st = d["roofline"]["stage_ms_per_step"]
B = d["config"]["frames_per_step_per_gpu"]
pk = d["roofline"]["per_kernel"]
o = []
o.append("# Round 1 -- measured on 1xB200 (gpurun box, 128 logical host CPUs), synthetic data\n")
o.append("Numbers come from `bench.py` (CUDA events on the launching stream, 20 timed steps after 3 warm-ups, three\n"
         "batches rotated = 2.5 GB of inputs + pyramids per rotation >> 126 MB L2) and from ncu captures of the same\n"
         "code taken under `gpurun` (files in this directory).  SM clock 1965 MHz throughout, no throttle reason active.\n")
o.append("## Headline (`bench_r1_n1.json`, reference arm `bench_r1_reference_arm.json`)\n")
o.append("| quantity | value |\n|---|---|")
o.append(f"| frames/s, inputs resident in HBM (`value`) | **{d['value']:.0f}** ({d['ms_per_step']:.2f} ms per {B}-frame step, {d['gpu_launches']} kernel launches / 20 steps) |")
o.append(f"| frames/s end to end through the host-buffer C ABI (`e2e`, {d['e2e']['host_threads']} host threads, {d['e2e']['h2d_bytes_per_step']/1e6:.0f} MB H2D + {d['e2e']['d2h_bytes_per_step']/1e6:.1f} MB D2H per step) | **{d['e2e']['value']:.0f}** |")
o.append(f"| CPU port of the reference path, same box (`--impl reference`) | {r['value']:.0f} frames/s on {r['cpu_baseline']['cores']} std::threads (best of 4..128; 1 thread = 25 frames/s) |")
c4, c5 = d["lba"]["config4"], d["lba"]["config5"]
o.append(f"| LocalBA config 4 ({c4['config']}) | **{c4['value']:.0f}** LM iterations/s ({c4['ms_total']:.1f} ms for optimize(10), {c4['trials']} trials) vs {d['lba']['cpu_baseline_config4']['value']:.1f} on one CPU thread (g2o is single-threaded) |")
o.append(f"| LocalBA config 5 ({c5['config']}), 1 GPU | {c5['value']:.0f} LM iterations/s ({c5['ms_total']:.1f} ms) |")
o.append(f"| 2xB200 (`bench_r1_n2.json`) | {n2['value']:.0f} frames/s resident; config 5 sharded by landmark + ncclAllReduce: {n2['lba']['config5']['value']:.0f} LM iterations/s |")
if d.get("stereo") and "error" not in d["stereo"]:
    sx = d["stereo"]
    o.append(f"| SURVEY 8(f-1) `Frame::ComputeStereoMatches`, {sx['config']} | **{sx['stereo_match_us_per_pair']:.1f} us per pair** on the device-resident extractor outputs (3 launches per 64 pairs, {sx['matches_per_pair']:.0f} matches/pair) vs {sx['cpu_baseline']['stereo_match_ms_per_pair']:.2f} ms on one CPU thread; whole stereo frame (2 x extract + match): {sx['pairs_per_s']:.0f} pairs/s; bench-time parity vs oracle: {sx['cpu_baseline']['parity_pair0']} |")
if d.get("pose_optimization") and "error" not in d["pose_optimization"]:
    px = d["pose_optimization"]
    pc = px["cpu_baseline"]
    o.append(f"| SURVEY 8(f-2) `Optimizer::PoseOptimization`, {px['config']} | **{px['value']:.0f} calls/s** in one kernel launch ({px['ms_per_step']:.2f} ms per 128 frames, {px['lm_trials_per_frame']:.0f} LM trials per frame); {px['e2e_value']:.0f} calls/s through the host-buffer ABI; CPU port {pc['ms_per_call']:.2f} ms per call on one thread; bench-time parity vs oracle: {[v for k, v in pc.items() if k.startswith('parity')][0]} |")
if not d.get("is_in_frustum") and os.path.exists(os.path.join(P, "bench_r1_f3_leg.json")):
    # the 8(f-3) leg was added after the 20-step line was taken: its numbers come from a later short run
    d["is_in_frustum"] = json.load(open(os.path.join(P, "bench_r1_f3_leg.json"))).get("is_in_frustum")
if d.get("is_in_frustum") and "error" not in d["is_in_frustum"]:
    fl, fs = d["is_in_frustum"]["local_map_5000"], d["is_in_frustum"]["stream_1M"]
    o.append(f"| SURVEY 8(f-3) `Frame::isInFrustum` over a local map | 5000 points: kernel {fl['kernel_us']:.0f} us, host-buffer call {fl['host_call_us']:.0f} us (CPU port {fl.get('cpu_port_us', float('nan')):.0f} us); 2^20 points: kernel {fs['kernel_us']:.0f} us = **{fs['kernel_GBps_algorithmic']:.0f} GB/s** algorithmic (57 B/point), CPU port {fs.get('cpu_port_us', float('nan'))/1e3:.1f} ms; bench-time parity vs oracle: {fl.get('parity')} / {fs.get('parity')} |")
if d.get("compute_bow") and "error" not in d["compute_bow"]:
    bw = d["compute_bow"]
    o.append(f"| SURVEY 8(f-4) `Frame::ComputeBoW` (DBoW2 transform), {bw['config']} | kernels {bw['kernels_us']:.0f} us, call with descriptors already on the device {bw['call_us_descriptors_on_device']:.0f} us, CPU port {bw.get('cpu_port_us', float('nan')):.0f} us; bench-time parity vs oracle: {bw.get('parity')} |")
o.append(f"\n## Where a step goes (CUDA events per stage, ms per {B}-frame step)\n")
o.append("| stage | ms/step | us/frame | algorithmic GB/s | frac of HBM peak (6556 GB/s measured) |\n|---|---|---|---|---|")
for k in ["pyramid", "fast", "octree", "blur", "describe", "layout", "match_last(th15)", "match_local(th3)"]:
    key = k if k in pk else ("match_local" if k.startswith("match_local") else None)
    g = pk[key]["GB/s"] if key in pk else None
    o.append(f"| {k} | {st[k]:.3f} | {st[k]*1000/B:.1f} | {'' if g is None else '%.0f' % g} | {'' if g is None else '%.3f' % pk[key]['frac_of_hbm']} |")
o.append("\n(per-stage events are taken with everything on one stream; in the timed `value` run the blur and the two\n"
         "matchers run on side streams, which is why the stages add up to more than `ms_per_step`.)\n")
rf = d["roofline"]
o.append(f"`roofline` of the bench line: dominant kernel = `{rf['kernel']}` ({rf['share_of_step']*100:.0f} % of the step), "
         f"{rf['achieved']:.0f} GB/s of algorithmic traffic = **{rf['frac']*100:.1f} % of the measured HBM peak**; DRAM traffic "
         f"{traffic/1e6:.0f} MB per 64-frame launch (ncu capture at batch 64) vs {rf['algorithmic_bytes_per_launch']/1e6*64/B:.0f} MB algorithmic: no re-reads. "
         "ncu explains the gap to the HBM roofline: the kernel is instruction-issue bound (70 % of issue slots busy, <3 % DRAM throughput).\n")
o.append("## ncu launch list (`launches_r1.csv`, `--metrics gpu__time_duration.sum --clock-control none`, one `bench.py --steps 2` run)\n")
o.append(launch)
tot = sum(st[k] for k in ["pyramid", "fast", "octree", "blur", "describe", "layout", "match_last(th15)", "match_local(th3)"])
o.append("\nEvent-based shares of the same stages for comparison (the launch list is cold-cache and serialises the side "
         "streams, so the shares must agree, not the absolute times): "
         + ", ".join(f"{k} {100 * st[k] / tot:.0f} %" for k in ["fast", "octree", "blur", "pyramid", "describe"]) + ".\n")
o.append("## ncu `--set full`, extractor kernels at batch 64 (`extract_r1.ncu-rep`)\n")
o.append(extract_raw)
o.append("\n## ncu `--set full`, LBA kernels on config 4 (`lba_r1.ncu-rep`)\n")
o.append(lba_raw)
if os.path.exists(os.path.join(P, "f_rows_r1.ncu-rep")):
    o.append("\n## ncu `--set full`, SURVEY 8(f) kernels (`f_rows_r1.ncu-rep`: ComputeStereoMatches on 64 pairs, "
             "PoseOptimization on 128 frames x 1000 edges; `scripts/profile_f.py`)\n")
    o.append(run("raw", os.path.join(P, "f_rows_r1.ncu-rep")))
o.append("""
## What moved during the round (us per 720p frame at batch 64, CUDA events)

| kernel | first correct version | end of round | what did it |
|---|---|---|---|
| FAST cells | 36 | 10.4 | compaction queue instead of divergent heavy path; iniTh pass first and minTh pass only for empty cells (the reference's own two cv::FAST calls); packed-byte VABSDIFF4 pre-test; tile through TMA (UTMALDG.3D + mbarrier) |
| describe (IC angle + rBRIEF) | 15 | 4.7 | **pattern table moved from `__constant__` (lane-divergent index = 32 replays per load) to a lane-transposed shared-memory copy** (ncu source view: 45 % of samples on `LDC.64`) |
| blur | 11.5 | 3.2 | 4 pixels / thread, aligned word loads and stores (6.3); then the taps as byte vectors on the integer dot-product unit: IDP.4A horizontal, IDP.2A vertical on row-pair words, PRMT packing (68 -> ~16 lane instructions per pixel) |
| octree | 8.5 | 6.5 | node arrays in shared memory, 4-way batched point loops, level-major launch order |
| level-0 copy | 4.1 | 0.5 | one vectorised kernel instead of 64 cudaMemcpy2DAsync |
| matchers (both) | 12.2 | 10.5 (overlapped on two side streams) | per-keypoint state of the resolution rounds in shared memory |
| LBA Schur pairs (config 4, per trial) | 241 | 90 | 4 index/operand loads in flight per warp, two DMMA accumulators, 8 warps |
| LBA solve (config 4, per trial) | 590 | 290 | warp-per-row panel solve, left-looking diagonal block, shared-memory back-substitution |

Tried and reverted (measured slower or no gain): L2-sized sub-batches (no stage is HBM bound), orient/brief kernel
split, forcing 64 registers on describe, fewer CTAs / 4-accumulator ILP in the LDLT diagonal block, staging the
describe patch in shared memory before the constant-memory fix, running two to four sub-batches on concurrent
(also prioritised) stream lanes to hide the latency-bound octree under the other lanes' FAST grids (27.5 k -> 27.7 k fps
plain, 25.4 k with priorities).
""")
open(os.path.join(P, "r1_summary.md"), "w").write("\n".join(o))
print("written", len(o))
