#!/usr/bin/env python3
"""profiles/r2_summary.md from the round-2 bench lines and ncu captures copied into profiles/."""
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


def load(name):
    p = os.path.join(P, name)
    return json.load(open(p)) if os.path.exists(p) else None


def metric(text, kernel, key):
    sec = text[text.index("### " + kernel):]
    nxt = sec.find("###", 4)
    sec = sec if nxt < 0 else sec[:nxt]
    m = re.search(re.escape(key) + r" = ([\d.,]+) ?(\w*)", sec)
    if not m:
        return None
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m.group(2), 1)
    return float(m.group(1).replace(",", "")) * mul


d, ref, n2, n4 = load("bench_r2_n1.json"), load("bench_r2_reference_arm.json"), load("bench_r2_n2.json"), load("bench_r2_n4.json")
r1 = load("bench_r1_n1.json")
ext = run("raw", os.path.join(P, "extract_r2.ncu-rep"))
lba = run("raw", os.path.join(P, "lba_r2.ncu-rep"))
octr = run("raw", os.path.join(P, "octree_r2.ncu-rep"))
rsz = run("raw", os.path.join(P, "resize_r2.ncu-rep"))
launch = run("launches", os.path.join(P, "launches_r2.csv"))
launch_lba = run("launches", os.path.join(P, "launches_lba_r2.csv"))

# DRAM traffic of the dominant kernel (the capture ran 3 lanes of a 64-frame batch: 22 frames per launch)
FRAMES_PER_CAPTURED_LAUNCH = 22
tr = metric(ext, "void fast_warp_kernel<64>", "dram__bytes_read.sum") + metric(ext, "void fast_warp_kernel<64>", "dram__bytes_write.sum")
json.dump({"kernel": "fast_warp_kernel<64>", "batch": FRAMES_PER_CAPTURED_LAUNCH, "dram_bytes_per_launch": tr,
           "source": "profiles/extract_r2.ncu-rep (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum; "
                     "64-frame batch in 3 lanes = 22 frames per captured launch)"},
          open(os.path.join(P, "fast_traffic_r2.json"), "w"), indent=1)
tp = metric(lba, "schur_pairs_kernel", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
json.dump({"kernel": "schur_pairs_kernel", "config": "config 5 (200 KF x 80000 landmarks)",
           "sm__pipe_tensor_cycles_active_pct": tp, "source": "profiles/lba_r2.ncu-rep (ncu --set full)"},
          open(os.path.join(P, "schur_pairs_r2.json"), "w"), indent=1)

st = d["roofline"]["stage_ms_per_step"]
B = d["config"]["frames_per_step_per_gpu"]
pk = d["roofline"]["per_kernel"]
o = []
o.append("# Round 2 -- measured on 1xB200 (gpurun box, 128 logical host CPUs, 2 NUMA nodes), synthetic data\n")
o.append("Numbers come from `bench.py` (CUDA events on the launching stream, 20 timed steps after 3 warm-ups, 256 distinct\n"
         "frames per GPU in 8 streams, two 128-frame batches rotated) and from ncu captures of the same code taken under\n"
         "`gpurun` (files in this directory).  SM clock %s MHz throughout, throttle reasons: %s.\n"
         % (d["clocks"]["sm_mhz"], d["clocks"]["reasons"] or "none"))
o.append("## Headline (`bench_r2_n1.json`, reference arm `bench_r2_reference_arm.json`)\n")
o.append("| quantity | round 2 | round 1 |\n|---|---|---|")
o.append(f"| frames/s, inputs resident in HBM (`value`) | **{d['value']:.0f}** ({d['ms_per_step']:.2f} ms per {B}-frame step) | {r1['value']:.0f} ({r1['ms_per_step']:.2f} ms) |")
o.append(f"| frames/s end to end through the host-buffer C ABI (`e2e`, {d['e2e']['host_threads']} host threads, {d['e2e']['h2d_bytes_per_step']/1e6:.0f} MB H2D + {d['e2e']['d2h_bytes_per_step']/1e6:.1f} MB D2H per step) | **{d['e2e']['value']:.0f}** | {r1['e2e']['value']:.0f} ({r1['e2e']['h2d_bytes_per_step']/1e6:.0f} MB H2D) |")
o.append(f"| CPU arm, same box (`--impl reference`): the reference's own `ORBextractor.cc` object code (`oracle/_ref`) on {ref['cpu_baseline']['cores']} std::threads + oracle matchers | {ref['value']:.0f} frames/s (kind \"{ref['cpu_baseline']['kind']}\") | 364-390 (oracle port) |")
c4, c5 = d["lba"]["config4"], d["lba"]["config5"]
r4, r5 = r1["lba"]["config4"], r1["lba"]["config5"]
o.append(f"| LocalBA config 4 ({c4['config']}) | **{c4['value']:.0f}** LM iterations/s ({c4['ms_total']:.2f} ms for optimize(10)); through the C ABI incl. host structure build: {c4['e2e']['value']:.0f} | {r4['value']:.0f} ({r4['ms_total']:.2f} ms) |")
o.append(f"| LocalBA config 5 ({c5['config']}) | **{c5['value']:.0f}** LM iterations/s ({c5['ms_total']:.2f} ms); C ABI: {c5['e2e']['value']:.0f} (host prep {c5['e2e']['ms_host_prep']:.1f} ms on 8 threads) | {r5['value']:.0f} ({r5['ms_total']:.2f} ms) |")
o.append(f"| reduced solve (`ms_solve` per optimize(10)), config 4 / 5 | {c4['stage_ms']['ms_solve']:.2f} / {c5['stage_ms']['ms_solve']:.2f} ms ({c5['reduced_solver']}) | {r4['stage_ms']['ms_solve']:.2f} / {r5['stage_ms']['ms_solve']:.2f} ms (dense cooperative) |")
o.append(f"| Schur roofline (config 5) | {c5['roofline']['achieved']:.2f} TFLOP/s useful of {c5['roofline']['peak']:.1f} TFLOP/s measured fp64 DMMA peak = **{100*c5['roofline']['frac']:.1f} %**; tensor pipe active {tp:.1f} % (ncu) | 714 GF/s, no denominator |")
for name, x in (("2", n2), ("4", n4)):
    if x:
        c = x["lba"]["config5"]
        o.append(f"| {name}xB200 | {x['value']:.0f} frames/s resident, {x['e2e']['value']:.0f} e2e; config 5 sharded by landmark: {c['value']:.0f} LM iterations/s, all-reduce {c['allreduce_bytes_per_trial']/1e6:.2f} MB per trial, `sharded_equals_single` = {c.get('sharded_equals_single')} (max dpose {c.get('max_abs_dpose_vs_single'):.1e}); configs[2] stereo streams: {x['stereo']['value']:.0f} pairs/s | |")
sx = d["stereo"]
o.append(f"| configs[2] leg on one GPU ({sx['config']}) | {sx['value']:.0f} stereo pairs/s incl. SearchForTriangulation ({sx['triangulate_us_per_kf_pair']:.0f} us per keyframe pair through host buffers); ComputeStereoMatches {sx['stereo_match_us_per_pair']:.1f} us/pair; parity vs oracle in the bench: {sx['cpu_baseline']['parity_pair0']} / {sx['cpu_baseline']['parity_triangulation0']} | 15423 pairs/s without triangulation |")
lat = d["latency_batch1"]
o.append("| batch-1 latency through the host-buffer ABI (us, GPU / one CPU thread) | " + "; ".join(
    f"{k}: {v['gpu_us']:.0f} / {v['cpu_port_us']:.0f}" for k, v in lat.items() if isinstance(v, dict)) + " | not measured |")
fl = d["is_in_frustum"]["local_map_5000"]
o.append(f"| `isInFrustum`, 5000-point local map, host-buffer call | {fl['host_call_us']:.0f} us (kernel {fl['kernel_us']:.1f} us; CPU port {fl['cpu_port_us']:.0f} us) | 112 us (CPU 87 us) |")
o.append(f"\n## Where a step goes (CUDA events per stage in a serial profiling pass, ms per {B}-frame step)\n")
o.append("| stage | ms/step | us/frame | round 1 us/frame | algorithmic GB/s | frac of HBM peak (6556 GB/s measured) |\n|---|---|---|---|---|---|")
r1st = r1["roofline"]["stage_ms_per_step"]
for k in ["pyramid", "fast", "octree", "blur", "describe", "layout", "match_last(th15)", "match_local(th3)"]:
    key = k if k in pk else ("match_local" if k.startswith("match_local") else None)
    g = pk[key]["GB/s"] if key in pk else None
    o.append(f"| {k} | {st[k]:.3f} | {st[k]*1000/B:.2f} | {r1st[k]*1000/128:.2f} | {'' if g is None else '%.0f' % g} | {'' if g is None else '%.3f' % pk[key]['frac_of_hbm']} |")
o.append("\nIn the timed `value` run the blur runs on a side stream, the 128-frame batch runs as 3 lanes, and the two matchers\n"
         "run on their own streams concurrently with the NEXT step's extraction (double-buffered extractor results), which is\n"
         "why `ms_per_step` is below the sum of the stages.\n")
rf = d["roofline"]
o.append(f"`roofline` of the bench line: dominant kernel = `{rf['kernel']}` ({rf['share_of_step']*100:.0f} % of the step), "
         f"{rf['achieved']:.0f} GB/s of algorithmic traffic = **{rf['frac']*100:.1f} % of the measured HBM peak** (round 1: 4.2 %); DRAM traffic "
         f"{tr/1e6/FRAMES_PER_CAPTURED_LAUNCH:.2f} MB per frame (ncu) vs 2.85 MB algorithmic: no re-reads.\n")
o.append("## ncu, full captures (`extract_r2.ncu-rep`: fast_warp_kernel + describe_tma_kernel; `octree_r2.ncu-rep`; `resize_r2.ncu-rep`; `lba_r2.ncu-rep`)\n")
o.append(ext)
o.append(octr)
o.append(rsz)
o.append(lba)
o.append("## ncu launch list of one bench run (`launches_r2.csv`, `--metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)\n")
o.append(launch)
o.append("## ncu launch list of one LocalBA config-5 solve (`launches_lba_r2.csv`)\n")
o.append(launch_lba)
extra = os.path.join(P, "r2_notes.md")
if os.path.exists(extra):
    o.append(open(extra).read())
open(os.path.join(P, "r2_summary.md"), "w").write("\n".join(o) + "\n")
print("wrote profiles/r2_summary.md")
