#!/usr/bin/env python3
"""profiles/r2b_summary.md: second half of round 2 -- bench lines, ncu launch lists and full captures copied into
profiles/ (scripts/final_measure.sh writes them under gpurun_out/ on the GPU box)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def load(name):
    return json.loads(open(os.path.join(P, name)).read().strip().splitlines()[-1])


def run(*a):
    return subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py")] + list(a),
                          capture_output=True, text=True).stdout


d, n2, ref = load("bench_r2b_n1.json"), load("bench_r2b_n2.json"), load("bench_r2b_reference_arm.json")
old = load("bench_r2_n1.json")
st, ost = d["roofline"]["stage_ms_per_step"], old["roofline"]["stage_ms_per_step"]
B = d["config"]["frames_per_step_per_gpu"]
c4, c5, o4, o5 = d["lba"]["config4"], d["lba"]["config5"], old["lba"]["config4"], old["lba"]["config5"]
o = []
o.append("# Round 2, second half -- measured on 1xB200 (gpurun box), synthetic data\n")
o.append("Same method as `r2_summary.md` (bench.py: CUDA events on the launching stream, 20 timed steps after 3 warm-ups, 256\n"
         "distinct frames per GPU; ncu captures of the same code under `gpurun`).  SM clock %.0f MHz, throttle reasons: %s.\n"
         % (d["clocks"]["sm_mhz"], d["clocks"]["reasons"] or "none"))
o.append("## Headline (`bench_r2b_n1.json`, reference arm `bench_r2b_reference_arm.json`, two GPUs `bench_r2b_n2.json`)\n")
o.append("| quantity | now | `r2_summary.md` |\n|---|---|---|")
o.append("| frames/s, inputs resident in HBM (`value`) | **%.0f** (%.2f ms per %d-frame step) | %.0f (%.2f ms) |"
         % (d["value"], d["ms_per_step"], B, old["value"], old["ms_per_step"]))
o.append("| frames/s end to end through the host-buffer C ABI (`e2e`) | **%.0f** | %.0f |" % (d["e2e"]["value"], old["e2e"]["value"]))
o.append("| CPU arm, same box (`--impl reference`, %s threads) | %.0f frames/s | |" % (ref["cpu_baseline"]["cores"], ref["value"]))
for name, c, oc in (("4 (50 KF x 20000 landmarks)", c4, o4), ("5 (200 KF x 80000 landmarks)", c5, o5)):
    o.append("| LocalBA config %s | **%.0f** LM iterations/s (%.2f ms for optimize(10)); through the C ABI incl. host structure build: %.0f (host prep %.1f ms) | %.0f (%.2f ms); C ABI %.0f (host prep %.1f ms) |"
             % (name, c["value"], c["ms_total"], c["e2e"]["value"], c["e2e"]["ms_host_prep"], oc["value"], oc["ms_total"],
                oc["e2e"]["value"], oc["e2e"]["ms_host_prep"]))
    o.append("| ... stages (ms per optimize(10)): linearize / Schur / reduced solve / update | %.2f / %.2f / %.2f / %.2f (%s) | %.2f / %.2f / %.2f / %.2f |"
             % (c["stage_ms"]["ms_linearize"], c["stage_ms"]["ms_schur"], c["stage_ms"]["ms_solve"], c["stage_ms"]["ms_update"],
                c["reduced_solver"], oc["stage_ms"]["ms_linearize"], oc["stage_ms"]["ms_schur"], oc["stage_ms"]["ms_solve"],
                oc["stage_ms"]["ms_update"]))
rig = d["lba"].get("config4_fisheye_rig")
if rig and "value" in rig:
    o.append("| LocalBA, %s (row a17) | **%.0f** LM iterations/s (%.2f ms); oracle port on one host thread: %.1f; parity vs oracle: same iteration / trial counts %s, relative error of the landmark / translation deltas %.1e / %.1e (bar 1e-4) | not built |"
             % (rig["config"], rig["value"], rig["ms_total"], rig["cpu_port"]["value"], rig["parity_vs_oracle"]["same_iterations_and_trials"],
                rig["parity_vs_oracle"]["rel_delta_points"], rig["parity_vs_oracle"]["rel_delta_translations"]))
n5 = n2["lba"]["config5"]
o.append("| 2xB200 | %.0f frames/s resident, %.0f e2e; config 5 sharded by landmark: %.0f LM iterations/s (%.2f ms), `sharded_equals_single` = %s (max dpose %.1e); configs[2] stereo streams: %.0f pairs/s | 72706 / 45258; 893 |"
         % (n2["value"], n2["e2e"]["value"], n5["value"], n5["ms_total"], n5["sharded_equals_single"],
            n5["max_abs_dpose_vs_single"], n2["stereo"]["value"]))
if os.path.exists(os.path.join(P, "bench_r2b_n4.json")):
    n4 = load("bench_r2b_n4.json")
    m5 = n4["lba"]["config5"]
    o.append("| 4xB200 (before the ranks were lined up ahead of the timed region, see DESIGN.md 5) | %.0f frames/s resident, %.0f e2e; config 5 sharded: %.0f LM iterations/s (%.2f ms; linearize / Schur / solve / update %.2f / %.2f / %.2f / %.2f), `sharded_equals_single` = %s | 160569 / 72044; 916 |"
             % (n4["value"], n4["e2e"]["value"], m5["value"], m5["ms_total"], m5["stage_ms"]["ms_linearize"], m5["stage_ms"]["ms_schur"],
                m5["stage_ms"]["ms_solve"], m5["stage_ms"]["ms_update"], m5["sharded_equals_single"]))
for nn, f in ((2, "bench_r2c_n2.json"), (4, "bench_r2c_n4.json")):
    if os.path.exists(os.path.join(P, f)):
        x = load(f)
        m5 = x["lba"]["config5"]
        o.append("| %dxB200, ranks lined up before the timed region | %.0f frames/s resident, %.0f e2e; config 5 sharded: **%.0f** LM iterations/s (%.2f ms; linearize / Schur / solve / update %.2f / %.2f / %.2f / %.2f), `sharded_equals_single` = %s; through the C ABI %.0f | |"
                 % (nn, x["value"], x["e2e"]["value"], m5["value"], m5["ms_total"], m5["stage_ms"]["ms_linearize"], m5["stage_ms"]["ms_schur"],
                    m5["stage_ms"]["ms_solve"], m5["stage_ms"]["ms_update"], m5["sharded_equals_single"], m5["e2e"]["value"]))
o.append("")
o.append("## Where a step goes (CUDA events per stage in a serial profiling pass, ms per %d-frame step)\n" % B)
o.append("| stage | ms/step | us/frame | `r2_summary.md` us/frame |\n|---|---|---|---|")
for k in st:
    if k in ("h2d", "d2h"):
        continue
    o.append("| %s | %.3f | %.2f | %.2f |" % (k, st[k], st[k] * 1e3 / B, ost.get(k, 0) * 1e3 / B))
r = d["roofline"]
o.append("\n`roofline` of the bench line: dominant kernel = `%s` (%.0f %% of the step), %.0f GB/s of algorithmic traffic = **%.1f %% of the measured HBM peak** (%.0f GB/s).\n"
         % (r["kernel"], 100 * r["share_of_step"], r["achieved"], 100 * r["frac"], r["peak"]))
for title, f in (("ncu launch list of one bench run without the LBA / stereo legs (`launches_r2b.csv`; cold-cache, serialised: compare shares)", "launches_r2b.csv"),
                 ("ncu launch list of three LocalBA config-5 solves (`launches_lba_r2b.csv`)", "launches_lba_r2b.csv")):
    if os.path.exists(os.path.join(P, f)):
        o.append("## " + title + "\n")
        o.append(run("launches", os.path.join(P, f)))
for title, f in (("ncu, full captures of the LocalBA kernels (`lba_r2b.ncu-rep`)", "lba_r2b.ncu-rep"),
                 ("ncu, full captures of `resize_words_kernel` / `fast_warp_kernel` (`extract_r2b.ncu-rep`)", "extract_r2b.ncu-rep")):
    if os.path.exists(os.path.join(P, f)):
        o.append("## " + title + "\n")
        o.append(run("raw", os.path.join(P, f)))
notes = os.path.join(P, "r2b_notes.md")
if os.path.exists(notes):
    o.append(open(notes).read())
open(os.path.join(P, "r2b_summary.md"), "w").write("\n".join(o) + "\n")
print("wrote profiles/r2b_summary.md")
