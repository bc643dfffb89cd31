"""Regenerates profiles/r1_scan_strip_ncu.{md,json} and profiles/r1_launch_summary.md from the scratch captures in gpurun_out/
(scan_r1_aw_final.ncu-rep, r1_launches.csv) and profiles/r1_bench.json.  Run in the build container after a gpurun capture."""
import collections
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(ROOT, "gpurun_out", "scan_r1_aw_final.ncu-rep")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
g = lambda k: (vals[hdr.index(k)], units[hdr.index(k)]) if k in hdr else ("n/a", "")
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
tab = "\n".join(f"| `{k}` | {g(k)[0]} {g(k)[1]} |" for k in want)


def num(k):
    v, u = g(k)
    v = float(v.replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1, "us": 1e-3, "msecond": 1, "usecond": 1e-3}.get(u, 1)


rd, wr, ms = num("dram__bytes_read.sum"), num("dram__bytes_write.sum"), num("gpu__time_duration.sum")
b = json.load(open(os.path.join(ROOT, "profiles", "r1_bench.json")))
alg = b["roofline"]["algorithmic_bytes_per_launch"]
cmd = "ncu --set full --clock-control none --import-source on -k regex:hevcb_scan_strip_kernel -s 3 -c 1 python bench.py --steps 1 --warmup 3 --no-sweep --no-parse --no-rewrite --no-insert"
md = f"""# ncu --set full, hevcb_scan_strip_kernel, round 1 (analyser / writer design)

`{cmd}` on one B200 (4 GiB buffer, 16 KiB NALs, 258 048 NALs).
Report file: `gpurun_out/scan_r1_aw_final.ncu-rep` (scratch; numbers copied here by `tools/make_profile_summaries.py`).  Times under ncu are cold-cache and serialised; the bench value (CUDA events, no profiler) is in `profiles/r1_bench.json`.

| metric | value |
|---|---|
{tab}

DRAM traffic per launch = read + write = {rd / 1e9:.3f} + {wr / 1e9:.3f} = **{(rd + wr) / 1e9:.3f} GB**; algorithmic bytes (N_in + N_rbsp + 24*NALs) = {alg / 1e9:.3f} GB -> traffic / algorithmic = {(rd + wr) / alg:.3f}.
The image is written once.  The input is loaded twice (once by an analyser CTA, once by a writer CTA) but the second load is served by L2: the analysers stay within a 1100-tile (34 MiB) window of the writers' progress counter.  Measured sensitivity (same command, `HEVCB_SCAN_WINDOW`, 2 stages per CTA): 1200 tiles -> DRAM reads 1.27x the input, 1100 -> 1.11x, both at the same speed within 1 % (the kernel is not DRAM-bound); at 1050 and below the pipeline (2 tiles in flight per CTA in both roles = 592 tiles, plus one scanner batch of 320) is throttled.

Where the time goes now (warp-state sampling of this capture, `--page source`): 31 % of all samples are the analysers' worker warps waiting for their bulk copies (`mbarrier.try_wait` loop), 15 % the writers' workers at the \"prefix ready\" barrier, 2 % the analysers' control warps polling the writers' progress (the L2 window); no single arithmetic instruction holds more than 1.5 %.  Issue slots are ~37 % used: the pass is bound by the latency of the memory system under the mixed load (the bare load pipeline alone sustains 4.0 TB/s, see DESIGN 4.1), not by DRAM bandwidth or instruction issue.  The previous single-role pipeline spent 38 % of all warp samples at one CTA barrier in lock-step with the scanner (`profiles/r1_scan_strip_ncu_v1.md`).

SASS evidence of the async-copy path: `UBLKCP.S.G` (cp.async.bulk), `SYNCS.ARRIVE.TRANS64` / `SYNCS.PHASECHK.TRANS64.TRYWAIT` (mbarrier), `REDUX`, `VOTE`, `LDG.E.128.STRONG.GPU` (tile-state polls) in `cuobjdump -sass hevcbitstream_b200/libhevcb200.so`.
"""
open(os.path.join(ROOT, "profiles", "r1_scan_strip_ncu.md"), "w").write(md)
json.dump({"kernel": "hevcb_scan_strip_kernel", "command": cmd, "workload": "nal16k", "size_gib": 4.0, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "traffic_bytes_per_launch": rd + wr, "duration_ms_under_ncu": ms}, open(os.path.join(ROOT, "profiles", "r1_scan_strip_ncu.json"), "w"), indent=1)
print("scan:", rd, wr, ms, (rd + wr) / alg)

rows = [r for r in csv.reader(open(os.path.join(ROOT, "profiles", "r1_launches.csv"))) if r]
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
c = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    short = r[c["Kernel Name"]].split("(")[0].split("::")[-1][:70]
    v = float(r[c["Metric Value"]].replace(",", ""))
    u = r[c["Metric Unit"]]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u in ("ms", "msecond") else v)
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
lines = ["# ncu launch list summary, round 1 (analyser / writer scan kernel)", "",
         "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv python bench.py --steps 2 --warmup 3 --no-sweep --e2e-gib 0.25` (B200, cold-cache serialised launch times: compare SHARES, not absolutes).",
         "Raw CSV: `profiles/r1_launches.csv` (warm-up + timed steps of the headline workload, then the e2e, parse, insert and rewrite sections of the bench).", "",
         "| kernel | launches | total us | share |", "|---|---|---|---|"]
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| `{k}` | {n} | {us:.1f} | {100 * us / tot:.1f}% |")
per = lambda name: agg.get(name, [1, 0.0])[1] / max(agg.get(name, [1, 0.0])[0], 1)
step = sum(per(k) for k in ("hevcb_scan_strip_kernel", "hevcb_scan_emit_kernel", "hevcb_scan_finalize_kernel", "hevcb_scan_init_kernel"))
lines += ["", f"Per scan step the launches are: memset (tile states), `hevcb_scan_init_kernel` ({per('hevcb_scan_init_kernel'):.1f} us), `hevcb_scan_strip_kernel` "
          f"({per('hevcb_scan_strip_kernel'):.1f} us on average over the differently sized scans of the run, cooperative), `hevcb_scan_emit_kernel` ({per('hevcb_scan_emit_kernel'):.1f} us), "
          f"`hevcb_scan_finalize_kernel` ({per('hevcb_scan_finalize_kernel'):.1f} us): the strip kernel is {100 * per('hevcb_scan_strip_kernel') / step:.1f}% of a step's kernel time, "
          "which is the share `bench.py` attributes the roofline to (its CUDA-event time covers all of them)."]
open(os.path.join(ROOT, "profiles", "r1_launch_summary.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[7:14]))
