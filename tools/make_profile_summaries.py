"""Regenerates the round-2 profile summaries under profiles/ from the scratch captures in gpurun_out/ (ncu --set full reports of
hevcb_scan_strip_kernel on the nal16k / nal64 / epb_dense_4k bench workloads, the ncu launch list of a bench run) and
profiles/r2_bench.json.  Run in the build container after the gpurun captures (commands are quoted in the outputs)."""
import collections
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1, "us": 1e-3, "msecond": 1, "usecond": 1e-3}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    g = lambda k: (vals[hdr.index(k)], units[hdr.index(k)]) if k in hdr else ("n/a", "")
    num = lambda k: float(g(k)[0].replace(",", "")) * UNIT.get(g(k)[1], 1)
    return g, num


def top_lines(rep, n=8):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    res, hdr, fil = [], None, ""
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fil = r[1]
        elif len(r) > 6 and r[0] == "Line No":
            hdr = r
            si = r.index("# Samples")
        elif hdr and len(r) > si and r[0].isdigit():
            try:
                res.append((int(r[si]), os.path.basename(fil), int(r[0]), r[1].strip()[:90]))
            except ValueError:
                pass
    tot = sum(x[0] for x in res) or 1
    return tot, sorted(res, reverse=True)[:n]


bench = json.load(open(os.path.join(ROOT, "profiles", "r2_bench.json")))
alg16 = bench["roofline"]["algorithmic_bytes_per_launch"]
sections = []
for wl, title in (("nal16k", "headline: 4 GiB, 16 KiB NALs"), ("nal64", "4 GiB, 64-byte NALs (every chunk needs the exact masks)"),
                  ("dense", "4 GiB, EPB-dense payload 00 00 03 01 (a quarter of the bytes removed)")):
    rep = os.path.join(ROOT, "gpurun_out", f"r2_scan_{wl}.ncu-rep")
    if not os.path.exists(rep):
        continue
    g, num = raw(rep)
    rd, wr, ms = num("dram__bytes_read.sum"), num("dram__bytes_write.sum"), num("gpu__time_duration.sum")
    tab = "\n".join(f"| `{k}` | {g(k)[0]} {g(k)[1]} |" for k in WANT)
    tot, tops = top_lines(rep)
    tl = "\n".join(f"| {100 * c / tot:.1f} % | `{f}:{ln}` | `{src}` |" for c, f, ln, src in tops)
    wname = {"nal16k": "nal16k", "nal64": "nal64", "dense": "epb_dense_4k"}[wl]
    cmd = (f"ncu --set full --clock-control none --import-source on -k regex:hevcb_scan_strip_kernel -s 3 -c 1 python bench.py --workload {wname} --steps 1 "
           "--warmup 3 --no-sweep --no-parse --no-rewrite --no-insert --no-cpu-extras")
    sec = f"""## {title}

`{cmd}` (report: `gpurun_out/r2_scan_{wl}.ncu-rep`, scratch).

| metric | value |
|---|---|
{tab}

DRAM traffic per launch: read {rd / 1e9:.3f} GB + write {wr / 1e9:.3f} GB = **{(rd + wr) / 1e9:.3f} GB**.
"""
    if wl == "nal16k":
        sec += (f"Algorithmic bytes (N_in + N_rbsp + 24 x NALs) = {alg16 / 1e9:.3f} GB -> traffic / algorithmic = **{(rd + wr) / alg16:.3f}**: the input is read from "
                "HBM once (the tile stays in shared memory until its prefix has arrived; the L2 prefetch pulls every byte exactly once), the image is written once.\n")
        json.dump({"kernel": "hevcb_scan_strip_kernel", "command": cmd, "workload": "nal16k", "size_gib": 4.0, "dram_bytes_read": rd, "dram_bytes_write": wr,
                   "traffic_bytes_per_launch": rd + wr, "duration_ms_under_ncu": ms}, open(os.path.join(ROOT, "profiles", "r2_scan_strip_ncu.json"), "w"), indent=1)
    sec += f"\nWarp-state samples by source line (top {len(tops)} of {tot} samples):\n\n| share | line | source |\n|---|---|---|\n{tl}\n"
    sections.append(sec)
sw = bench.get("sweep", {})
swt = "\n".join(f"| {k} | {v['input_GBps']} | {v['algorithmic_GBps']} | {v['frac_of_peak']} |" for k, v in sw.items())
md = f"""# ncu --set full, hevcb_scan_strip_kernel, round 2 (single-pass ring pipeline)

Times under ncu are cold-cache and serialised; the bench values (CUDA events, no profiler) are in `profiles/r2_bench.json`:
headline {bench['value']:.0f} GB/s of input = {bench['roofline']['achieved']:.0f} GB/s algorithmic = **{bench['roofline']['frac']:.3f}** of the measured HBM peak ({bench['roofline']['peak']} GB/s).

| workload | input GB/s | algorithmic GB/s | fraction of measured peak |
|---|---|---|---|
{swt}

""" + "\n".join(sections) + """
SASS evidence of the async-copy path (`cuobjdump -sass hevcbitstream_b200/libhevcb200.so`): `UBLKCP.S.G` (cp.async.bulk global -> shared), `UBLKPF` (cp.async.bulk.prefetch.L2),
`SYNCS.ARRIVE.TRANS64` / `SYNCS.PHASECHK.TRANS64.TRYWAIT` (mbarrier), `REDUX`, `VOTE`, `LDG.E.128.STRONG.GPU` / `STG.E.128.STRONG.GPU` (tile-state words).
"""
open(os.path.join(ROOT, "profiles", "r2_scan_strip_ncu.md"), "w").write(md)

# ---- the single-pass insertion kernel (fused_assemble_kernel) on the headline stream
rep_i = os.path.join(ROOT, "gpurun_out", "r2_insert_fused.ncu-rep")
if os.path.exists(rep_i):
    g, num = raw(rep_i)
    rd, wr, ms = num("dram__bytes_read.sum"), num("dram__bytes_write.sum"), num("gpu__time_duration.sum")
    tab = "\n".join(f"| `{k}` | {g(k)[0]} {g(k)[1]} |" for k in WANT)
    tot, tops = top_lines(rep_i, 10)
    tl = "\n".join(f"| {100 * c / tot:.1f} % | `{f}:{ln}` | `{src}` |" for c, f, ln, src in tops)
    ins = bench.get("insert", {})
    alg = float(ins.get("out_bytes", 0)) * 2.0
    cmd_i = ("ncu --set full --clock-control none --import-source on -k regex:fused_assemble_kernel -s 2 -c 1 python bench.py --steps 4 --warmup 3 --no-sweep "
             "--no-parse --no-rewrite --no-cpu-extras --e2e-gib 0.25")
    md_i = f"""# ncu --set full, fused_assemble_kernel (single-pass rbsp_to_nal), round 2

`{cmd_i}` (report: `gpurun_out/r2_insert_fused.ncu-rep`, scratch).  Workload: `hevcb_insert_device` over the image and extents of the 4 GiB
headline stream (16 KiB NALs); the output must equal the original stream (checked in `bench.py`).
Bench value (CUDA events, no profiler, all 7 launches of a call): {ins.get('ms', 0):.3f} ms = {ins.get('input_GBps', 0):.0f} GB/s of input = **{ins.get('frac_of_peak', 0):.3f}** of the
measured HBM peak when scored as SURVEY 8d says (N_rbsp + N_nal); round 1 (count + scan + write, two reads of the payload): 2.77 ms, 0.47.

| metric | value |
|---|---|
{tab}

DRAM traffic per launch: read {rd / 1e9:.3f} GB + write {wr / 1e9:.3f} GB = **{(rd + wr) / 1e9:.3f} GB** = {(rd + wr) / alg if alg else 0:.3f}x the algorithmic bytes
({alg / 1e9:.3f} GB): every source byte is read once (TMA bulk copies into shared memory), every output byte written once.

Warp-state samples by source line (top {len(tops)} of {tot} samples):

| share | line | source |
|---|---|---|
{tl}

Reading: the samples at the `__syncthreads()` in front of `long long o = sm.excl;` are the seven warps of a tile waiting for warp 0 to
receive the tile's exclusive prefix from the scanner CTA: about a third of a tile's lifetime.  The scanner answers within one or two
L2 round trips of the last aggregate of its batch; what a tile waits for is the slowest EARLIER tile (in-order prefix), at six
33 KiB tiles per SM (all of the shared memory).  SASS: `UBLKCP.S.G` (cp.async.bulk), `SYNCS.*TRANS64` (mbarrier), `LDG/STG.E.64.STRONG.GPU`
(tile states).
"""
    open(os.path.join(ROOT, "profiles", "r2_insert_ncu.md"), "w").write(md_i)

rows = [r for r in csv.reader(open(os.path.join(ROOT, "profiles", "r2_launches.csv"))) if r]
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
lines = ["# ncu launch list summary, round 2", "",
         "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv python bench.py --steps 2 --warmup 3 --no-sweep --e2e-gib 0.25 --no-cpu-extras` "
         "(B200, cold-cache serialised launch times: compare SHARES, not absolutes).",
         "Raw CSV: `profiles/r2_launches.csv` (warm-up + timed steps of the headline workload, then the e2e, parse, insert and rewrite sections of the bench).", "",
         "| kernel | launches | total us | share |", "|---|---|---|---|"]
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| `{k}` | {n} | {us:.1f} | {100 * us / tot:.1f}% |")
per = lambda name: agg.get(name, [1, 0.0])[1] / max(agg.get(name, [1, 0.0])[0], 1)
step = sum(per(k) for k in ("hevcb_scan_strip_kernel", "hevcb_scan_emit_kernel", "hevcb_scan_finalize_kernel", "hevcb_scan_init_kernel"))
lines += ["", f"Per scan step the launches are: memset (tile states), `hevcb_scan_init_kernel` ({per('hevcb_scan_init_kernel'):.1f} us), `hevcb_scan_strip_kernel` "
          f"({per('hevcb_scan_strip_kernel'):.1f} us on average over the differently sized scans of the run, cooperative), `hevcb_scan_emit_kernel` ({per('hevcb_scan_emit_kernel'):.1f} us), "
          f"`hevcb_scan_finalize_kernel` ({per('hevcb_scan_finalize_kernel'):.1f} us): the strip kernel is {100 * per('hevcb_scan_strip_kernel') / step:.1f}% of a step's kernel time, "
          "which is the share `bench.py` attributes the roofline to (its CUDA-event time covers all of them)."]
open(os.path.join(ROOT, "profiles", "r2_launch_summary.md"), "w").write("\n".join(lines) + "\n")
print(md[:1500])
print("\n".join(lines[7:16]))
