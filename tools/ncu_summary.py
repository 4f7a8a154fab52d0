"""Turns an `ncu --set full --import-source on` report of tools/profile_target.py into the tables kept under
profiles/: one column per k_bounce launch of a wave, the DRAM traffic per launch (profiles/traverse_traffic.json,
read by bench.py's `roofline.traffic`) and, per launch, where the warp instructions go (shade / walk / primitive
tests / filing) from the SASS page of the report.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_summary.md [--traffic profiles/traverse_traffic.json]
"""
import csv, io, json, subprocess, sys

METRICS = [
    ("duration [ms]", "gpu__time_duration.sum", 1.0),
    ("registers / thread", "launch__registers_per_thread", 1.0),
    ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    ("lanes / instruction", "smsp__thread_inst_executed_per_inst_executed.ratio", 1.0),
    ("IPC (of 4)", "sm__inst_executed.avg.per_cycle_elapsed", 1.0),
    ("L1 wavefront pipe %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("L1 hit % (all sectors, stores count as misses)", "l1tex__t_sector_hit_rate.pct", 1.0),
    ("L2 hit %", "lts__t_sector_hit_rate.pct", 1.0),
    ("DRAM read [GB]", "dram__bytes_read.sum", 1.0),
    ("DRAM write [GB]", "dram__bytes_write.sum", 1.0),
    ("FMA pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1.0),
    ("ALU pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1.0),
    ("FP64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 1.0),
    ("L2 throughput % of peak", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("SM throughput % of peak", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
]
UNIT_TO_GB = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}
UNIT_TO_MS = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "second": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}


def ncu_csv(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def raw_page(rep):
    rows = ncu_csv(rep, "--page", "raw")
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[2:]:
        if len(r) != len(hdr) or "k_bounce" not in r[col["Kernel Name"]]:
            continue
        d = {"name": r[col["Kernel Name"]]}
        for label, m, _ in METRICS:
            if m not in col:
                continue
            try:
                v, u = float(r[col[m]].replace(",", "")), units[col[m]]
            except ValueError:      # empty or "no data"
                continue
            if m.startswith("dram__bytes"):
                v *= UNIT_TO_GB.get(u, 1.0)
            if m.startswith("gpu__time"):
                v *= UNIT_TO_MS.get(u, 1.0)
            d[label] = v
        stalls = {}
        for h in hdr:
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                stalls[h[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(r[col[h]].replace(",", "") or 0)
        tot = sum(stalls.values()) or 1.0
        d["stalls"] = ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:4])
        launches.append(d)
    return launches


def sass_regions(rep, launch_index):
    """Warp instructions and stall samples of one launch by code region.  The regions are found from landmarks in
    the SASS: the first 256-bit node load (walk), the primitive-record load after the node loop (primitive tests),
    the last warp vote of the walk loop (filing starts after it)."""
    rows = ncu_csv(rep, "--page", "source", "--kernel-id", f":::{launch_index + 1}")
    # several kernels may match the id filter; keep the first section
    sec, hdr = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            if sec:
                break
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            sec.append(r)
    if not sec:
        return None
    c = {h: i for i, h in enumerate(hdr)}
    I = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
    ins = [(r[c["Source"]], I(r[c["# Samples"]]), I(r[c["Instructions Executed"]]), I(r[c["Thread Instructions Executed"]])) for r in sec]
    node = [i for i, x in enumerate(ins) if ".256" in x[0] and "LDG" in x[0]]
    if not node:
        return None
    votes = [i for i, x in enumerate(ins) if x[0].strip().startswith("VOTE.ANY P")]
    walk_begin = max([v for v in votes if v < node[0]], default=node[0])
    walk_end = max([v for v in votes if v > node[0]], default=node[-1])
    leaf_begin = next((i for i in range(node[-1], walk_end) if "LDG.E.128.CONSTANT" in ins[i][0]), walk_end)
    subr = next((i for i, x in enumerate(ins) if i > walk_end and x[0].strip().startswith("EXIT")), len(ins))
    regions = [("acquire + shade + ray set-up", 0, walk_begin), ("walk: inner nodes", walk_begin, leaf_begin),
               ("walk: primitive tests", leaf_begin, walk_end + 1), ("file hits", walk_end + 1, subr), ("division / sqrt slow paths", subr, len(ins))]
    ts, ti = sum(x[1] for x in ins) or 1, sum(x[2] for x in ins) or 1
    out = []
    for name, lo, hi in regions:
        s = sum(x[1] for x in ins[lo:hi]); n = sum(x[2] for x in ins[lo:hi]); t = sum(x[3] for x in ins[lo:hi])
        out.append((name, 100.0 * n / ti, 100.0 * s / ts, t / max(n, 1)))
    return out


def main():
    rep, out_md = sys.argv[1], sys.argv[2]
    traffic_path = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
    wave_paths = int(sys.argv[sys.argv.index("--wave-paths") + 1]) if "--wave-paths" in sys.argv else 1920 * 1080 * 16
    what = sys.argv[sys.argv.index("--what") + 1] if "--what" in sys.argv else "tools/profile_target.py (C2, 1920x1080, 16 spp = 33.2 M paths, depth 8)"
    L = raw_page(rep)
    if not L:
        sys.exit("no k_bounce launches in " + rep)
    n = len(L)
    lines = [f"## `ncu --set full --clock-control none --import-source on`: the {n} k_bounce launches of one wave",
             "(segment 0 = camera rays; 1..depth-1 = shade the previous segment's hits + trace; last = shade only)", "",
             "| metric | " + " | ".join(f"seg {i}" for i in range(n)) + " |", "|---|" + "---|" * n]
    for label, _, _ in METRICS:
        lines.append(f"| {label} | " + " | ".join(f"{d.get(label, float('nan')):.4g}" for d in L) + " |")
    lines.append("| DRAM [GB/s] (read + write) / duration | " + " | ".join(f"{(d['DRAM read [GB]'] + d['DRAM write [GB]']) / (d['duration [ms]'] * 1e-3):.0f}" for d in L) + " |")
    lines.append("| top stall reasons | " + " | ".join(d["stalls"] for d in L) + " |")
    rd, wr = sum(d["DRAM read [GB]"] for d in L), sum(d["DRAM write [GB]"] for d in L)
    lines += ["", f"DRAM traffic of the wave: {rd + wr:.2f} GB (read {rd:.2f}, write {wr:.2f}); sum of durations {sum(d['duration [ms]'] for d in L):.3f} ms.", ""]
    for k in (0, 1, 2):
        if k >= n:
            break
        reg = sass_regions(rep, k)
        if not reg:
            continue
        lines += [f"### where launch {k} spends its warp instructions (SASS page, regions by landmark instructions)", "",
                  "| region | warp instructions % | stall samples % | lanes / instruction |", "|---|---|---|---|"]
        lines += [f"| {nm} | {wi:.1f} | {ss:.1f} | {ln:.1f} |" for nm, wi, ss, ln in reg]
        lines.append("")
    open(out_md, "a").write("\n".join(lines) + "\n")
    if traffic_path:
        per = [(d["DRAM read [GB]"] + d["DRAM write [GB]"]) * 1e9 for d in L]
        json.dump({"kernel": "k_bounce", "dram_bytes_per_launch": sum(per) / n, "per_launch_bytes": per, "launches": n,
                   "dram_bytes_per_wave": sum(per), "wave_paths": wave_paths,
                   "ipc_time_weighted": sum(d["IPC (of 4)"] * d["duration [ms]"] for d in L) / sum(d["duration [ms]"] for d in L),
                   "lanes_per_instruction_time_weighted": sum(d["lanes / instruction"] * d["duration [ms]"] for d in L) / sum(d["duration [ms]"] for d in L),
                   "source": f"ncu --set full ({rep}), the {n} k_bounce launches of one wave of {what}: dram__bytes_read.sum + dram__bytes_write.sum"},
                  open(traffic_path, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
