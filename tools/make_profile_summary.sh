#!/bin/bash
# (round 1 helper; round 2 assembled profiles/r02_*.md from tools/ncu_summary.py directly) Turns the files a profiling gpurun call left in gpurun_out/ into profiles/r01_<tag>_summary.md (+ the raw launch list,
# the bench line and profiles/traverse_traffic.json).   usage: tools/make_profile_summary.sh v12 "title text"
set -e
tag="$1"; title="$2"
cd "$(dirname "$0")/.."
out=profiles/r01_${tag}_summary.md
printf "# r01 %s -- %s\n\n\`python tools/profile_target.py\` (C2 scene, 1920x1080, 16 spp = one 33.2 M-path wave per frame, depth 8); plain run on the same box: %s.  bench.py's step is one 64-spp wave of the same kernels (launches four times as long).\n\n" "$tag" "$title" "$(head -1 gpurun_out/plain_${tag}.log)" > $out
python tools/ncu_summary.py gpurun_out/prof_r01_${tag}.ncu-rep $out --traffic profiles/traverse_traffic.json > /dev/null
cp gpurun_out/launches_bench_${tag}.csv profiles/r01_${tag}_bench_launches.csv
cp gpurun_out/bench_${tag}.json profiles/r01_bench_1gpu.json
python - "$tag" <<'PY'
import csv, collections, json, sys
tag = sys.argv[1]
rows = list(csv.reader(open(f'profiles/r01_{tag}_bench_launches.csv')))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; c = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[hi + 2:]:
    if len(r) < len(hdr): continue
    k = r[c['Kernel Name']][:60]; v = float(r[c['Metric Value']].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'nsecond': 1e-3, 'msecond': 1e3}.get(r[c['Metric Unit']], 1)
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
b = json.load(open('profiles/r01_bench_1gpu.json'))
out = ["", f"## launch list of the benchmark itself (`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 3 --no-baselines`, raw: r01_{tag}_bench_launches.csv)", "",
       "| kernel | launches | total us | share |", "|---|---|---|---|"]
out += [f"| {k} | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% |" for k, a in agg.items()]
kb = sum(a[1] for k, a in agg.items() if 'k_bounce' in k)
out.append(f"\nk_bounce = {100 * kb / tot:.1f} % of the device time under ncu (every launch of the captured window: build, warm-up, timed steps, the instrumented 2-spp render of the fp32 view, the e2e leg); bench.py's CUDA events give {100 * b['roofline']['share_of_step']:.1f} % of the step (`roofline.share_of_step`).")
out.append(f"\nbench line of the same build (`profiles/r01_bench_1gpu.json`): {b['value']:.0f} Mpaths/s device-timed, {b['e2e']['value']:.0f} end to end, {b['ms_per_step']:.2f} ms per 64-spp frame, clocks {b['clocks']['sm_mhz']:.0f} MHz, reasons {b['clocks']['reasons']}.")
open(f'profiles/r01_{tag}_summary.md', 'a').write("\n".join(out) + "\n")
print("\n".join(out[-3:]))
PY
