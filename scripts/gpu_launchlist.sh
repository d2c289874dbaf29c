#!/bin/bash
# usage: scripts/gpu_launchlist.sh <tag> <cfg...>   -> per-kernel durations of bench_configs.py for the given configs
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches.csv \
    python scripts/bench_configs.py "$@" > $OUT/${TAG}_ncu.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("$OUT/${TAG}_launches.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hd = rows[h]; kn = hd.index("Kernel Name"); mv = hd.index("Metric Value"); gs = hd.index("Grid Size"); bs = hd.index("Block Size")
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= mv: continue
    k = (r[kn][:110], r[gs], r[bs])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[mv].replace(",", ""))
with open("$OUT/${TAG}_launch_summary.txt", "w") as f:
    for k, (n, t) in agg.items():
        line = "%6d x %10.1f us avg  grid %s block %s  %s" % (n, t / n / 1e3, k[1], k[2], k[0])
        print(line); f.write(line + "\n")
PY
