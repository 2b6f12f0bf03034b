#!/bin/bash
# launch list (per-kernel durations) of two steps of the headline config and of the full-pipeline config
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
O=gpurun_out
for c in pe150 pe250_full se100; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stamp|stat_kernel|lane_kernel|pair_kernel|maxlen" --csv --log-file $O/r02_launches_$c.csv \
    python bench.py --config $c --steps 2 --warmup 3 --no-e2e --no-cpu > $O/r02_launches_$c.log 2>&1
python - $O/r02_launches_$c.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
print(sys.argv[1], len(rows), "launches; last step:")
for r in rows[-14:]:
    print("   %-60s grid %-14s %10.1f us" % (r[4][:60], r[8], float(r[-1]) / 1e3))
PY
done
echo done
