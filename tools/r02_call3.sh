#!/bin/bash
# Round 2, GPU call (one B200): lane-per-read stat_kernel -- parity, bench of every config, ncu of the statistics kernel.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
O=gpurun_out
echo "== gpu tests (parity + bench size) =="
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02_pytest_gpu.log 2>&1; echo "exit $?"; tail -4 $O/r02_pytest_gpu.log
echo "== bench: the four configs =="
for c in pe150 se100 pe250_full pe150_err3; do
  timeout 900 python bench.py --config $c > $O/r02_bench_$c.json 2> $O/r02_bench_$c.err; echo "$c exit $?"
  python - <<PY
import json
try:
    j = json.load(open("$O/r02_bench_$c.json"))
    r = j["roofline"]
    print("$c", "value", round(j["value"], 1), j["unit"], "ms/step", round(j["ms_per_step"], 3), "e2e", j["e2e"] and round(j["e2e"]["value"] or 0, 1), j["e2e"] and j["e2e"].get("mode"),
          "cpu", j["cpu_baseline"] and round(j["cpu_baseline"]["value"], 2), "clocks", j["clocks"]["sm_mhz"])
    for p in r["phases"]:
        print("   ", round(p["ms"], 3), "ms", round(p["frac"], 3), "of peak |", p["launches"][:70])
except Exception as e:
    print("$c: no line", e)
PY
done
B="python bench.py --pairs 2000000 --no-e2e --no-cpu"
echo "== ncu: statistics kernel at qc_sample 0 =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stat_kernel -s 6 -c 2 -o $O/r02_stat_qc0_full \
    $B --qc-sample 0 --steps 1 --warmup 3 > $O/r02_stat_qc0_full.log 2>&1
ncu -i $O/r02_stat_qc0_full.ncu-rep --page raw --csv > $O/r02_stat_qc0_full_raw.csv 2>/dev/null
ncu -i $O/r02_stat_qc0_full.ncu-rep --page details > $O/r02_stat_qc0_full_details.txt 2>/dev/null
grep -n "stat_kernel<\|Duration\|Executed Ipc Active\|Issue Slots Busy\|Executed Instructions  " $O/r02_stat_qc0_full_details.txt | head -30
echo done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lane_kernel -s 3 -c 1 -o $O/r02_lane_full \
    $B --qc-sample 40000 --steps 1 --warmup 3 > $O/r02_lane_full.log 2>&1
ncu -i $O/r02_lane_full.ncu-rep --page raw --csv > $O/r02_lane_full_raw.csv 2>/dev/null
ncu -i $O/r02_lane_full.ncu-rep --page details > $O/r02_lane_full_details.txt 2>/dev/null
grep -n "lane_kernel<\|Duration\|Executed Ipc Active\|Issue Slots Busy\|Executed Instructions  \|Registers Per\|Achieved Occ" $O/r02_lane_full_details.txt | head
echo done2
