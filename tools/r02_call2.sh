#!/bin/bash
# Round 2, GPU call (one B200): new lane_kernel / stat_kernel -- parity suite, bench of every config, ncu evidence.
#   gpurun --timeout 2400 -- 'bash tools/r02_call2.sh'
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
O=gpurun_out
echo "== smoke =="
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "== gpu tests =="
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02_pytest_gpu.log 2>&1; echo "exit $?"; tail -4 $O/r02_pytest_gpu.log
echo "== bench: the four configs =="
for c in pe150 se100 pe250_full pe150_err3; do
  timeout 900 python bench.py --config $c > $O/r02_bench_$c.json 2> $O/r02_bench_$c.err; echo "$c exit $?"
  python - <<EOF
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
EOF
done
timeout 600 python bench.py --config pe150 --filter-kernel warp --no-e2e --no-cpu > $O/r02_bench_pe150_warp.json 2> $O/r02_bench_pe150_warp.err
B="python bench.py --pairs 2000000 --no-e2e --no-cpu"
echo "== launch list of one short bench run (our kernels only) =="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aqc --csv --log-file $O/r02_launches_pe150.csv \
    $B --qc-sample 40000 --steps 2 --warmup 3 > $O/r02_launches_bench.log 2>&1
grep -c aqc $O/r02_launches_pe150.csv
echo "== full captures =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lane_kernel -s 3 -c 1 -o $O/r02_lane_full \
    $B --qc-sample 40000 --steps 1 --warmup 3 > $O/r02_lane_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stat_kernel -s 12 -c 4 -o $O/r02_stat_full \
    $B --qc-sample 40000 --steps 1 --warmup 3 > $O/r02_stat_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stat_kernel -s 12 -c 4 -o $O/r02_stat_qc0_full \
    $B --qc-sample 0 --steps 1 --warmup 3 > $O/r02_stat_qc0_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lane_kernel -s 3 -c 1 -o $O/r02_lane_pe250_full \
    python bench.py --config pe250_full --pairs 1000000 --no-e2e --no-cpu --steps 1 --warmup 3 > $O/r02_lane_pe250_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lane_kernel -s 3 -c 1 -o $O/r02_lane_se100_full \
    python bench.py --config se100 --pairs 2000000 --no-e2e --no-cpu --steps 1 --warmup 3 > $O/r02_lane_se100_full.log 2>&1
for r in r02_lane_full r02_stat_full r02_stat_qc0_full r02_lane_pe250_full r02_lane_se100_full; do
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
  ncu -i $O/$r.ncu-rep --page details > $O/${r}_details.txt 2>/dev/null
done
rm -f $O/r02_lane_pe250_full.ncu-rep $O/r02_lane_se100_full.ncu-rep $O/r02_stat_full.ncu-rep
echo done
