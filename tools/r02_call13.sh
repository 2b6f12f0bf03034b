#!/bin/bash
# Round 2, final 1-GPU call: smoke, parity suite, the four bench lines (with e2e and CPU arm), reference arm, launch lists, DRAM traffic.
#   gpurun --timeout 2400 -- 'bash tools/r02_call13.sh'
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
O=gpurun_out
echo "== smoke =="
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== gpu tests =="
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02_final_pytest_gpu.log 2>&1; echo "exit $?"; tail -3 $O/r02_final_pytest_gpu.log
echo "== bench: the four configs =="
for c in pe150 se100 pe250_full pe150_err3; do
  timeout 900 python bench.py --config $c > $O/r02_final_bench_$c.json 2> $O/r02_final_bench_$c.err; echo "$c exit $?"
  python - <<PY
import json
try:
    j = json.load(open("$O/r02_final_bench_$c.json"))
    r = j["roofline"]
    print("$c", "value", round(j["value"], 1), j["unit"], "ms/step", round(j["ms_per_step"], 3), "e2e", j["e2e"] and round(j["e2e"]["value"] or 0, 1), j["e2e"] and j["e2e"].get("mode"),
          "cpu", j["cpu_baseline"] and round(j["cpu_baseline"]["value"], 2), "clocks", j["clocks"]["sm_mhz"], "launches", j["gpu_launches"])
    for p in r["phases"]:
        print("   ", round(p["ms"], 3), "ms", round(p["frac"], 3), "of peak |", p["launches"][:70])
except Exception as e:
    print("$c: no line", e)
PY
done
echo "== default invocation, as the driver runs it =="
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_final_bench_default.json 2>/dev/null; head -c 600 $O/r02_final_bench_default.json; echo
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/r02_final_bench_reference.json 2>/dev/null; head -c 700 $O/r02_final_bench_reference.json; echo
B="python bench.py --pairs 2000000 --no-e2e --no-cpu"
echo "== launch lists =="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aqc --csv --log-file $O/r02_final_launches_pe150.csv \
    $B --qc-sample 40000 --steps 2 --warmup 3 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aqc --csv --log-file $O/r02_final_launches_se100.csv \
    python bench.py --config se100 --pairs 2000000 --no-e2e --no-cpu --steps 2 --warmup 3 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aqc --csv --log-file $O/r02_final_launches_pe250_full.csv \
    python bench.py --config pe250_full --pairs 1000000 --no-e2e --no-cpu --steps 2 --warmup 3 > /dev/null 2>&1
grep -c aqc $O/r02_final_launches_pe150.csv $O/r02_final_launches_se100.csv $O/r02_final_launches_pe250_full.csv
echo "== DRAM bytes per launch (qc0, 2 M pairs) =="
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"stat_kernel|lane_kernel" -s 8 -c 6 --csv --log-file $O/r02_final_dram_qc0.csv \
    $B --qc-sample 0 --steps 2 --warmup 2 > /dev/null 2>&1
tail -20 $O/r02_final_dram_qc0.csv | cut -c1-220
echo done
