#!/bin/bash
# Round 2, GPU call (one B200): stat_kernel with the four-bases-per-step word path -- parity, bench of every config, ncu.
#   gpurun --timeout 1800 -- 'bash tools/r02_call11.sh'
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
O=gpurun_out
echo "== smoke =="
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== gpu tests =="
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02_pytest_gpu_v4.log 2>&1; echo "exit $?"; tail -4 $O/r02_pytest_gpu_v4.log
echo "== bench =="
for c in pe150 se100 pe250_full pe150_err3; do
  timeout 900 python bench.py --config $c > $O/r02_v4_bench_$c.json 2> $O/r02_v4_bench_$c.err; echo "$c exit $?"
  python - <<PY
import json
try:
    j = json.load(open("$O/r02_v4_bench_$c.json"))
    r = j["roofline"]
    print("$c", "value", round(j["value"], 1), j["unit"], "ms/step", round(j["ms_per_step"], 3), "e2e", j["e2e"] and round(j["e2e"]["value"] or 0, 1), "clocks", j["clocks"]["sm_mhz"])
    for p in r["phases"]:
        print("   ", round(p["ms"], 3), "ms", round(p["frac"], 3), "of peak |", p["launches"][:70])
except Exception as e:
    print("$c: no line", e)
PY
done
B="python bench.py --pairs 2000000 --no-e2e --no-cpu"
timeout 600 $B --qc-sample 0 --steps 5 --warmup 3 > $O/r02_v4_bench_qc0.json 2>/dev/null
python -c "
import json; j=json.load(open('$O/r02_v4_bench_qc0.json')); print('qc0 2M pairs', round(j['value'],1), [ (round(p['ms'],3), p['launches'][:40]) for p in j['roofline']['phases']])"
echo "== ncu: stat_kernel =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stat_kernel -s 12 -c 2 -o $O/r02_v4_stat_qc0_full \
    $B --qc-sample 0 --steps 1 --warmup 3 > $O/r02_v4_stat_qc0_full.log 2>&1
ncu -i $O/r02_v4_stat_qc0_full.ncu-rep --page raw --csv > $O/r02_v4_stat_qc0_full_raw.csv 2>/dev/null
ncu -i $O/r02_v4_stat_qc0_full.ncu-rep --page details > $O/r02_v4_stat_qc0_full_details.txt 2>/dev/null
ncu -i $O/r02_v4_stat_qc0_full.ncu-rep --page source --csv > $O/r02_v4_stat_qc0_full_source.csv 2>/dev/null
echo done
