#!/bin/bash
# Round 2, GPU call (one B200): stat_kernel word path with head/tail from registers -- parity, warps-per-CTA sweep, ncu.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
O=gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1]))
    print(sys.argv[1].split("/")[-1], "value", round(j["value"], 1), "ms/step", round(j["ms_per_step"], 3), " phases:", [round(p["ms"], 3) for p in j["roofline"]["phases"]])
except Exception as e:
    print(sys.argv[1], "no line", e)
PY
}
echo "== gpu tests =="
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02_pytest_gpu_v7.log 2>&1; echo "exit $?"; tail -3 $O/r02_pytest_gpu_v7.log
B="python bench.py --pairs 2000000 --no-e2e --no-cpu"
for v in default st24; do
  L=""; [ $v != default ] && L="$PWD/gpurun_variants/libaqc_$v.so"
  for c in pe150 se100 pe250_full; do
    AQC_LIB_PATH=$L timeout 600 python bench.py --config $c --no-e2e --no-cpu > $O/r02_v7_${c}_$v.json 2> $O/r02_v7_${c}_$v.err; show $O/r02_v7_${c}_$v.json
  done
  AQC_LIB_PATH=$L timeout 600 $B --qc-sample 0 --steps 5 --warmup 3 > $O/r02_v7_qc0_$v.json 2>/dev/null; show $O/r02_v7_qc0_$v.json
done
echo "== ncu: stat_kernel =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stat_kernel -s 12 -c 2 -o $O/r02_v7_stat_qc0_full \
    $B --qc-sample 0 --steps 1 --warmup 3 > $O/r02_v7_stat_qc0_full.log 2>&1
ncu -i $O/r02_v7_stat_qc0_full.ncu-rep --page raw --csv > $O/r02_v7_stat_qc0_full_raw.csv 2>/dev/null
ncu -i $O/r02_v7_stat_qc0_full.ncu-rep --page details > $O/r02_v7_stat_qc0_full_details.txt 2>/dev/null
ncu -i $O/r02_v7_stat_qc0_full.ncu-rep --page source --csv > $O/r02_v7_stat_qc0_full_source.csv 2>/dev/null
echo done
