#!/bin/bash
# lane_kernel as one CTA per SM: parity, resident numbers of every config, CTA-size sweep on pe150
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
O=gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1]))
    print(sys.argv[1].split("/")[-1], "value", round(j["value"], 1), "ms/step", round(j["ms_per_step"], 3), " phases:", [round(p["ms"], 3) for p in j["roofline"]["phases"]], " filter frac", round(j["roofline"]["phases"][1]["frac"], 3))
except Exception as e:
    print(sys.argv[1], "no line", e)
PY
}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zzz_bench_size.py -x -q -m gpu > $O/r02_pytest_gpu.log 2>&1; echo "exit $?"; tail -2 $O/r02_pytest_gpu.log
for c in pe150 se100 pe250_full pe150_err3; do
  timeout 600 python bench.py --config $c --no-e2e --no-cpu > $O/r02_q_$c.json 2> $O/r02_q_$c.err; show $O/r02_q_$c.json
done
for v in w16 w18 w22; do
  AQC_LIB_PATH=$PWD/gpurun_variants/libaqc_$v.so timeout 600 python bench.py --config pe150 --no-e2e --no-cpu > $O/r02_q_pe150_$v.json 2> $O/r02_q_pe150_$v.err; show $O/r02_q_pe150_$v.json
done
for w in 8 10 11; do
  AQC_LANE_WARPS=$w timeout 600 python bench.py --config pe250_full --no-e2e --no-cpu > $O/r02_q_pe250_w$w.json 2> $O/r02_q_pe250_w$w.err; show $O/r02_q_pe250_w$w.json
done
echo done
