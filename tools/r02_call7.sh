#!/bin/bash
# lane_kernel CTA-shape experiments (resident numbers) + shard parity re-check happens in the next multi-GPU call
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
timeout 600 python bench.py --config pe150 --no-e2e --no-cpu > $O/r02_q_pe150.json 2> $O/r02_q_pe150.err; show $O/r02_q_pe150.json
for v in w20sync w20 w10sync; do
  AQC_LIB_PATH=$PWD/gpurun_variants/libaqc_$v.so timeout 600 python bench.py --config pe150 --no-e2e --no-cpu > $O/r02_q_pe150_$v.json 2> $O/r02_q_pe150_$v.err; show $O/r02_q_pe150_$v.json
done
AQC_LIB_PATH=$PWD/gpurun_variants/libaqc_w20sync.so timeout 600 python -m pytest tests/test_gpu_zzz_bench_size.py -x -q -m gpu -k "default_path and 40000" 2>&1 | tail -2
echo done
