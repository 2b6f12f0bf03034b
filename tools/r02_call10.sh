#!/bin/bash
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
for c in pe150 se100 pe250_full; do
  timeout 600 python bench.py --config $c --no-e2e --no-cpu > $O/r02_q_${c}.json 2> $O/r02_q_${c}.err; show $O/r02_q_${c}.json
done
for v in st16 st20; do
  for c in pe150 se100 pe250_full; do
    AQC_LIB_PATH=$PWD/gpurun_variants/libaqc_$v.so timeout 600 python bench.py --config $c --no-e2e --no-cpu > $O/r02_q_${c}_$v.json 2> $O/r02_q_${c}_$v.err; show $O/r02_q_${c}_$v.json
  done
done
echo done
