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
for v in imad st24 st28; do
  for c in pe150 se100; do
    AQC_LIB_PATH=$PWD/gpurun_variants/libaqc_$v.so timeout 600 python bench.py --config $c --no-e2e --no-cpu > $O/r02_q_${c}_$v.json 2> $O/r02_q_${c}_$v.err; show $O/r02_q_${c}_$v.json
  done
done
echo done
