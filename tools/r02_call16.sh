#!/bin/bash
# Round 2, GPU call (one B200): lane_kernel prefilter with 1 / 2 / 4 / 8 independent shift-in chains.
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
for v in default ch1 ch2 ch8; do
  L=""; [ $v != default ] && L="$PWD/gpurun_variants/libaqc_$v.so"
  for c in pe150 pe150_err3 pe250_full; do
    AQC_LIB_PATH=$L timeout 600 python bench.py --config $c --no-e2e --no-cpu > $O/r02_ch_${c}_$v.json 2> $O/r02_ch_${c}_$v.err; show $O/r02_ch_${c}_$v.json
  done
done
echo done
