#!/bin/bash
# Round 2, GPU call (one B200): full parity suite + smoke + default bench with the final library.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1800 python -m pytest tests -x -q -m gpu > $O/r02_last_pytest_gpu.log 2>&1; echo "exit $?"; tail -3 $O/r02_last_pytest_gpu.log
timeout 900 python bench.py > $O/r02_last_bench.json 2> $O/r02_last_bench.err; echo "bench exit $?"
python - <<PY
import json
j = json.load(open("$O/r02_last_bench.json"))
print("value", round(j["value"], 1), "ms/step", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"], 1), j["e2e"]["mode"], "frac", round(j["roofline"]["frac"], 3), "traffic", j["roofline"]["traffic"], "clocks", j["clocks"])
PY
echo done
