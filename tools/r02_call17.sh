#!/bin/bash
# clock sampling: default run and a run whose timed region is shorter than the sampling period
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python bench.py --no-e2e --no-cpu > $O/r02_clk_default.json 2>/dev/null; python -c "
import json; j=json.load(open('$O/r02_clk_default.json')); print(round(j['value'],1), j['clocks'])"
timeout 600 python bench.py --no-e2e --no-cpu --steps 3 --warmup 3 > $O/r02_clk_short.json 2>/dev/null; python -c "
import json; j=json.load(open('$O/r02_clk_short.json')); print(round(j['value'],1), j['clocks'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 1 --steps 20 --warmup 5 --no-e2e --no-cpu > $O/r02_clk_torchrun.json 2>/dev/null; python -c "
import json; j=json.load(open('$O/r02_clk_torchrun.json')); print(round(j['value'],1), j['clocks'])"
echo done
