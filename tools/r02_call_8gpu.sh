#!/bin/bash
# Round 2, 8-GPU call: BASELINE configs[3] (PE250 20 M pairs, full pipeline) and configs[4] (PE150 200 M pairs, 3 % error)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi topo -m > $O/topo8.txt 2>&1; nproc; free -g | head -2
for c in pe150 pe250_full pe150_err3; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --config $c --steps 10 --warmup 3 > $O/r02_bench8_$c.json 2> $O/r02_bench8_$c.err; echo "$c exit $?"
  python - $O/r02_bench8_$c.json <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1]))
    print(sys.argv[1].split("/")[-1], "value", round(j["value"], 1), j["unit"], "ms/step", round(j["ms_per_step"], 3), "e2e", j["e2e"] and round(j["e2e"]["value"] or 0, 1),
          "per-gpu h2d GB/s", j["e2e"] and round(j["e2e"].get("per_gpu_h2d_GBps", 0), 1), "shard_parity", j.get("shard_parity"), j.get("numa_rank0"), "clocks", j["clocks"])
except Exception as e:
    print(sys.argv[1], "no line", e)
PY
  grep -v "OMP_NUM\|\*\*\*\*\|NCCL version" $O/r02_bench8_$c.err | tail -3
done
echo done
