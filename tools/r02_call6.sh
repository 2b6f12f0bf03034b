#!/bin/bash
# Round 2, 2-GPU call: NCCL CLI path, directory mode over the GPUs, sharded bench with shard_parity
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi topo -m > $O/topo2.txt 2>&1; nproc
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_golden.py -x -q -m gpu -k "two_gpus or spreads" > $O/r02_pytest_gpu2.log 2>&1; echo "exit $?"; tail -4 $O/r02_pytest_gpu2.log
for c in pe150 pe250_full; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --config $c --steps 10 --warmup 3 > $O/r02_bench2_$c.json 2> $O/r02_bench2_$c.err; echo "$c exit $?"
  python - $O/r02_bench2_$c.json <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1]))
    print(sys.argv[1].split("/")[-1], "value", round(j["value"], 1), "ms/step", round(j["ms_per_step"], 3), "e2e", j["e2e"] and round(j["e2e"]["value"] or 0, 1),
          "per-gpu h2d GB/s", j["e2e"] and round(j["e2e"].get("per_gpu_h2d_GBps", 0), 1), "shard_parity", j.get("shard_parity"), j.get("shard_parity_note"), j.get("numa_rank0"))
except Exception as e:
    print(sys.argv[1], "no line", e)
PY
  tail -3 $O/r02_bench2_$c.err
done
echo done
