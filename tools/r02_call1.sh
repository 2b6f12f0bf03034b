#!/bin/bash
# Round 2, GPU call 1 (one B200): ncu evidence for the kernels that carry the bench number.
#   gpurun --timeout 1200 -- 'bash tools/r02_call1.sh'
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
O=gpurun_out
B="python bench.py --filter-kernel lane2 --stat-kernel warp --pairs 2000000 --qc-sample 40000 --no-e2e --no-cpu"

echo "== launch list of one short bench run =="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r02_launches.csv \
    $B --steps 2 --warmup 3 > $O/r02_launches_bench.log 2>&1
tail -25 $O/r02_launches.csv | cut -c1-300

echo "== full captures =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lane2_kernel -s 3 -c 1 -o $O/r02_lane2_full \
    $B --steps 1 --warmup 3 > $O/r02_lane2_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 6 -c 2 -o $O/r02_pair_full \
    $B --steps 1 --warmup 3 > $O/r02_pair_full.log 2>&1
for r in r02_lane2_full r02_pair_full; do
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
  ncu -i $O/$r.ncu-rep --page details > $O/${r}_details.txt 2>/dev/null
done

echo "== statistics-heavy mode (qc_sample 0) launch list =="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_launches_qc0.csv \
    python bench.py --filter-kernel lane2 --stat-kernel warp --pairs 2000000 --qc-sample 0 --no-e2e --no-cpu --steps 1 --warmup 1 > $O/r02_launches_qc0_bench.log 2>&1
tail -12 $O/r02_launches_qc0.csv | cut -c1-300

echo "== plain bench (default flags, what the driver runs) =="
timeout 900 python bench.py > $O/r02_bench0.json 2> $O/r02_bench0.err; tail -c 1500 $O/r02_bench0.json
nvidia-smi topo -m > $O/topo.txt 2>&1; lscpu > $O/lscpu.txt 2>&1; numactl -H > $O/numa.txt 2>&1
echo done
