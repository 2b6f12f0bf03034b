#!/bin/bash
# First GPU call of round 2 (one B200): everything round 1 could not measure for lane_kernel, in one go.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# Outputs under gpurun_out/: lane_parity.log (oracle matrix on hardware), bench.json (auto-selected kernel), bench_warp.json,
# launches.csv (ncu launch list of one bench step), lane_full.ncu-rep (+ lane_full_raw.csv), stat_lane_full.ncu-rep, parity_*.log /
# full_*.json (lane2 and the stat_kernel = 2 candidates), sweep (lane tuning variants).
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."

echo "== lane kernel: full oracle matrix ==" | tee gpurun_out/lane_parity.log
timeout 900 python tests/lane_gpu_check.py parity >> gpurun_out/lane_parity.log 2>&1; echo "exit $?" | tee -a gpurun_out/lane_parity.log

echo "== lane2 kernel and the lane-per-read statistics (stat_kernel = 2): oracle matrix, then 2 M pairs vs the warp kernels =="
for c in lane2 lane_st2 lane_st3; do timeout 600 python tests/lane_gpu_check.py parity $c > gpurun_out/parity_$c.log 2>&1; echo "$c parity exit $?"; tail -1 gpurun_out/parity_$c.log; done
for c in lane2 warp_st2 lane_st2 lane_st3 lane2_st3; do timeout 300 python tests/lane_gpu_check.py full 2000000 $c > gpurun_out/full_$c.json 2> gpurun_out/full_$c.err; echo "$c full exit $?"; cat gpurun_out/full_$c.json; done

echo "== packed transport of the host-buffer entry (AQC_BATCH_PACK_BASES / _QUALS) vs the oracle =="
AQC_CHUNK_PAIRS=3000 timeout 300 python tests/lane_gpu_check.py pack > gpurun_out/pack_parity.log 2>&1; echo "pack parity exit $?"; tail -1 gpurun_out/pack_parity.log

echo "== gpu test suite =="
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log

echo "== bench: auto selection, then the warp kernel for reference =="
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json
timeout 600 python bench.py --filter-kernel warp --no-cpu > gpurun_out/bench_warp.json 2> gpurun_out/bench_warp.err

echo "== ncu: launch list of one short bench run, then a full capture of lane_kernel =="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --filter-kernel lane --pairs 2000000 --qc-sample 40000 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lane_kernel -s 3 -c 1 -o gpurun_out/lane_full \
    python bench.py --filter-kernel lane --pairs 2000000 --qc-sample 40000 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/lane_full_bench.log 2>&1
ncu -i gpurun_out/lane_full.ncu-rep --page raw --csv > gpurun_out/lane_full_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stat_lane_kernel -s 3 -c 1 -o gpurun_out/stat_lane_full \
    python bench.py --filter-kernel lane --stat-kernel lane_post --pairs 2000000 --qc-sample 40000 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/stat_lane_full_bench.log 2>&1
ncu -i gpurun_out/stat_lane_full.ncu-rep --page raw --csv > gpurun_out/stat_lane_full_raw.csv 2>/dev/null

echo "== lane tuning variants (built here with nvcc, benchmarked with parity) =="
python tools/variant_sweep.py build lane_base= lane_imad=-DAQC_LANE_IMAD_SHIFT lane_2plane=-DAQC_LANE_TWO_PLANE_FILTER lane_unrollconv=-DAQC_LANE_UNROLL_CONVERT > gpurun_out/sweep_build.log 2>&1
timeout 1200 python tools/variant_sweep.py run --filter-kernel lane --steps 10 > gpurun_out/sweep_run.log 2>&1
for w in 1 2 3; do AQC_LANE_WARPS=$w timeout 200 python bench.py --filter-kernel lane --no-e2e --no-cpu --steps 10 > gpurun_out/bench_lane_warps$w.json 2>/dev/null; done
echo done
