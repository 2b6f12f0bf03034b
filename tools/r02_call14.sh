#!/bin/bash
# Round 2, GPU call (one B200): the device FASTQ parser -- parity tests, throughput.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parse.py -x -q 2>&1 | tail -5
timeout 600 python tools/parse_bench.py --mb 1024 > $O/r02_parse_bench.json 2> $O/r02_parse_bench.err; cat $O/r02_parse_bench.json; tail -3 $O/r02_parse_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"newline|record_kernel|gather|scan_" -c 24 --csv --log-file $O/r02_parse_launches.csv \
    python tools/parse_bench.py --mb 1024 --reps 1 > /dev/null 2>&1
grep -v "^==" $O/r02_parse_launches.csv | python -c "
import csv, sys
rows = list(csv.DictReader(sys.stdin))
from collections import OrderedDict
agg = OrderedDict()
for r in rows[-33:]:
    k = (r['ID'], r['Kernel Name'][:40]); agg.setdefault(k, {})[r['Metric Name']] = r['Metric Value']
for k, v in agg.items(): print(k, v)
"
echo done
