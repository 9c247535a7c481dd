#!/bin/bash
# Debug helper: device-resident throughput of bench.py as a function of the number of concurrent streams.
for st in 1 2 3 4; do
  timeout 300 python bench.py --no-cpu-baseline --streams $st 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('streams', $st, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'])"
done
