#!/bin/bash
# sweep of the SpMV launch shape (development aid): prints avg SpMV launch ms per setting
run() { python bench.py --no-cpu --no-e2e --steps 4 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['roofline']['avg_launch_ms'],4), 'ms  step', round(d['ms_per_step'],2))"; }
for rows in $@; do RXG_SPMV_ROWS=$rows run "rows=$rows"; done
