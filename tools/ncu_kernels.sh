#!/bin/bash
# ncu --set full captures of the main kernels (one launch each) on the bench workload, plus the launch list of bench.py
# itself; run under gpurun.  Only compact JSON summaries are kept: gpurun_out/ is limited to 64 MiB.
# usage: bash tools/ncu_kernels.sh [round tag, default r02] [kernel:skip ...]
R=${1:-r02}
shift
SPECS=${@:-"k_spmv_win:3 k_cg_dots:2 k_pairlist:2 k_hessian:1 k_enbond:1 k_e4b_eval:1 k_e3b_eval:1 k_ehb_eval:1 k_boprim:1 k_cg_update1:2"}
mkdir -p gpurun_out/ncu_$R
for spec in $SPECS; do
  IFS=: read name skip <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -o /tmp/${name}_s$skip -f python tools/profile_step.py --steps 2 > /tmp/${name}_s$skip.log 2>&1
  python tools/ncu_summarize.py /tmp/${name}_s$skip.ncu-rep > gpurun_out/ncu_$R/${name}_s$skip.json
done
# the launch list of the bench command (cold-cache, serialised per-launch times: shares of the step, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
python tools/agg_launches.py gpurun_out/${R}_launches.csv 60 > gpurun_out/${R}_launches_summary.txt
ls -la gpurun_out/ncu_$R/ | tail -20; head -30 gpurun_out/${R}_launches_summary.txt
