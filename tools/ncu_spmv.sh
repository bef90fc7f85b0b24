#!/bin/bash
# ncu --set full of one k_spmv_rows16 launch for two staging variants; keeps the raw metric page (CSV) and the summaries.
mkdir -p gpurun_out/ncu
for rows in 4 40; do
  RXG_SPMV_ROWS=$rows ncu --set full --clock-control none --import-source on -k regex:k_spmv_rows16 -s 5 -c 1 -o /tmp/spmv16_r$rows -f python tools/profile_step.py --steps 1 > /tmp/spmv16_r$rows.log 2>&1
  ncu -i /tmp/spmv16_r$rows.ncu-rep --page raw --csv > gpurun_out/ncu/spmv16_r${rows}_raw.csv
  python tools/ncu_summarize.py /tmp/spmv16_r$rows.ncu-rep > gpurun_out/ncu/k_spmv_rows16_r$rows.json
done
ls -la gpurun_out/ncu/
