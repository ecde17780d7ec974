#!/bin/bash
# GPU call: programmatic dependent launch re-measured on the current kernels; final launch list
mkdir -p gpurun_out
timeout 250 python tools/exp_ab.py --knob acnn_set_pdl --values 0,2,1 > gpurun_out/ab_pdl.txt 2>&1; cat gpurun_out/ab_pdl.txt
timeout 420 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02f_launches_dram.csv python tools/profile_step.py --ncu > gpurun_out/ncu13.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02f_launches_dram.csv
timeout 300 python tools/exp_skip.py > gpurun_out/exp_skip_final.txt 2>&1; tail -32 gpurun_out/exp_skip_final.txt
