#!/bin/bash
# GPU call: round-2 final ncu launch list (time + DRAM / L2 traffic) of one training step, and
# --set full captures of the top kernels (raw pages exported on the box: the .ncu-rep files are large)
mkdir -p gpurun_out
timeout 420 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02f_launches_dram.csv python tools/profile_step.py --ncu > gpurun_out/ncu6.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02f_launches_dram.csv
cap() { # name regex H Cin Cout k which
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o gpurun_out/r02f_$1 python tools/profile_layer.py --H $3 --Cin $4 --Cout $5 --k $6 --which $7 > gpurun_out/ncu6_$1.log 2>&1
  ncu -i gpurun_out/r02f_$1.ncu-rep --page raw --csv > gpurun_out/r02f_$1.raw.csv 2>/dev/null
}
cap halo_112_64_32_fprop conv_halo 112 64 32 3 fprop
cap halo_56_64_128_fprop conv_halo 56 64 128 3 fprop
cap conv_14_512_1024_fprop conv_gemm_kernel 14 512 1024 3 fprop
cap conv_28_64_256_1x1_fprop conv_gemm_kernel 28 64 256 1 fprop
cap wgrad_14_512_1024 wgrad_gemm 14 512 1024 3 wgrad
ls -la gpurun_out | tail -20
