#!/bin/bash
# GPU call: split epilogue of conv_gemm_kernel (tests, A/B), then the round-2 final ncu evidence
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_conv_gemm_gpu.py -x -q > gpurun_out/tests7.log 2>&1; echo "rc=$?" >> gpurun_out/tests7.log
tail -3 gpurun_out/tests7.log
timeout 200 python tools/exp_ab.py --knob acnn_set_conv_split_epilogue --values 0,1,2 > gpurun_out/ab_split_epi.txt 2>&1; cat gpurun_out/ab_split_epi.txt
for shp in "28 64 256" "14 128 512" "7 256 1024" "7 1024 256" "14 512 128"; do set -- $shp
  for sp in 0 2; do echo "== H=$1 $2->$3 k1 split=$sp"; ACNN_CONV_SPLIT_EPI=$sp timeout 60 python tools/profile_layer.py --H $1 --Cin $2 --Cout $3 --k 1 --which fprop,dgrad --reps 10; done
done > gpurun_out/split_layers.txt 2>&1
cat gpurun_out/split_layers.txt
timeout 420 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02f_launches_dram.csv python tools/profile_step.py --ncu > gpurun_out/ncu7.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02f_launches_dram.csv
cap() { # name regex H Cin Cout k which
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o /tmp/r02f_$1 python tools/profile_layer.py --H $3 --Cin $4 --Cout $5 --k $6 --which $7 > gpurun_out/ncu7_$1.log 2>&1
  ncu -i /tmp/r02f_$1.ncu-rep --page raw --csv > gpurun_out/r02f_$1.raw.csv 2>/dev/null
  rm -f /tmp/r02f_$1.ncu-rep
}
cap halo_112_64_32_fprop conv_halo 112 64 32 3 fprop
cap halo_56_64_128_fprop conv_halo 56 64 128 3 fprop
cap conv_14_512_1024_fprop conv_gemm_kernel 14 512 1024 3 fprop
cap conv_28_64_256_1x1_fprop conv_gemm_kernel 28 64 256 1 fprop
cap wgrad_14_512_1024 wgrad_gemm 14 512 1024 3 wgrad
du -sh gpurun_out
