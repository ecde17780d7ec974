#!/bin/bash
# GPU call: halo split-epilogue tests, layer timings (split 0 / 1), A/B of the step
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gemm_gpu.py -x -q -k halo > gpurun_out/conv_tests3.log 2>&1; echo "rc=$?" >> gpurun_out/conv_tests3.log
tail -3 gpurun_out/conv_tests3.log
for shp in "112 64 32" "56 64 128" "112 64 64" "112 32 64" "56 32 64"; do set -- $shp
  for sp in 0 1; do echo "== H=$1 $2->$3 halo=2 split=$sp"; ACNN_CONV_HALO=2 ACNN_CONV_HALO_SPLIT=$sp timeout 60 python tools/profile_layer.py --H $1 --Cin $2 --Cout $3 --k 3 --which fprop,dgrad --reps 5; done
done > gpurun_out/halo_layers3.txt 2>&1
cat gpurun_out/halo_layers3.txt
timeout 150 python tools/exp_ab.py --knob acnn_set_conv_halo_split --values 0,1 > gpurun_out/ab_halo_split.txt 2>&1; cat gpurun_out/ab_halo_split.txt
timeout 150 python tools/exp_ab.py --knob acnn_set_conv_halo --values 0,1 > gpurun_out/ab_halo_v3.txt 2>&1; cat gpurun_out/ab_halo_v3.txt
ACNN_CONV_HALO=2 timeout 150 ncu --set full --import-source on -k regex:conv_halo -s 2 -c 1 -f -o gpurun_out/halo3_112_64_32 python tools/profile_layer.py --H 112 --Cin 64 --Cout 32 --k 3 --which fprop > gpurun_out/ncu3.log 2>&1
