#!/bin/bash
# GPU call: conv kernel tests (halo v2 loops, double-buffered output staging), A/Bs of the two knobs,
# halo-vs-im2col layer timings, ncu --set full of the halo kernel, bench lines of the other configs
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gemm_gpu.py -x -q > gpurun_out/conv_tests.log 2>&1; echo "rc=$?" >> gpurun_out/conv_tests.log
tail -3 gpurun_out/conv_tests.log
for shp in "112 64 32" "56 64 128" "56 32 64" "112 64 64" "112 32 64"; do set -- $shp
  for m in 0 2; do echo "== H=$1 $2->$3 halo=$m"; ACNN_CONV_HALO=$m timeout 60 python tools/profile_layer.py --H $1 --Cin $2 --Cout $3 --k 3 --which fprop,dgrad --reps 5; done
done > gpurun_out/halo_layers.txt 2>&1
cat gpurun_out/halo_layers.txt
timeout 150 python tools/exp_ab.py --knob acnn_set_conv_out_bufs --values 1,0,2 > gpurun_out/ab_out_bufs.txt 2>&1; cat gpurun_out/ab_out_bufs.txt
timeout 120 python tools/exp_ab.py --knob acnn_set_conv_halo --values 0,1 > gpurun_out/ab_halo_v2.txt 2>&1; cat gpurun_out/ab_halo_v2.txt
ACNN_CONV_HALO=2 timeout 150 ncu --set full --import-source on -k regex:conv_halo -s 2 -c 1 -f -o gpurun_out/halo_112_64_32 python tools/profile_layer.py --H 112 --Cin 64 --Cout 32 --k 3 --which fprop > gpurun_out/ncu1.log 2>&1
for cfg in c1 c2 c5; do timeout 200 python bench.py --config $cfg --no-cpu-baseline > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err; done
tail -c 400 gpurun_out/bench_c5.json
