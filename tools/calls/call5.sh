#!/bin/bash
# GPU call: mixed-precision epilogue arithmetic (FHADD / FHFMA / HSET2): conv + plan tests, step timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gemm_gpu.py tests/test_plan_gpu.py -x -q > gpurun_out/tests5.log 2>&1; echo "rc=$?" >> gpurun_out/tests5.log
tail -3 gpurun_out/tests5.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_c3_v5.json 2> gpurun_out/bench_c3_v5.err; cut -c1-330 gpurun_out/bench_c3_v5.json
timeout 120 python tools/profile_step.py --csv gpurun_out/conv_layers_v5.csv > gpurun_out/opbreak_v5.txt 2>&1; head -8 gpurun_out/opbreak_v5.txt; grep "conv GEMMs" gpurun_out/opbreak_v5.txt
timeout 150 ncu --set full --import-source on -k regex:conv_gemm_kernel -s 2 -c 1 -f -o gpurun_out/ncu_1x1_64_256 python tools/profile_layer.py --H 28 --Cin 64 --Cout 256 --k 1 --which fprop > gpurun_out/ncu5.log 2>&1
