#!/bin/bash
# GPU call: A/B of the halo rule on the step, full GPU test suite, bench c3
mkdir -p gpurun_out
timeout 150 python tools/exp_ab.py --knob acnn_set_conv_halo --values 0,1 > gpurun_out/ab_halo_v4.txt 2>&1; cat gpurun_out/ab_halo_v4.txt
(timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gpu_tests.log 2>&1; echo tests rc=$? >> gpurun_out/r02c_gpu_tests.log); tail -3 gpurun_out/r02c_gpu_tests.log
timeout 200 python bench.py > gpurun_out/bench_c3_v4.json 2> gpurun_out/bench_c3_v4.err; cut -c1-400 gpurun_out/bench_c3_v4.json
timeout 120 python tools/profile_step.py --csv gpurun_out/conv_layers_v4.csv > gpurun_out/opbreak_v4.txt 2>&1
