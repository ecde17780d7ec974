#!/bin/bash
# final validation of the native (model-level C ABI) product path: full GPU suite, bench lines of c3 / c1 / c2 / c5
mkdir -p gpurun_out
(timeout 700 python -m pytest tests -m gpu -q -s --durations=12 > gpurun_out/r02h_gpu_tests.log 2>&1; echo tests rc=$? >> gpurun_out/r02h_gpu_tests.log); tail -30 gpurun_out/r02h_gpu_tests.log; grep -h "C2 full size\|C-ABI step" gpurun_out/r02h_gpu_tests.log
timeout 300 python bench.py > gpurun_out/bench_c3_native.json 2> gpurun_out/bench_c3_native.err; cut -c1-400 gpurun_out/bench_c3_native.json; tail -2 gpurun_out/bench_c3_native.err
for cfg in c1 c2 c5; do timeout 200 python bench.py --config $cfg --no-cpu-baseline > gpurun_out/bench_${cfg}_native.json 2> gpurun_out/bench_${cfg}_native.err; cut -c1-260 gpurun_out/bench_${cfg}_native.json; echo; done
