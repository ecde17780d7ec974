#!/bin/bash
# GPU call (final validation of the committed defaults): full GPU test suite, smoke, bench lines
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_gpu_tests.log 2>&1; echo tests rc=$? >> gpurun_out/r02e_gpu_tests.log); tail -3 gpurun_out/r02e_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_c3_final.json 2> gpurun_out/bench_c3_final.err; cut -c1-300 gpurun_out/bench_c3_final.json
for cfg in c1 c2 c5; do timeout 200 python bench.py --config $cfg --no-cpu-baseline > gpurun_out/bench_${cfg}_final.json 2> gpurun_out/bench_${cfg}_final.err; cut -c1-260 gpurun_out/bench_${cfg}_final.json; echo; done
timeout 120 python tools/profile_step.py --csv gpurun_out/conv_layers_final.csv > gpurun_out/opbreak_final.txt 2>&1; head -12 gpurun_out/opbreak_final.txt; grep "conv GEMMs" gpurun_out/opbreak_final.txt
