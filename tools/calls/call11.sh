#!/bin/bash
# GPU call: split epilogue for two-M-tile CTA tiles -- tests, then interleaved A/B on the step
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_conv_gemm_gpu.py -x -q > gpurun_out/tests11.log 2>&1; echo "rc=$?" >> gpurun_out/tests11.log
tail -3 gpurun_out/tests11.log
timeout 250 python tools/exp_ab.py --knob acnn_set_conv_split_mt2 --values 0,1,2 > gpurun_out/ab_split_mt2.txt 2>&1; cat gpurun_out/ab_split_mt2.txt
