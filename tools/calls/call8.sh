#!/bin/bash
# GPU call: is one M tile + split epilogue better than two M tiles per CTA?  (interleaved A/B arms)
mkdir -p gpurun_out
timeout 300 python tools/exp_ab.py --configs "acnn_set_conv_mtiles=-1,acnn_set_conv_split_epilogue=1;acnn_set_conv_mtiles=1,acnn_set_conv_split_epilogue=2;acnn_set_conv_mtiles=1,acnn_set_conv_split_epilogue=1;acnn_set_conv_mtiles=-1,acnn_set_conv_split_epilogue=2" > gpurun_out/ab_mt_split.txt 2>&1; cat gpurun_out/ab_mt_split.txt
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_c3_v8.json 2> gpurun_out/bench_c3_v8.err; cut -c1-330 gpurun_out/bench_c3_v8.json
