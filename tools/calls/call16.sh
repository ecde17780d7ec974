#!/bin/bash
# model-level C ABI on the GPU: native vs Python executor bit-identity, C-ABI host-array sequence, smoke
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_native_model_gpu.py tests/test_dp_nccl_gpu.py -q -s > gpurun_out/r02g_native_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02g_native_tests.log; tail -25 gpurun_out/r02g_native_tests.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02g_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r02g_smoke.log; tail -5 gpurun_out/r02g_smoke.log
