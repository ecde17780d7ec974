#!/bin/bash
# printed parity numbers of the full-size C2 test and the C-ABI tests; eager latency native vs python executor
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_parity_fp32_gpu.py tests/test_native_model_gpu.py -q -s -k "full_size or c_abi or c_host" > gpurun_out/r02i_parity_prints.log 2>&1; echo "rc=$?" >> gpurun_out/r02i_parity_prints.log; grep -h "C2 full size\|C-ABI step\|step [0-9]:\|passed\|failed" gpurun_out/r02i_parity_prints.log | cut -c1-300
timeout 120 python tools/exp_eager_latency.py > gpurun_out/r02i_eager_latency.txt 2>&1; cat gpurun_out/r02i_eager_latency.txt | tail -6
