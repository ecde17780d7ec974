"""Eager (no CUDA graph) latency of the public call `model(inputs, training=False)` -- BASELINE config 1,
vanilla ResNet-50, batch 1, 224 x 224 -- through the two executors: the library's launch records
(native, the default) and the Python plan + one ctypes call per op.  Same kernels, same plan; what differs
is the host time per launch, which is what bounds a batch-1 forward that is not graph-captured."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from assembled_cnn_b200.model_fns import Model

x = (torch.randn(1, 224, 224, 3) * 64).clamp(-124, 152).cuda()
for batch in (1, 8):
    xb = x.expand(batch, -1, -1, -1).contiguous()
    for native in (True, False):
        model = Model(50, resnet_version=1, native=native)
        for _ in range(5):
            model(xb, training=False)
        torch.cuda.synchronize()
        ts = []
        for _ in range(30):
            t0 = time.perf_counter()
            out = model(xb, training=False)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        # host time to ENQUEUE the forward (no sync): what the CPU spends per call
        t0 = time.perf_counter()
        for _ in range(30):
            model(xb, training=False)
        enq = (time.perf_counter() - t0) * 1e3 / 30
        torch.cuda.synchronize()
        ts.sort()
        print("batch %d, %s executor: eager model(x, False) latency median %.3f ms (min %.3f), host enqueue "
              "%.3f ms per call" % (batch, "native (acnn_run_ops)" if native else "python (ctypes per op)",
                                    ts[len(ts) // 2], ts[0], enq))
