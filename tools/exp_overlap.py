import os, sys, time
os.environ.setdefault('ACNN_NATIVE_PLAN', '0')   # per-op hooks live in the Python executor (same plan, same launches)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from assembled_cnn_b200.hparams import params_from_flags
from assembled_cnn_b200.model_fns import Model, Trainer
from assembled_cnn_b200 import runtime as R
from bench import MODEL_FLAGS, TRAIN_FLAGS, synth_batch

def measure(overlap, graph):
    orig = R.Runtime.run
    if not overlap:
        def run_no(self, ops, overlap_wgrad=False):
            return orig(self, ops, False)
        R.Runtime.run = run_no
    try:
        params = params_from_flags(batch_size=256, **MODEL_FLAGS, **TRAIN_FLAGS)
        model = Model(50, num_classes=1001, resnet_version=2, use_sk_block=True, anti_alias_type="sconv", anti_alias_filter_size=3)
        tr = Trainer(model, params, 224, 224, use_cuda_graph=graph)
        x, y = synth_batch(tr.input_batch, 224, 1234)
        x, y = x.cuda(), y.cuda()
        for _ in range(4): tr.train_step(x, y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): tr.train_step(x, y)
        e1.record(); torch.cuda.synchronize()
        print("overlap=%s graph=%s: %.2f ms/step" % (overlap, graph, e0.elapsed_time(e1) / 10), flush=True)
        del tr, model
        torch.cuda.empty_cache()
    finally:
        R.Runtime.run = orig

for ov in (False, True):
    for g in (True, False):
        measure(ov, g)
