#!/bin/bash
# EXPERIMENT: CTA pairs (tcgen05 cta_group::2) for the N = 256 conv tiles: training-step time and the
# per-layer conv timings with ACNN_CONV_PAIRS=0 / 1.
for v in 0 1 0 1; do
  ACNN_CONV_PAIRS=$v python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from assembled_cnn_b200.hparams import params_from_flags
from assembled_cnn_b200.model_fns import Model, Trainer
from bench import MODEL_FLAGS, TRAIN_FLAGS, synth_batch
params = params_from_flags(batch_size=256, **MODEL_FLAGS, **TRAIN_FLAGS)
model = Model(50, num_classes=1001, resnet_version=2, use_sk_block=True, anti_alias_type="sconv", anti_alias_filter_size=3)
tr = Trainer(model, params, 224, 224, use_cuda_graph=True)
x, y = synth_batch(tr.input_batch, 224, 1234); x, y = x.cuda(), y.cuda()
for _ in range(3): tr.train_step(x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(15): tr.train_step(x, y)
e1.record(); torch.cuda.synchronize()
print("ACNN_CONV_PAIRS=%s : %.3f ms/step  loss %s" % (os.environ.get("ACNN_CONV_PAIRS"), e0.elapsed_time(e1) / 15, tr.train_step(x, y).tolist()), flush=True)
PY
done
