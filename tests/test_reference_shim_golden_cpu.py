"""The oracle against golden vectors produced by RUNNING THE REFERENCE'S OWN MODEL CODE
(nets/resnet_model.py + nets/blocks.py + nets/model_helper.py) through a TF-1.14 API stand-in
(tests/golden/tf1_shim, generator tests/golden/make_reference_shim_golden.py; executed in the build
container where /root/reference is mounted -- this test only reads the committed json).

Pinned: variable names / shapes / initializer kinds / creation order of 8 model configurations
(including the Assemble-ResNet-50 of the north star, ResNet-D, SE, proj anti-alias, zero-gamma, R101,
R152), and the logits + updated BN moving statistics for seeded inputs and seeded variable values, in
inference and in training mode."""
import hashlib
import importlib.util
import json
import os

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location(
    "make_reference_shim_golden", os.path.join(HERE, "golden", "make_reference_shim_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_shim_golden.json")))

KIND = {"conv_kernel": "variance_scaling", "dense_kernel": "glorot_uniform", "dense_bias": "zeros",
        "beta": "zeros", "moving_mean": "zeros", "moving_variance": "ones"}


def _oracle(flags, use_resnet_d, size):
    from oracle import model as M
    model, vs = M.build(seed=1, input_hw=size, use_resnet_d=use_resnet_d, **flags)
    return M, model, vs


def _close(a, b, rel, abs_=1e-6):
    return abs(a - b) <= max(rel * max(abs(a), abs(b)), abs_)


@pytest.mark.parametrize("name", sorted(mg.CONFIGS))
def test_variable_inventory_matches_reference_code(name):
    flags, d, batch, size = mg.CONFIGS[name]
    gold = GOLD[name]
    _, _, vs = _oracle(flags, d, size)
    names = list(vs.vars)
    assert len(names) == gold["num_variables"]
    assert names[:3] == gold["first_names"] and names[-2:] == gold["last_names"]
    h = hashlib.sha256()
    zero_gammas = 0
    for n in names:
        kind = vs.kind[n]
        init = KIND.get(kind)
        if kind == "gamma":
            init = "zeros" if float(vs.vars[n].abs().sum()) == 0.0 else "ones"
            zero_gammas += init == "zeros"
        h.update(("%s|%s|%s|%d\n" % (n, ",".join(map(str, vs.vars[n].shape)), init,
                                     bool(vs.trainable[n]))).encode())
    assert zero_gammas == gold["zero_init_gammas"]
    assert h.hexdigest() == gold["names_sha256"]


@pytest.mark.parametrize("name", [n for n in sorted(mg.CONFIGS) if "152" not in n and "101" not in n])
def test_forward_matches_reference_code(name):
    flags, d, batch, size = mg.CONFIGS[name]
    gold = GOLD[name]
    M, model, vs = _oracle(flags, d, size)
    for i, n in enumerate(list(vs.vars)):
        vs.vars[n] = mg.seeded_value(i, n, tuple(vs.vars[n].shape))
    x = mg.seeded_input(batch, size)
    with torch.no_grad():
        y = M.forward(model, vs, x, training=False, use_resnet_d=d)
    got = mg.digest(y)
    for k in ("sum", "abs_sum", "first", "last"):
        assert _close(got[k], gold["eval_logits"][k], 2e-4, 1e-4), (k, got[k], gold["eval_logits"][k])
    for a, b in zip(y[0, :8].tolist(), gold["eval_logits_row0_head"]):
        assert _close(a, b, 2e-4, 1e-4)
    # training mode: batch statistics + moving-average updates (UPDATE_OPS)
    with torch.no_grad():
        y = M.forward(model, vs, x, training=True, use_resnet_d=d)
    got = mg.digest(y)
    assert _close(got["abs_sum"], gold["train_logits"]["abs_sum"], 5e-3), (got, gold["train_logits"])
    after = dict(vs.vars)
    after.update(model.bn_updates)          # the UPDATE_OPS of this step
    mm = torch.cat([after[n].flatten() for n in vs.vars if n.endswith("moving_mean")])
    mv = torch.cat([after[n].flatten() for n in vs.vars if n.endswith("moving_variance")])
    for k in ("sum", "abs_sum"):
        assert _close(mg.digest(mm)[k], gold["moving_mean_after"][k], 1e-4)
        assert _close(mg.digest(mv)[k], gold["moving_variance_after"][k], 1e-4)
