"""CPU: the loss oracle against the reference-generated goldens, and the drop-in Tacotron2Loss class (PyTorch formulation on
CPU tensors) against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle.loss_oracle import make_loss_case, tacotron2_loss
from tacotron2_subword_b200 import Tacotron2Loss

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("alignloss", ["", "L2"])
def test_loss_oracle_matches_reference_golden(alignloss):
    z = np.load(os.path.join(GOLDEN, f"loss_{alignloss or 'default'}.npz"))
    c = make_loss_case(int(z["B"]), int(z["T"]), int(z["T_in"]), int(z["seed"]))
    losses, grads = tacotron2_loss(c, alignloss)
    for k in ("total", "mel_loss", "gate_loss"):
        assert abs(float(losses[k]) - float(z[k])) <= 2e-6 * max(1.0, abs(float(z[k]))), k
    for k in ("align_loss", "align_bert_loss"):
        if alignloss == "L2":
            assert abs(float(losses[k]) - float(z[k])) <= 2e-6
        else:
            assert losses[k] is None and np.isnan(float(z[k]))
    for k, g in grads.items():
        if g is None:
            assert f"g_{k}/max" not in z.files
            continue
        scale = float(z[f"g_{k}/max"])
        assert abs(float(g.abs().max()) - scale) <= 1e-5 * scale
        assert abs(float(g.sum()) - float(z[f"g_{k}/sum"])) <= 1e-5 * scale * g.numel() ** 0.5
        assert float((g.reshape(-1)[:16].float() - torch.from_numpy(z[f"g_{k}/head"])).abs().max()) <= 1e-5 * scale


@pytest.mark.parametrize("alignloss", ["", "L2"])
def test_loss_class_cpu_path_matches_oracle(alignloss):
    c = make_loss_case(2, 17, 9, 5)
    outs = [c[k].clone().requires_grad_(True) for k in ("mel", "mel_postnet", "gate", "align", "align_bert")]
    total, mel_loss, gate_loss, al, alb = Tacotron2Loss(alignloss)(outs, (c["mel_target"], c["gate_target"], c["align_target"]), None, 0)
    total.backward()
    losses, grads = tacotron2_loss(c, alignloss)
    assert abs(float(total) - float(losses["total"])) <= 1e-5
    assert (al is None) == (alignloss == "") and (alb is None) == (alignloss == "")
    for o, k in zip(outs, ("mel", "mel_postnet", "gate", "align", "align_bert")):
        if grads[k] is None:
            assert o.grad is None
        else:
            assert float((o.grad.double() - grads[k]).abs().max()) <= 1e-6 * float(grads[k].abs().max()) + 1e-9


def test_loss_class_l2_stops_at_40000_iterations():
    c = make_loss_case(2, 5, 4, 6)
    outs = [c[k] for k in ("mel", "mel_postnet", "gate", "align", "align_bert")]
    out = Tacotron2Loss("L2")(outs, (c["mel_target"], c["gate_target"], c["align_target"]), None, 40000)
    assert out[3] is None and out[4] is None           # loss_function.py:29 `iters < 40000`
