"""CPU restatement of the reference loss (TEST INFRASTRUCTURE ONLY -- see oracle/README.md).

/root/reference/loss_function.py:12-66 (``Tacotron2Loss.forward``), variants "" (default, hparams.py:66) and "L2":
    mel_loss  = mean((mel - target)^2) + mean((mel_postnet - target)^2)                        :22
    gate_loss = mean(max(x,0) - x*y + log(1 + exp(-|x|)))  over gate.view(-1,1)                :23  (nn.BCEWithLogitsLoss)
    L2, iters < 40000:  align_loss = mean((align - align_target)^2), same for align_bert        :30-32
    total = mel_loss + gate_loss [+ align_loss + align_bert_loss]                               :59-66
plus the analytic gradients autograd would produce for them (what the fused CUDA sweep writes).

Pinned by tests/golden/loss_*.npz: values and gradient digests from the unmodified reference class + torch.autograd
(oracle/make_golden.py::run_reference_loss).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch


def make_loss_case(B: int, T: int, T_in: int, seed: int, n_mel: int = 80) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    gate_t = (torch.rand(B, T, generator=g) > 0.9).float()
    return {"mel": r(B, n_mel, T), "mel_postnet": r(B, n_mel, T), "gate": 2.0 * r(B, T), "mel_target": r(B, n_mel, T),
            "gate_target": gate_t, "align": torch.rand(B, T, T_in, generator=g), "align_bert": torch.rand(B, T, T_in, generator=g),
            "align_target": torch.rand(B, T, T_in, generator=g)}


def tacotron2_loss(c: Dict[str, torch.Tensor], alignloss: str = "", iters: int = 0, dtype=torch.float64):
    """-> (dict of loss terms, dict of gradients of the TOTAL w.r.t. every model output)."""
    mel, post, gate = c["mel"].to(dtype), c["mel_postnet"].to(dtype), c["gate"].to(dtype)
    y, gy = c["mel_target"].to(dtype), c["gate_target"].to(dtype)
    n_mel, n_gate = mel.numel(), gate.numel()
    mel_loss = ((mel - y) ** 2).mean() + ((post - y) ** 2).mean()
    gate_loss = (gate.clamp(min=0) - gate * gy + torch.log1p(torch.exp(-gate.abs()))).mean()
    losses = {"mel_loss": mel_loss, "gate_loss": gate_loss, "align_loss": None, "align_bert_loss": None}
    grads = {"mel": 2 * (mel - y) / n_mel, "mel_postnet": 2 * (post - y) / n_mel, "gate": (torch.sigmoid(gate) - gy) / n_gate,
             "align": None, "align_bert": None}
    total = mel_loss + gate_loss
    if alignloss == "L2" and iters < 40000:
        at = c["align_target"].to(dtype)
        for k in ("align", "align_bert"):
            a = c[k].to(dtype)
            losses[k + "_loss"] = ((a - at) ** 2).mean()
            grads[k] = 2 * (a - at) / a.numel()
            total = total + losses[k + "_loss"]
    losses["total"] = total
    return losses, grads
