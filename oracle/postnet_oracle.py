"""CPU restatement of the reference Postnet in eval mode (TEST INFRASTRUCTURE ONLY -- see oracle/README.md).

/root/reference/model.py:27-70: five Conv1d(k=5, 'same' padding) + BatchNorm1d, tanh after all but the last,
F.dropout(0.5, self.training) after each (identity in eval).  BERT_Tacotron2.forward adds the result to the decoder mel
(model.py:557-558) and parse_output zeroes everything beyond output_lengths (model.py:531-541).

Pinned by tests/golden/postnet_eval.npz (outputs of the unmodified reference class, oracle/make_golden.py).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

BN_EPS = 1e-5   # nn.BatchNorm1d default, model.py:41,52,63


def postnet_weight_shapes(n_mel: int = 80, dim: int = 512, k: int = 5, n_convs: int = 5) -> Dict[str, tuple]:
    chans = [n_mel] + [dim] * (n_convs - 1) + [n_mel]
    out = {}
    for i in range(n_convs):
        p = f"convolutions.{i}."
        out[p + "0.conv.weight"] = (chans[i + 1], chans[i], k)
        out[p + "0.conv.bias"] = (chans[i + 1],)
        for n in ("weight", "bias", "running_mean", "running_var"):
            out[p + "1." + n] = (chans[i + 1],)
    return out


def make_postnet_weights(seed: int = 1234, n_mel: int = 80, dim: int = 512, k: int = 5, n_convs: int = 5) -> Dict[str, torch.Tensor]:
    """Seeded weights + non-trivial BatchNorm statistics (a freshly initialised BN would hide folding bugs)."""
    g = torch.Generator().manual_seed(seed)
    w = {}
    for key, shape in postnet_weight_shapes(n_mel, dim, k, n_convs).items():
        if key.endswith("conv.weight"):
            bound = (6.0 / ((shape[0] + shape[1]) * shape[2])) ** 0.5 * (5.0 / 3.0)      # xavier_uniform, tanh gain
            w[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif key.endswith("conv.bias"):
            w[key] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        elif key.endswith("1.weight"):
            w[key] = 0.5 + torch.rand(shape, generator=g)
        elif key.endswith("1.bias"):
            w[key] = (torch.rand(shape, generator=g) * 2 - 1) * 0.2
        elif key.endswith("running_mean"):
            w[key] = torch.randn(shape, generator=g) * 0.1
        else:
            w[key] = 0.5 + torch.rand(shape, generator=g)
    return w


def postnet_eval(w: Dict[str, torch.Tensor], x: torch.Tensor, n_convs: int = 5) -> torch.Tensor:
    """model.py:64-70 with self.training == False.  x [B, n_mel, T] -> [B, n_mel, T]."""
    for i in range(n_convs):
        p = f"convolutions.{i}."
        k = w[p + "0.conv.weight"].shape[2]
        x = F.conv1d(x, w[p + "0.conv.weight"], w[p + "0.conv.bias"], padding=(k - 1) // 2)          # layers.py:26-27
        x = (x - w[p + "1.running_mean"][None, :, None]) / torch.sqrt(w[p + "1.running_var"][None, :, None] + BN_EPS)
        x = x * w[p + "1.weight"][None, :, None] + w[p + "1.bias"][None, :, None]
        if i < n_convs - 1:
            x = torch.tanh(x)
    return x


def mel_postnet(w: Dict[str, torch.Tensor], mel: torch.Tensor, output_lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mel + postnet(mel) (model.py:557-558), zero beyond output_lengths when given (model.py:531-541)."""
    out = mel + postnet_eval(w, mel)
    if output_lengths is not None:
        T = mel.shape[2]
        invalid = torch.arange(T)[None, :] >= output_lengths[:, None]
        out = out.masked_fill(invalid[:, None, :], 0.0)
    return out
