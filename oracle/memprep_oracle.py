"""CPU restatement of the decoder-input step (TEST INFRASTRUCTURE ONLY -- see oracle/README.md).

/root/reference/model.py:548-549 (and :553-554 for the sub-word stream):
    decoder_inputs = linear_converter(cat([encoder_outputs, phoneme_embeddings_cls], 2))      LinearNorm(512 + 768 -> 512, bias)
/root/reference/model.py:258-261 (Decoder.initialize_decoder_states):
    processed_memory = attention_layer.memory_layer(memory)                                   LinearNorm(512 -> 128, no bias)
LinearNorm is torch.nn.Linear (layers.py:8-20).

Pinned by tests/golden/memprep.npz: outputs of the unmodified reference modules (layers.LinearNorm instances wired as in
model.py), written by oracle/make_golden.py.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F


def make_memprep_weights(seed: int = 1234, enc: int = 512, cls: int = 768, attn: int = 128) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)

    def xavier(o, i, gain=1.0):
        bound = gain * (6.0 / (i + o)) ** 0.5
        return (torch.rand(o, i, generator=g) * 2 - 1) * bound

    return {"linear_converter.linear_layer.weight": xavier(enc, enc + cls),
            "linear_converter.linear_layer.bias": (torch.rand(enc, generator=g) * 2 - 1) * (1.0 / (enc + cls)) ** 0.5,
            "memory_layer.linear_layer.weight": xavier(attn, enc, gain=5.0 / 3.0)}


def make_memprep_inputs(B: int, T: int, seed: int = 1, enc: int = 512, cls: int = 768) -> Tuple[torch.Tensor, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    # encoder outputs are BiLSTM activations (|x| < 1); BERT embeddings have a heavier tail
    return torch.tanh(torch.randn(B, T, enc, generator=g)), 0.6 * torch.randn(B, T, cls, generator=g)


def memory_prepare(w: Dict[str, torch.Tensor], encoder_outputs: torch.Tensor, cls_embeddings: torch.Tensor,
                   dtype: torch.dtype = torch.float32):
    """-> (memory [B,T,enc], processed_memory [B,T,attn]); model.py:548-549 then :258-261."""
    x = torch.cat([encoder_outputs.to(dtype), cls_embeddings.to(dtype)], 2)                                    # model.py:548
    memory = F.linear(x, w["linear_converter.linear_layer.weight"].to(dtype), w["linear_converter.linear_layer.bias"].to(dtype))   # :549
    pm = F.linear(memory, w["memory_layer.linear_layer.weight"].to(dtype))                                     # model.py:258-261
    return memory, pm
