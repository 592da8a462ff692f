"""Deterministic synthetic weights, inputs and dropout/noise replay plans.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): imported by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``; never by the product package.

Everything is drawn from a CPU ``torch.Generator`` with an explicit seed and an
explicit, fixed draw order, so the same tensors can be regenerated on the GPU
box (same image, same torch) without shipping 170 MB of weights.  The golden
fixtures under ``tests/golden/`` store a checksum of the weights they were made
with so RNG drift is detected instead of silently breaking parity.

Shapes/key names follow the reference ``Decoder`` state_dict
(/root/reference/model.py:142-207, attention.py:7-37, 305-322).
"""
from __future__ import annotations

import hashlib
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

SMA = "StepwiseMonotonicAttention"
LSA = "LocationSensitiveAttention"


@dataclass
class DecoderDims:
    """Decoder hyper-parameters the hot path reads (hparams.py:55,72-90)."""
    n_mel: int = 80
    enc: int = 512          # encoder_embedding_dim
    arnn: int = 1024        # attention_rnn_dim
    drnn: int = 1024        # decoder_rnn_dim
    prenet: int = 256
    attn: int = 128         # attention_dim
    loc_filters: int = 32
    loc_kernel: int = 31
    streams: int = 2        # 2 = BERT_Tacotron2 (char + sub-word), 1 = Tacotron2 compat


def decoder_weight_shapes(attention: str = SMA, dims: DecoderDims = DecoderDims()) -> Dict[str, tuple]:
    """Ordered {state_dict key (without the ``decoder.`` prefix): shape}."""
    d = dims
    s: Dict[str, tuple] = {}
    streams = ["", "_bert"] if d.streams == 2 else [""]
    for sfx in streams:
        s[f"prenet{sfx}.layers.0.linear_layer.weight"] = (d.prenet, d.n_mel)
        s[f"prenet{sfx}.layers.1.linear_layer.weight"] = (d.prenet, d.prenet)
    for sfx in streams:
        s[f"attention_rnn{sfx}.weight_ih"] = (4 * d.arnn, d.prenet + d.enc)
        s[f"attention_rnn{sfx}.weight_hh"] = (4 * d.arnn, d.arnn)
        s[f"attention_rnn{sfx}.bias_ih"] = (4 * d.arnn,)
        s[f"attention_rnn{sfx}.bias_hh"] = (4 * d.arnn,)
    for sfx in streams:
        p = f"attention_layer{sfx}."
        if attention == SMA:
            s[p + "memory_layer.linear_layer.weight"] = (d.attn, d.enc)
            s[p + "v.weight"] = (1, d.attn)
            s[p + "query_layer.linear_layer.weight"] = (d.attn, d.arnn)
        else:
            s[p + "query_layer.linear_layer.weight"] = (d.attn, d.arnn)
            s[p + "memory_layer.linear_layer.weight"] = (d.attn, d.enc)
            s[p + "v.linear_layer.weight"] = (1, d.attn)
            s[p + "location_layer.location_conv.conv.weight"] = (d.loc_filters, 2, d.loc_kernel)
            s[p + "location_layer.location_dense.linear_layer.weight"] = (d.attn, d.loc_filters)
    dec_in = d.streams * (d.arnn + d.enc)
    s["decoder_rnn.weight_ih"] = (4 * d.drnn, dec_in)
    s["decoder_rnn.weight_hh"] = (4 * d.drnn, d.drnn)
    s["decoder_rnn.bias_ih"] = (4 * d.drnn,)
    s["decoder_rnn.bias_hh"] = (4 * d.drnn,)
    if d.streams == 2:
        # dead in decode() (model.py:375-378) but part of the checkpoint contract
        s["decoder_rnn_bert.weight_ih"] = (4 * d.drnn, d.arnn + d.enc)
        s["decoder_rnn_bert.weight_hh"] = (4 * d.drnn, d.drnn)
        s["decoder_rnn_bert.bias_ih"] = (4 * d.drnn,)
        s["decoder_rnn_bert.bias_hh"] = (4 * d.drnn,)
    proj_in = d.drnn + d.streams * d.enc
    s["linear_projection.linear_layer.weight"] = (d.n_mel, proj_in)
    s["linear_projection.linear_layer.bias"] = (d.n_mel,)
    s["gate_layer.linear_layer.weight"] = (1, proj_in)
    s["gate_layer.linear_layer.bias"] = (1,)
    return s


def _bound_for(key: str, shape: tuple, dims: DecoderDims) -> float:
    """Uniform bound mimicking the reference initialisers (xavier for LinearNorm /
    ConvNorm, layers.py:13-15,34-35; 1/sqrt(H) for nn.LSTMCell; 1/sqrt(fan_in)
    for plain nn.Linear weights and biases)."""
    if "_rnn" in key:
        return 1.0 / math.sqrt(dims.arnn)
    if key.endswith("bias"):
        return 1.0 / math.sqrt(dims.drnn + dims.streams * dims.enc)
    if key.endswith("v.weight"):
        return 1.0 / math.sqrt(shape[1])
    gain = 5.0 / 3.0 if any(t in key for t in ("query_layer", "memory_layer", "location_dense")) else 1.0
    if len(shape) == 3:
        fan_in, fan_out = shape[1] * shape[2], shape[0] * shape[2]
    else:
        fan_out, fan_in = shape
    return gain * math.sqrt(6.0 / (fan_in + fan_out))


def make_decoder_weights(attention: str = SMA, seed: int = 1234,
                         dims: DecoderDims = DecoderDims(),
                         gate_bias: Optional[float] = None) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for key, shape in decoder_weight_shapes(attention, dims).items():
        b = _bound_for(key, shape, dims)
        out[key] = (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * b
    if gate_bias is not None:
        out["gate_layer.linear_layer.bias"].fill_(gate_bias)
    return out


def weights_checksum(w: Dict[str, torch.Tensor]) -> str:
    h = hashlib.sha256()
    for k in sorted(w):
        h.update(k.encode())
        h.update(w[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


def make_inputs(B: int, T_in: int, T_sub: int, T: int, seed: int = 1234, ragged: bool = False,
                dims: DecoderDims = DecoderDims()):
    """SURVEY.md 8d synthetic inputs: memory/embeddings 0.5*randn, mels randn.

    Returns dict(memory[B,T_in,enc], embeddings[B,T_sub,enc], mels[B,n_mel,T],
    memory_lengths[B], bert_lengths[B], output_lengths[B]) -- lengths int64.
    With ``ragged`` lengths are U{ceil(L/2)..L} with row 0 full length (mask
    width = max length, utils.py:10-14).
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    memory = 0.5 * torch.randn(B, T_in, dims.enc, generator=g)
    embeddings = 0.5 * torch.randn(B, max(T_sub, 1), dims.enc, generator=g)[:, :T_sub]
    mels = torch.randn(B, dims.n_mel, T, generator=g)

    def lens(L):
        if not ragged or B == 1:
            return torch.full((B,), L, dtype=torch.int64)
        lo = (L + 1) // 2
        v = torch.randint(lo, L + 1, (B,), generator=g, dtype=torch.int64)
        v[0] = L
        return v

    return dict(memory=memory, embeddings=embeddings, mels=mels,
                memory_lengths=lens(T_in), bert_lengths=lens(T_sub), output_lengths=lens(T))


@dataclass
class DropoutPlan:
    """Pre-drawn Bernoulli keep-masks and Gaussian noise, replayed into the
    reference, the oracle and the CUDA path alike (SURVEY.md 8c "Determinism").

    prenet_keep[s][l] : uint8 [T_p, B, prenet]   s = stream (0 char, 1 bert), l = layer.
                        Teacher-forced T_p = T+1 (model.py:412-413 runs the prenet
                        on the go-frame + all T targets); free-running T_p = max steps
                        (model.py:449-450, 470-471: frame t uses row t).
    lstm_keep         : uint8 [T, 6, B, H] or None.  Order per frame (model.py:341-346,
                        372-373): attn_h, attn_c, attn_h_bert, attn_c_bert, dec_h, dec_c.
    sma_noise[s]      : float32 [T, B, T_s] or None (attention.py:346-348).
    """
    prenet_keep: List[List[torch.Tensor]]
    lstm_keep: Optional[torch.Tensor] = None
    sma_noise: Optional[List[torch.Tensor]] = None


def make_dropout_plan(B: int, T_p: int, T: int, T_in: int, T_sub: int, training: bool,
                      seed: int = 4321, p_att: float = 0.1, p_dec: float = 0.1,
                      dims: DecoderDims = DecoderDims()) -> DropoutPlan:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    prenet_keep = [[(torch.rand(T_p, B, dims.prenet, generator=g) >= 0.5).to(torch.uint8)
                    for _l in range(2)] for _s in range(dims.streams)]
    lstm_keep = None
    noise = None
    if training:
        assert dims.arnn == dims.drnn
        r = torch.rand(T, 6, B, dims.arnn, generator=g)
        p = torch.tensor([p_att] * 4 + [p_dec] * 2).view(1, 6, 1, 1)
        lstm_keep = (r >= p).to(torch.uint8)
        noise = [torch.randn(T, B, L, generator=g) for L in ((T_in, T_sub)[:dims.streams])]
    return DropoutPlan(prenet_keep=prenet_keep, lstm_keep=lstm_keep, sma_noise=noise)


# ---------------------------------------------------------------------------------------------------
# Gradient parity plumbing shared by oracle/make_golden.py and tests/ (no reference import needed)
# ---------------------------------------------------------------------------------------------------
def seeded_loss(outs, seed: int):
    """sum_o <o, N(0,1)> over (mel, gate, align, align_bert) with a CPU generator -- identical for every implementation."""
    g = torch.Generator().manual_seed(seed)
    total = 0.0
    for o in outs:
        if o is None:
            continue
        total = total + (o * torch.randn(o.shape, generator=g).to(o.device)).sum()
    return total


def grad_digest(name: str, g: torch.Tensor) -> dict:
    g = g.detach().to("cpu", torch.float64).reshape(-1)
    r = torch.randn(g.numel(), generator=torch.Generator().manual_seed(len(name) * 7919 + g.numel()), dtype=torch.float64)
    return {name + "/max": np.array(float(g.abs().max())), name + "/sum": np.array(float(g.sum())),
            name + "/proj": np.array(float((g * r).sum() / np.sqrt(g.numel()))), name + "/head": g[:32].numpy().copy()}
