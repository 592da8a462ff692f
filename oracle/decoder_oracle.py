"""CPU oracle: a functional restatement of the reference Tacotron2 dual-stream decoder.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this file; the
product package never does and fails loudly when its CUDA library is missing.

PARITY PIN: the reference repository ships no golden vectors for this path
(SURVEY.md section 4 / 8c), so the oracle is pinned against *outputs of the reference
itself*: ``oracle/make_golden.py`` imports the unmodified ``/root/reference``
modules (through ``oracle/ref_shim.py``) in the build container, replays the same
dropout masks into them and commits the results under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against those fixtures, and (when
``/root/reference`` is present) ``tests/test_oracle_vs_reference.py`` re-runs the
reference live.

Third-party arithmetic restated here (not under /root/reference): PyTorch's
``nn.LSTMCell`` (gate order i,f,g,o; gates = W_ih x + b_ih + W_hh h + b_hh),
``nn.Linear``, ``nn.Conv1d``, ``F.dropout`` (keep * 1/(1-p)), ``F.softmax`` --
README.md:21 of the reference pins torch==1.8.0; the semantics are unchanged in
the torch 2.11 used here.

Every function cites the reference lines it follows.  The structure is deliberately
different from the reference (explicit state, explicit masks, no module
objects) -- it is a restatement, not a copy.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .synth import LSA, SMA, DecoderDims, DropoutPlan

Tensor = torch.Tensor


def lengths_to_mask(lengths: Tensor) -> Tensor:
    """utils.py:10-14 -- bool [B, max(lengths)], True on valid positions."""
    max_len = int(lengths.max().item())
    ids = torch.arange(max_len, dtype=torch.long)
    return ids.unsqueeze(0) < lengths.unsqueeze(1)


def _keep_scale(x: Tensor, keep: Tensor, p: float) -> Tensor:
    """F.dropout with an externally supplied keep-mask: x * keep * 1/(1-p)."""
    scale = torch.tensor(1.0 / (1.0 - p), dtype=x.dtype)
    return x * (keep.to(x.dtype) * scale)


def prenet(x: Tensor, w0: Tensor, w1: Tensor, keep0: Tensor, keep1: Tensor) -> Tensor:
    """model.py:13-24 -- two bias-free linears, ReLU, dropout p=0.5 ALWAYS on."""
    x = _keep_scale(F.relu(F.linear(x, w0)), keep0, 0.5)
    return _keep_scale(F.relu(F.linear(x, w1)), keep1, 0.5)


def lstm_cell(x: Tensor, h: Tensor, c: Tensor, w_ih: Tensor, w_hh: Tensor,
              b_ih: Tensor, b_hh: Tensor) -> Tuple[Tensor, Tensor]:
    """torch.nn.LSTMCell as called at model.py:340, 344, 371."""
    gates = F.linear(x, w_ih, b_ih) + F.linear(h, w_hh, b_hh)
    i, f, g, o = gates.chunk(4, dim=1)
    c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h_new = torch.sigmoid(o) * torch.tanh(c_new)
    return h_new, c_new


def sma_step(h: Tensor, memory: Tensor, pm: Tensor, alpha_prev: Tensor, wq: Tensor, v: Tensor,
             invalid: Optional[Tensor], noise: Optional[Tensor], truncate: bool = False):
    """attention.py:365-398 (forward), :354-363 (energies), :340-352 (p = sigmoid(e [+2*N(0,1)])),
    :330-338 (alpha'_j = alpha_j p_j + alpha_{j-1}(1 - p_{j-1}); mass past the end is dropped).

    ``invalid`` True = padded position (energy -> -inf before the sigmoid, attention.py:388-389).
    ``truncate`` (batched free-running extension, no reference equivalent): padded
    positions do not exist, i.e. alpha' is forced to 0 there so each utterance equals its
    own batch-1 reference run on unpadded memory.
    """
    q = F.linear(h, wq).unsqueeze(1)                                  # [B,1,A]
    e = F.linear(torch.tanh(q + pm), v).squeeze(-1)                   # [B,T]
    if invalid is not None:
        e = e.masked_fill(invalid, float("-inf"))
    if noise is not None:
        e = e + noise * 2.0                                            # sigmoid_noise = 2.0 (attention.py:316)
    p = torch.sigmoid(e)
    moved = alpha_prev[:, :-1] * (1.0 - p[:, :-1])
    alpha = alpha_prev * p + F.pad(moved, (1, 0))
    if truncate and invalid is not None:
        alpha = alpha.masked_fill(invalid, 0.0)
    ctx = torch.bmm(alpha.unsqueeze(1), memory).squeeze(1)
    return ctx, alpha


def lsa_step(h: Tensor, memory: Tensor, pm: Tensor, a_prev: Tensor, a_cum: Tensor, wq: Tensor,
             v: Tensor, wconv: Tensor, wdense: Tensor, invalid: Optional[Tensor]):
    """attention.py:64-85 (forward), :42-62 (energies), :19-23 (location layer):
    softmax_T( v . tanh(Wq h + dense(conv1d([a_prev; a_cum])) + pm) ) with -inf on padding."""
    q = F.linear(h, wq).unsqueeze(1)
    cat = torch.stack((a_prev, a_cum), dim=1)                          # model.py:351
    loc = F.conv1d(cat, wconv, padding=(wconv.shape[-1] - 1) // 2)     # [B,F,T]
    loc = F.linear(loc.transpose(1, 2), wdense)                        # [B,T,A]
    e = F.linear(torch.tanh(q + loc + pm), v).squeeze(-1)
    if invalid is not None:
        e = e.masked_fill(invalid, float("-inf"))
    alpha = F.softmax(e, dim=1)
    ctx = torch.bmm(alpha.unsqueeze(1), memory).squeeze(1)
    return ctx, alpha


@dataclass
class _Stream:
    memory: Tensor
    pm: Tensor            # processed memory (memory_layer applied once, model.py:258-261)
    invalid: Optional[Tensor]
    h: Tensor
    c: Tensor
    ctx: Tensor
    a_prev: Tensor        # decoder-side attention_weights (zeros at t=0, model.py:249)
    a_cum: Tensor
    sma_alpha: Tensor     # SMA module state, one-hot(0) at t=0 (attention.py:324-328)
    sfx: str


class DecoderOracle:
    """Explicit-state restatement of model.Decoder (model.py:128-492)."""

    def __init__(self, weights: Dict[str, Tensor], attention: str = SMA,
                 dims: DecoderDims = DecoderDims(), p_att: float = 0.1, p_dec: float = 0.1,
                 dtype: torch.dtype = torch.float32):
        self.w = {k: v.detach().to("cpu", dtype) for k, v in weights.items()}
        self.attention = attention
        self.d = dims
        self.p_att, self.p_dec = p_att, p_dec
        self.dtype = dtype
        self.sfx = ["", "_bert"][: dims.streams]

    # -- model.py:223-270 -------------------------------------------------------------
    def _init_stream(self, sfx: str, memory: Tensor, invalid: Optional[Tensor]) -> _Stream:
        B, T, _ = memory.shape
        z = lambda *s: torch.zeros(*s, dtype=self.dtype)
        pm = F.linear(memory, self.w[f"attention_layer{sfx}.memory_layer.linear_layer.weight"])
        onehot = z(B, T)
        onehot[:, 0] = 1.0
        return _Stream(memory=memory, pm=pm, invalid=invalid, h=z(B, self.d.arnn), c=z(B, self.d.arnn),
                       ctx=z(B, self.d.enc), a_prev=z(B, T), a_cum=z(B, T), sma_alpha=onehot, sfx=sfx)

    # -- model.py:322-390 -------------------------------------------------------------
    def _decode(self, streams: List[_Stream], pre: List[Tensor], h2: Tensor, c2: Tensor,
                lstm_keep: Optional[Tensor], noise: Optional[List[Tensor]], truncate: bool):
        w = self.w
        for i, (s, x) in enumerate(zip(streams, pre)):
            n = f"attention_rnn{s.sfx}."
            s.h, s.c = lstm_cell(torch.cat((x, s.ctx), -1), s.h, s.c, w[n + "weight_ih"],
                                 w[n + "weight_hh"], w[n + "bias_ih"], w[n + "bias_hh"])   # :337-344
            if lstm_keep is not None:                                                     # :341-346
                s.h = _keep_scale(s.h, lstm_keep[2 * i], self.p_att)
                s.c = _keep_scale(s.c, lstm_keep[2 * i + 1], self.p_att)
        for i, s in enumerate(streams):                                                   # :351-359
            a = f"attention_layer{s.sfx}."
            if self.attention == SMA:
                s.ctx, alpha = sma_step(s.h, s.memory, s.pm, s.sma_alpha,
                                        w[a + "query_layer.linear_layer.weight"], w[a + "v.weight"],
                                        s.invalid, None if noise is None else noise[i], truncate)
                s.sma_alpha = alpha
            else:
                s.ctx, alpha = lsa_step(s.h, s.memory, s.pm, s.a_prev, s.a_cum,
                                        w[a + "query_layer.linear_layer.weight"],
                                        w[a + "v.linear_layer.weight"],
                                        w[a + "location_layer.location_conv.conv.weight"],
                                        w[a + "location_layer.location_dense.linear_layer.weight"],
                                        s.invalid)
            s.a_prev = alpha
            s.a_cum = s.a_cum + alpha
        x2 = torch.cat([t for s in streams for t in (s.h, s.ctx)], -1)                   # :362
        h2, c2 = lstm_cell(x2, h2, c2, w["decoder_rnn.weight_ih"], w["decoder_rnn.weight_hh"],
                           w["decoder_rnn.bias_ih"], w["decoder_rnn.bias_hh"])            # :371
        if lstm_keep is not None:                                                         # :372-373
            h2 = _keep_scale(h2, lstm_keep[4], self.p_dec)
            c2 = _keep_scale(c2, lstm_keep[5], self.p_dec)
        y = torch.cat([h2] + [s.ctx for s in streams], dim=1)                             # :382
        mel = F.linear(y, w["linear_projection.linear_layer.weight"],
                       w["linear_projection.linear_layer.bias"])                          # :385
        gate = F.linear(y, w["gate_layer.linear_layer.weight"], w["gate_layer.linear_layer.bias"])  # :388
        return mel, gate, h2, c2

    def _prenet(self, sfx: str, x: Tensor, keep0: Tensor, keep1: Tensor) -> Tensor:
        return prenet(x, self.w[f"prenet{sfx}.layers.0.linear_layer.weight"],
                      self.w[f"prenet{sfx}.layers.1.linear_layer.weight"], keep0, keep1)

    # -- model.py:392-428 -------------------------------------------------------------
    def forward(self, memory: Tensor, embeddings: Optional[Tensor], decoder_inputs: Tensor,
                memory_lengths: Tensor, bert_lengths: Optional[Tensor], plan: DropoutPlan,
                training: bool = False):
        """Teacher-forced.  Returns mel [B,n_mel,T], gate [B,T], align [B,T,T_in], align_bert [B,T,T_sub]."""
        dt = self.dtype
        mems = [memory.to(dt)] + ([embeddings.to(dt)] if self.d.streams == 2 else [])
        lens = [memory_lengths] + ([bert_lengths] if self.d.streams == 2 else [])
        B, _, T = decoder_inputs.shape
        frames = decoder_inputs.to(dt).permute(2, 0, 1)                                   # :283-287
        frames = torch.cat((torch.zeros(1, B, self.d.n_mel, dtype=dt), frames), 0)        # :407-411
        pre = [self._prenet(s, frames, plan.prenet_keep[i][0], plan.prenet_keep[i][1])
               for i, s in enumerate(self.sfx)]                                           # :412-413
        streams = []
        for s, m, l in zip(self.sfx, mems, lens):
            valid = lengths_to_mask(l)                                                    # :414
            assert valid.shape[1] == m.shape[1], "memory width must equal max(lengths) (utils.py:11)"
            streams.append(self._init_stream(s, m, ~valid))
        h2 = torch.zeros(B, self.d.drnn, dtype=dt)
        c2 = torch.zeros(B, self.d.drnn, dtype=dt)
        mels, gates, aligns = [], [], [[] for _ in streams]
        for t in range(T):                                                                # :417-424
            keep = plan.lstm_keep[t] if (training and plan.lstm_keep is not None) else None
            noise = [n[t].to(dt) for n in plan.sma_noise] if (training and plan.sma_noise is not None) else None
            mel, gate, h2, c2 = self._decode(streams, [p[t] for p in pre], h2, c2, keep, noise, False)
            mels.append(mel)
            gates.append(gate.squeeze(1))
            for a, s in zip(aligns, streams):
                a.append(s.a_prev)
        mel_out = torch.stack(mels).permute(1, 2, 0).contiguous()                         # :290-320
        gate_out = torch.stack(gates).transpose(0, 1).contiguous()
        al = [torch.stack(a).transpose(0, 1).contiguous() for a in aligns]
        return (mel_out, gate_out, al[0], al[1] if len(al) > 1 else None)

    # -- model.py:430-492 -------------------------------------------------------------
    def inference(self, memory: Tensor, embeddings: Optional[Tensor], plan: DropoutPlan,
                  max_decoder_steps: int = 1000, gate_threshold: float = 0.001,
                  plan_batch_index: int = 0):
        """Free-running, batch 1 (the reference is batch-1 only, model.py:461,480).
        Returns mel [1,n_mel,T], gate [1,T,1], align, align_bert, INFER_FLAG."""
        dt = self.dtype
        assert memory.shape[0] == 1
        mems = [memory.to(dt)] + ([embeddings.to(dt)] if self.d.streams == 2 else [])
        streams = [self._init_stream(s, m, None) for s, m in zip(self.sfx, mems)]         # :446
        h2 = torch.zeros(1, self.d.drnn, dtype=dt)
        c2 = torch.zeros(1, self.d.drnn, dtype=dt)
        x = torch.zeros(1, self.d.n_mel, dtype=dt)                                        # :444-445
        mels, gates, aligns = [], [], [[] for _ in streams]
        flag = True
        b = plan_batch_index
        thr = torch.tensor(gate_threshold, dtype=torch.float32)
        while True:
            t = len(mels)
            pre = [self._prenet(s, x, plan.prenet_keep[i][0][t, b:b + 1], plan.prenet_keep[i][1][t, b:b + 1])
                   for i, s in enumerate(self.sfx)]                                       # :449-450, 470-471
            mel, gate, h2, c2 = self._decode(streams, pre, h2, c2, None, None, False)
            mels.append(mel)
            gates.append(gate)
            for a, s in zip(aligns, streams):
                a.append(s.a_prev)
            if bool(torch.sigmoid(gate.to(torch.float32)) > thr):                         # :461, 480 (strict >)
                break
            if len(mels) == max_decoder_steps:                                            # :463, 482
                flag = False
                break
            x = mel                                                                       # :459-460, 487-488
        mel_out = torch.stack(mels).permute(1, 2, 0).contiguous()
        gate_out = torch.stack(gates).transpose(0, 1).contiguous()                        # [1,T,1]
        al = [torch.stack(a).transpose(0, 1).contiguous() for a in aligns]
        return mel_out, gate_out, al[0], (al[1] if len(al) > 1 else None), flag

    def inference_batched(self, memory: Tensor, embeddings: Optional[Tensor], memory_lengths: Tensor,
                          bert_lengths: Optional[Tensor], plan: DropoutPlan, max_decoder_steps: int = 1000,
                          gate_threshold: float = 0.001):
        """Definition of batched free-running decoding: utterance b's result
        equals the batch-1 reference run on its own un-padded memory with its own masks."""
        outs = []
        for b in range(memory.shape[0]):
            m = memory[b:b + 1, : int(memory_lengths[b])]
            e = None
            if self.d.streams == 2:
                e = embeddings[b:b + 1, : int(bert_lengths[b])]
            outs.append(self.inference(m, e, plan, max_decoder_steps, gate_threshold, plan_batch_index=b))
        return outs


def output_padding_mask(output_lengths: Tensor, n_mel: int) -> Tensor:
    """model.py:531-541 -- True beyond each utterance's output length, [B,n_mel,T]."""
    invalid = ~lengths_to_mask(output_lengths)
    return invalid.unsqueeze(1).expand(-1, n_mel, -1)


def apply_output_mask(mel: Tensor, mel_post: Optional[Tensor], gate: Tensor, output_lengths: Tensor):
    """model.py:537-539 -- mel(s) <- 0.0, gate <- 1e3 beyond output_lengths."""
    m = output_padding_mask(output_lengths, mel.shape[1])
    mel = mel.masked_fill(m, 0.0)
    if mel_post is not None:
        mel_post = mel_post.masked_fill(m, 0.0)
    gate = gate.masked_fill(m[:, 0, :], 1e3)
    return mel, mel_post, gate
