"""Drop-in model classes: same constructor arguments, method signatures, hparams fields and
state_dict layout as /root/reference/model.py -- with the decoder's per-frame loop replaced
by the persistent sm_100a kernel behind the C ABI (include/taco2dec.h).

  Decoder.forward    <- model.py:392-428      Decoder.inference  <- model.py:430-492
  BERT_Tacotron2     <- model.py:494-582      Tacotron2 (compat) <- GTA.py:21,57-59 / inference.py:302,334

Encoder / Postnet / embeddings are plain PyTorch exactly as in the reference (out of scope for
the kernel work); they exist here so the wrapper classes are usable as a drop-in.
There is no CPU decoder: calling the decoder without the built library or a B200 raises.
"""
from __future__ import annotations

import ctypes as C
from math import sqrt
from typing import Optional

import torch
from torch import nn
from torch.nn import functional as F

from . import _cabi
from .attention import LSA, SMA, LocationSensitiveAttention, StepwiseMonotonicAttention
from .layers import ConvNorm, LinearNorm
from .utils import get_mask_from_lengths, to_gpu


class Prenet(nn.Module):
    """Weights of the 2-layer prenet (model.py:13-24).  ``forward`` is the plain PyTorch
    definition (dropout p=0.5 always on) for callers that use the module stand-alone; the
    decoder evaluates it inside the CUDA path."""

    def __init__(self, in_dim, sizes):
        super().__init__()
        dims = [in_dim] + list(sizes)
        self.layers = nn.ModuleList(LinearNorm(a, b, bias=False) for a, b in zip(dims[:-1], dims[1:]))

    def forward(self, x):
        for lin in self.layers:
            x = F.dropout(F.relu(lin(x)), p=0.5, training=True)
        return x


class DropoutReplay:
    """Externally drawn masks for parity runs (see include/taco2dec.h ``taco2dec_rng``).
    prenet_keep[s][l]: uint8 [rows,B,prenet]; lstm_keep: uint8 [T,6,B,H]; sma_noise[s]: f32 [T,B,T_s]."""

    def __init__(self, prenet_keep=None, lstm_keep=None, sma_noise=None):
        self.prenet_keep, self.lstm_keep, self.sma_noise = prenet_keep, lstm_keep, sma_noise

    def to(self, device):
        mv = lambda t, dt: None if t is None else t.to(device=device, dtype=dt).contiguous()
        pk = None if self.prenet_keep is None else [[mv(m, torch.uint8) for m in row] for row in self.prenet_keep]
        nz = None if self.sma_noise is None else [mv(n, torch.float32) for n in self.sma_noise]
        return DropoutReplay(pk, mv(self.lstm_keep, torch.uint8), nz)


class _Engine:
    """One C-ABI handle per (Decoder, device) + a reusable workspace."""

    def __init__(self, cfg: _cabi.Config, device: torch.device):
        self.lib = _cabi.load_library()
        self.device = device
        self.handle = C.c_void_p()
        _cabi.check(self.lib.taco2dec_create(C.byref(cfg), device.index or 0, C.byref(self.handle)))
        self.workspace: Optional[torch.Tensor] = None
        self.weights_key = None   # (data_ptr, version) of every bound tensor; re-bind only on change
        self.mode = None

    def get_workspace(self, B, T_in, T_sub, T, tf) -> torch.Tensor:
        need = int(self.lib.taco2dec_workspace_bytes(self.handle, B, T_in, T_sub, T, int(tf)))
        if need == 0:
            raise _cabi.Taco2DecError("invalid shape for workspace query")
        if self.workspace is None or self.workspace.numel() < need:
            self.workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self.workspace

    def launch_count(self) -> int:
        return int(self.lib.taco2dec_launch_count(self.handle))

    def set_mode(self, path: str, weight_dtype: str, batched_precision: str = "fp16") -> None:
        paths = {"auto": _cabi.PATH_AUTO, "generic": _cabi.PATH_GENERIC, "latency": _cabi.PATH_LATENCY,
                 "tensor": _cabi.PATH_TENSOR, "tensor_graph": _cabi.PATH_TENSOR_GRAPH}
        dts = {"fp32": _cabi.W_FP32, "fp16": _cabi.W_FP16}
        bps = {"fp16": 0, "fp32": 1}
        if (path, weight_dtype, batched_precision) != self.mode:
            _cabi.check(self.lib.taco2dec_set_mode(self.handle, paths[path], dts[weight_dtype]))
            _cabi.check(self.lib.taco2dec_set_batched_precision(self.handle, bps[batched_precision]))
            self.mode = (path, weight_dtype, batched_precision)
            self.weights_key = None   # packed streams depend on the storage type

    def last_path(self) -> str:
        return {0: "none", 1: "generic", 2: "latency", 3: "tensor", 4: "tensor_graph"}[
            int(self.lib.taco2dec_last_path(self.handle))]

    def set_profiling(self, on: bool) -> None:
        _cabi.check(self.lib.taco2dec_set_profiling(self.handle, int(on)))

    def phase_clocks(self):
        buf = (C.c_longlong * 16)()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _cabi.check(self.lib.taco2dec_read_phase_clocks(self.handle, C.c_void_p(stream), buf))
        return list(buf)

    def measure_machine(self):
        """(L2 -> SM read rate in GB/s, latency of one cross-CTA exchange through L2 in ns) measured on this device."""
        bw, hop = C.c_double(), C.c_double()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _cabi.check(self.lib.taco2dec_measure_machine(self.handle, C.c_void_p(stream), C.byref(bw), C.byref(hop)))
        return float(bw.value), float(hop.value)

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        _cabi.check(self.lib.taco2dec_last_kernel_ms(self.handle, C.byref(ms)))
        return float(ms.value)

    def __del__(self):
        try:
            if self.handle:
                self.lib.taco2dec_destroy(self.handle)
        except Exception:
            pass


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class _NoBackward(torch.autograd.Function):
    """Marks decoder outputs as differentiable so training code fails LOUDLY at backward()
    instead of silently skipping the decoder when the shape has no backward kernel (the BPTT
    kernels cover the tensor path: 2 <= B <= 128, default dims)."""

    @staticmethod
    def forward(ctx, anchor, *outs):
        return tuple(o.view_as(o) for o in outs)

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError(
            "tacotron2_subword_b200: decoder backward is implemented for the tensor path only "
            "(2 <= batch <= 128, default decoder dims)")


def _fview(buf: torch.Tensor, off: int, *shape) -> torch.Tensor:
    n = 1
    for d in shape:
        n *= d
    return buf[off: off + 4 * n].view(torch.float32).view(*shape)


class _DecoderTF(torch.autograd.Function):
    """Teacher-forced decoder pass with a hand-written backward (the reference relies on autograd over
    its per-frame Python loop, model.py:392-428 + train.py:245-256).  The reverse-time recurrence runs in
    ``taco2dec_backward``; the sums over (frame, utterance) that remain -- weight gradients, d memory, the
    hoisted prenet -- are plain GEMMs over the rows it leaves behind (see ``taco2dec_grad_layout``)."""

    @staticmethod
    def forward(ctx, dec, independent, memory, embeddings, dec_in, mlen, blen, *params):
        ctx.set_materialize_grads(False)
        outs, state = dec._run_tf(memory, embeddings, dec_in, mlen, blen, save=True, independent=independent)
        mel, gate, align, align_b = outs
        # The Function's own outputs must not be reachable from ctx (output -> grad_fn -> ctx -> output is a cycle the
        # reference counter cannot free, and it would pin the multi-GB `saved` buffer): keep detached aliases instead.
        state["align"] = align.detach()
        state["align_b"] = None if align_b is None else align_b.detach()
        ctx.dec, ctx.state = dec, state
        ctx.n_params = len(params)
        return (mel, gate, align) + ((align_b,) if align_b is not None else ())

    @staticmethod
    def backward(ctx, d_mel, d_gate, d_align=None, d_align_b=None):
        dec, st = ctx.dec, ctx.state
        if st is None:
            raise RuntimeError("decoder backward called twice: the saved activations are released after the first pass "
                               "(retain_graph is not supported by the hand-written BPTT)")
        ctx.state = None          # release the saved-activation buffer as soon as this pass is done
        eng, dev = st["eng"], st["dev"]
        dec.check(dev)            # a watchdog abort in the forward pass must not turn into silent garbage gradients
        B, T, T_in, T_sub = st["B"], st["T"], st["T_in"], st["T_sub"]
        S, H, E, P, A, M = dec.n_streams, dec.attention_rnn_dim, dec.encoder_embedding_dim, dec.prenet_dim, \
            dec.attention_dim, dec.n_mel_channels
        G = 4 * H
        z = lambda g, *shape: torch.zeros(*shape, device=dev) if g is None else g.to(torch.float32).contiguous()
        d_mel, d_gate = z(d_mel, B, T, M), z(d_gate, B, T)
        d_align = None if d_align is None else d_align.contiguous()
        d_align_b = None if d_align_b is None else d_align_b.contiguous()
        GL = _cabi.GradLayout()
        _cabi.check(eng.lib.taco2dec_grad_layout_query(eng.handle, B, T_in, T_sub, T, C.byref(GL)))
        gbuf = torch.empty(int(GL.total), dtype=torch.uint8, device=dev)
        a = _cabi.BwdArgs()
        a.B, a.T, a.T_in, a.T_sub = B, T, T_in, T_sub
        a.memory, a.embeddings = _ptr(st["mem"]), _ptr(st["emb"])
        a.memory_lengths, a.bert_lengths = _ptr(st["mlen"]), _ptr(st["blen"])
        a.training = st["training"]
        a.independent = st["independent"]
        a.rng = st["rng"]
        a.align, a.align_bert = _ptr(st["align"]), _ptr(st["align_b"])
        a.d_mel, a.d_gate, a.d_align, a.d_align_bert = _ptr(d_mel), _ptr(d_gate), _ptr(d_align), _ptr(d_align_b)
        a.saved, a.saved_bytes = _ptr(st["saved"]), st["saved"].numel()
        a.grads, a.grads_bytes = _ptr(gbuf), gbuf.numel()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            _cabi.check(eng.lib.taco2dec_backward(eng.handle, C.byref(a), C.c_void_p(stream.cuda_stream)))

        # ---- contractions over (frame, utterance) on the repo's own tcgen05 GEMM (taco2dec_wgrad_gemm): Y^T . X with K = T*B;
        #      the per-frame gradient rows and the saved activations are addressed in place through (frame, utterance) strides ----
        SL, sv = st["SL"], st["saved"]
        h1 = _fview(sv, SL.h1, T + 1, S, B, H)
        cx = _fview(sv, SL.ctx, T + 1, S, B, E)
        h2 = _fview(sv, SL.h2, T + 1, B, H)
        dg1 = _fview(gbuf, GL.dg1, S, T, B, G)
        dg2 = _fview(gbuf, GL.dg2, T, B, G)
        dq = _fview(gbuf, GL.dq, S, T, B, A)
        dctx = _fview(gbuf, GL.dctx, S, T, B, E)
        dpre = _fview(gbuf, GL.dpre, S, T, B, P)
        dv = _fview(gbuf, GL.dv, S, B, A)
        Ts = [T_in, T_sub]
        mems = [st["mem"], st["emb"]]
        aligns = [st["align"], st["align_b"]]
        x_frames = F.pad(st["dec_in"].permute(2, 0, 1), (0, 0, 0, 0, 1, 0))[:T].contiguous()     # go frame + targets [T, B, M]
        grads = {}
        d_mems = [None, None]
        # ``dec._grad_ready`` (set by distributed.GradientBucketer): called with (parameter, gradient) the moment a weight
        # gradient exists, largest tensors first, so that its all-reduce over NCCL runs while the remaining contractions
        # are still being computed (the reference reduces everything after the whole backward, distributed.py:160-167)
        publish = getattr(dec, "_grad_ready", None)
        lib, hnd, cstream = eng.lib, eng.handle, C.c_void_p(stream.cuda_stream)
        ws_box = [None, None]       # workspace, key of the Y operand whose packed image it holds

        def wgrad(Y3, X3, out, accumulate=False):
            """out[M, N] (a 2-D view, unit column stride) (+)= sum over (t, b) of Y3[t, b, :]^T X3[t, b, :]."""
            Tn, Bn, Mn = Y3.shape
            Nn = X3.shape[2]
            assert X3.shape[:2] == (Tn, Bn) and Y3.stride(2) == 1 and X3.stride(2) == 1 and out.stride(1) == 1
            need = int(lib.taco2dec_wgrad_workspace_bytes(hnd, Mn, Nn, Tn, Bn))
            if ws_box[0] is None or ws_box[0].numel() < need:
                ws_box[0], ws_box[1] = torch.empty(max(need, ws_need), dtype=torch.uint8, device=dev), None
            ykey = (Y3.data_ptr(), tuple(Y3.shape), tuple(Y3.stride()))
            reuse = ws_box[1] == ykey          # same gradient rows as the previous product: its packed operand is still there
            ws_box[1] = ykey
            _cabi.check(lib.taco2dec_wgrad_gemm(hnd, _ptr(Y3), Y3.stride(0), Y3.stride(1), Mn, _ptr(X3), X3.stride(0), X3.stride(1), Nn,
                                                Tn, Bn, _ptr(out), out.stride(0), int(accumulate), int(reuse), _ptr(ws_box[0]),
                                                ws_box[0].numel(), cstream))

        def put(param, g):
            grads[param] = g
            if publish is not None and param.requires_grad:
                publish(param, g)

        # one workspace for every product: sized for the largest (gate rows x the widest activation block)
        ws_need = int(lib.taco2dec_wgrad_workspace_bytes(hnd, G, H, T, B))
        with torch.cuda.device(dev):
            # decoder LSTM: dWd_ih = dg2^T [h1_0 | ctx_0 | h1_1 | ctx_1](t), dWd_hh = dg2^T h2(t-1)
            g_ih = torch.empty(G, S * (H + E), device=dev)
            for s in range(S):
                wgrad(dg2, h1[1:, s], g_ih[:, s * (H + E): s * (H + E) + H])
                wgrad(dg2, cx[1:, s], g_ih[:, s * (H + E) + H: (s + 1) * (H + E)])
            put(dec.decoder_rnn.weight_ih, g_ih)
            g_hh = torch.empty(G, H, device=dev)
            wgrad(dg2, h2[:T], g_hh)
            put(dec.decoder_rnn.weight_hh, g_hh)
            db = dg2.sum((0, 1))
            put(dec.decoder_rnn.bias_ih, db)
            put(dec.decoder_rnn.bias_hh, db.clone())
            for s, sfx in enumerate(dec._sfx()):
                pre_m, rnn, att = getattr(dec, "prenet" + sfx), getattr(dec, "attention_rnn" + sfx), \
                    getattr(dec, "attention_layer" + sfx)
                pre = _fview(sv, SL.pre[s], T + 1, B, P)[:T]
                pre0 = _fview(sv, SL.pre0[s], T + 1, B, P)[:T]
                # attention LSTM: dW_hh = dg1^T h1(t-1), dW_ih = dg1^T [prenet(t) | ctx(t-1)]
                g_hh = torch.empty(G, H, device=dev)
                wgrad(dg1[s], h1[:T, s], g_hh)
                put(rnn.weight_hh, g_hh)
                g_ih = torch.empty(G, P + E, device=dev)
                wgrad(dg1[s], pre, g_ih[:, :P])
                wgrad(dg1[s], cx[:T, s], g_ih[:, P:])
                put(rnn.weight_ih, g_ih)
                db = dg1[s].sum((0, 1))
                put(rnn.bias_ih, db)
                put(rnn.bias_hh, db.clone())
                g_q = torch.empty(A, H, device=dev)
                wgrad(dq[s], h1[1:, s], g_q)
                put(att.query_layer.linear_layer.weight, g_q)
                put(att.v_weight(), dv[s].sum(0).view(1, A))
                if dec.attention_kind == LSA:
                    LF, LK = dec.loc_filters, dec.loc_kernel
                    put(att.location_layer.location_dense.linear_layer.weight, _fview(gbuf, GL.dloc_dense, S, B, A, LF)[s].sum(0))
                    put(att.location_layer.location_conv.conv.weight, _fview(gbuf, GL.dloc_conv, S, B, LF, 2, LK)[s].sum(0))
                dpm = _fview(gbuf, GL.dpm[s], B, Ts[s], A)
                wm = att.memory_layer.linear_layer.weight
                g_m = torch.empty(A, E, device=dev)
                wgrad(dpm.reshape(B * Ts[s], 1, A), mems[s].reshape(B * Ts[s], 1, E), g_m)       # rows = (utterance, position)
                put(wm, g_m)
                # d memory[b] = alignments[b]^T . d ctx[:, b] + d processed_memory[b] . Wm      (attention.py:395, model.py:258-261)
                dm_s = torch.empty(B, Ts[s], E, device=dev)
                al = aligns[s]
                _cabi.check(lib.taco2dec_bmm_tn(_ptr(al), al.stride(0), al.stride(1), _ptr(dctx[s]), dctx[s].stride(0), dctx[s].stride(1),
                                                _ptr(dm_s), dm_s.stride(0), dm_s.stride(1), B, Ts[s], E, T, cstream))
                wmd = wm.detach()
                _cabi.check(lib.taco2dec_sgemm_nn(_ptr(dpm), A, _ptr(wmd), wmd.stride(0), _ptr(dm_s), E, B * Ts[s], E, A, None, 0, 1, cstream))
                d_mems[s] = dm_s
                # prenet: two bias-free linear + ReLU + always-on dropout layers (model.py:13-24); kept <=> output > 0
                w0, w1 = pre_m.layers[0].linear_layer.weight, pre_m.layers[1].linear_layer.weight
                dz1 = dpre[s] * 2.0 * (pre > 0)
                g_w1 = torch.empty(P, P, device=dev)
                wgrad(dz1, pre0, g_w1)
                put(w1, g_w1)
                dz0 = torch.empty(T, B, P, device=dev)
                w1d = w1.detach()
                _cabi.check(lib.taco2dec_sgemm_nn(_ptr(dz1), P, _ptr(w1d), w1d.stride(0), _ptr(dz0), P, T * B, P, P, _ptr(pre0), P, 0, cstream))
                g_w0 = torch.empty(P, M, device=dev)
                wgrad(dz0, x_frames, g_w0)
                put(w0, g_w0)
            # projection / gate: d[Wp ; wg] = [d mel | d gate]^T [h2(t) | ctx_0(t) | ctx_1(t)]
            KD = H + S * E
            g_p = torch.empty(M, KD, device=dev)
            g_g = torch.empty(1, KD, device=dev)
            dm3 = d_mel.transpose(0, 1)                      # [T, B, M] view of the [B, T, M] gradient
            dg3 = d_gate.t().unsqueeze(2)                    # [T, B, 1]
            for Y3, out in ((dm3, g_p), (dg3, g_g)):
                wgrad(Y3, h2[1:], out[:, :H])
                for s in range(S):
                    wgrad(Y3, cx[1:, s], out[:, H + s * E: H + (s + 1) * E])
            put(dec.linear_projection.linear_layer.weight, g_p)
            put(dec.linear_projection.linear_layer.bias, d_mel.sum((0, 1)))
            put(dec.gate_layer.linear_layer.weight, g_g)
            put(dec.gate_layer.linear_layer.bias, d_gate.sum().view(1))
        finish = getattr(dec, "_grad_ready_finish", None)
        if finish is not None:
            finish()          # the early all-reduces must have landed before autograd hands these tensors on
        param_grads = tuple(grads.get(p_) if p_.requires_grad else None for p_ in st["params"])
        need = ctx.needs_input_grad
        return (None, None, d_mems[0] if need[2] else None, d_mems[1] if (S == 2 and need[3]) else None, None, None, None) \
            + param_grads


class Decoder(nn.Module):
    """model.py:128-492.  Same parameters / state_dict; the loop runs on the GPU."""

    def __init__(self, hparams, n_streams: int = 2):
        super().__init__()
        hp = hparams
        self.n_streams = n_streams
        self.n_mel_channels = hp.n_mel_channels
        self.n_frames_per_step = hp.n_frames_per_step
        if self.n_frames_per_step != 1:
            raise ValueError("only n_frames_per_step == 1 is supported (as in the reference, hparams.py:76)")
        self.encoder_embedding_dim = hp.encoder_embedding_dim
        self.attention_rnn_dim = hp.attention_rnn_dim
        self.decoder_rnn_dim = hp.decoder_rnn_dim
        self.prenet_dim = hp.prenet_dim
        self.max_decoder_steps = hp.max_decoder_steps
        self.gate_threshold = hp.gate_threshold
        self.p_attention_dropout = hp.p_attention_dropout
        self.p_decoder_dropout = hp.p_decoder_dropout
        self.attention_dim = hp.attention_dim
        self.loc_filters = hp.attention_location_n_filters
        self.loc_kernel = hp.attention_location_kernel_size
        self.attention_kind = SMA if hp.attention == SMA else LSA
        if hp.attention not in (SMA, LSA):
            # every other choice crashes inside the reference's decode()
            raise ValueError(f"attention '{hp.attention}' is not runnable in the reference either; use SMA or LSA")
        att_cls = StepwiseMonotonicAttention if self.attention_kind == SMA else LocationSensitiveAttention
        sfx = ["", "_bert"][:n_streams]
        mel_in = hp.n_mel_channels * hp.n_frames_per_step
        for s in sfx:
            setattr(self, "prenet" + s, Prenet(mel_in, [hp.prenet_dim, hp.prenet_dim]))
        for s in sfx:
            setattr(self, "attention_rnn" + s,
                    nn.LSTMCell(hp.prenet_dim + hp.encoder_embedding_dim, hp.attention_rnn_dim))
        for s in sfx:
            # NB: the reference only builds attention_layer_bert for SMA (model.py:158-191) and
            # crashes otherwise; building it for LSA too is the fix SURVEY.md 0.4 prescribes.
            setattr(self, "attention_layer" + s,
                    att_cls(hp.attention_rnn_dim, hp.encoder_embedding_dim, hp.attention_dim,
                            hp.attention_location_n_filters, hp.attention_location_kernel_size))
        self.decoder_rnn = nn.LSTMCell(n_streams * (hp.attention_rnn_dim + hp.encoder_embedding_dim),
                                       hp.decoder_rnn_dim)
        if n_streams == 2:
            # dead in decode() (model.py:375-378) but part of the checkpoint layout
            self.decoder_rnn_bert = nn.LSTMCell(hp.attention_rnn_dim + hp.encoder_embedding_dim, hp.decoder_rnn_dim)
        proj_in = hp.decoder_rnn_dim + n_streams * hp.encoder_embedding_dim
        self.linear_projection = LinearNorm(proj_in, mel_in)
        self.gate_layer = LinearNorm(proj_in, 1, bias=True, w_init_gain="sigmoid")
        # -- extensions (not in the reference) --------------------------------------------
        # "auto": batch 1 -> latency path; 2 <= B <= 128 -> tensor path unless batched_precision == "fp32"; else generic
        self.decoder_path = "auto"      # "auto" | "generic" | "latency" | "tensor" | "tensor_graph"  (include/taco2dec.h)
        self.weight_dtype = "fp32"      # storage of the packed LSTM matrices on the LATENCY path (batch 1): "fp32" | "fp16"
        # Batched calls (2 <= B <= 128): "fp16" = tcgen05 path, LSTM / query operands (weights AND x, h) rounded to fp16,
        # fp32 accumulation, stated bound mel/gate <= 1e-3, alignments <= 2e-4 vs the fp32 reference; "fp32" = the generic
        # fp32-exact kernel (~15x slower at B = 64; no backward: training with it raises).
        self.batched_precision = "fp16"
        self.max_backward_rows = 128       # rows one BPTT call takes (tcgen05 N limit); larger batches run as sub-batches
        self.dropout_replay: Optional[DropoutReplay] = None  # parity runs: externally drawn masks
        self.rng_seed: Optional[int] = None                  # fixed Philox seed; None = fresh per call
        self.validate_lengths = True
        self._pm_hint = None            # (memory ptr, processed memory) pairs from MemoryPrep, consumed by the next call
        self.grad_gemm_tf32: Optional[bool] = None   # weight-gradient GEMMs in TF32; None = only with weight_dtype "fp16"
        self._engines = {}

    # ------------------------------------------------------------------------------------
    def _sfx(self):
        return ["", "_bert"][: self.n_streams]

    def _engine(self, device: torch.device) -> _Engine:
        if device.type != "cuda":
            raise _cabi.Taco2DecError(
                "the decoder runs only on a CUDA sm_100 device (no CPU fallback); move the module with .cuda()")
        key = device.index if device.index is not None else torch.cuda.current_device()
        eng = self._engines.get(key)
        if eng is None:
            cfg = _cabi.Config(self.n_mel_channels, self.encoder_embedding_dim, self.attention_rnn_dim,
                               self.decoder_rnn_dim, self.prenet_dim, self.attention_dim, self.loc_filters,
                               self.loc_kernel, _cabi.ATTN_SMA if self.attention_kind == SMA else _cabi.ATTN_LSA,
                               self.n_streams, float(self.p_attention_dropout), float(self.p_decoder_dropout))
            eng = _Engine(cfg, torch.device("cuda", key))
            self._engines[key] = eng
        return eng

    @staticmethod
    def _w(t: torch.Tensor) -> torch.Tensor:
        if t.dtype != torch.float32 or not t.is_contiguous() or t.data_ptr() % 16:
            raise _cabi.Taco2DecError("decoder parameters must be contiguous, 16-byte aligned fp32")
        return t

    def _weight_tensors(self):
        ts = []
        for s in self._sfx():
            pre, rnn, att = getattr(self, "prenet" + s), getattr(self, "attention_rnn" + s), getattr(self, "attention_layer" + s)
            ts += [pre.layers[0].linear_layer.weight, pre.layers[1].linear_layer.weight, rnn.weight_ih, rnn.weight_hh,
                   rnn.bias_ih, rnn.bias_hh, att.query_layer.linear_layer.weight, att.memory_layer.linear_layer.weight,
                   att.v_weight()]
            if self.attention_kind == LSA:
                ts += [att.location_layer.location_conv.conv.weight, att.location_layer.location_dense.linear_layer.weight]
        ts += [self.decoder_rnn.weight_ih, self.decoder_rnn.weight_hh, self.decoder_rnn.bias_ih, self.decoder_rnn.bias_hh,
               self.linear_projection.linear_layer.weight, self.linear_projection.linear_layer.bias,
               self.gate_layer.linear_layer.weight, self.gate_layer.linear_layer.bias]
        return ts

    def set_processed_memory(self, memory, pm, embeddings=None, pm_bert=None) -> None:
        """Hand the next decoder call the processed memory that ``MemoryPrep`` already computed for exactly these tensors
        (``memory_layer(memory)``, model.py:258-261); ignored unless the call receives the same storage."""
        self._pm_hint = [(memory.data_ptr(), tuple(memory.shape[:2]), pm),
                         None if embeddings is None else (embeddings.data_ptr(), tuple(embeddings.shape[:2]), pm_bert)]

    def _take_pm_hint(self, mem, emb):
        hint, self._pm_hint = self._pm_hint, None
        out = [None, None]
        if hint is None:
            return out
        for i, t in enumerate((mem, emb)):
            h = hint[i] if i < len(hint) else None
            if t is not None and h is not None and h[0] == t.data_ptr() and h[1] == tuple(t.shape[:2]) and h[2].is_contiguous():
                out[i] = h[2]
        return out

    def invalidate_weights(self) -> None:
        """Force the library-owned weight re-layouts (latency-path streams, fp16 / bf16 tiles) to be rebuilt on the next
        call.  Needed after in-place updates that autograd does not see (``p.data.add_()``, ``p.data.copy_()`` -- used by
        some optimizers and EMA swaps -- do not bump ``_version``).  In training mode every call re-packs anyway."""
        for eng in self._engines.values():
            eng.weights_key = None

    def check(self, device=None, sync: bool = True) -> None:
        """Synchronise and raise ``Taco2DecError`` (TACO2DEC_E_ABORTED) if an in-kernel watchdog fired in any decoder
        call since the last check -- the teacher-forced entry points return without synchronising, so callers that are
        about to consume the outputs on the host (GTA writer, backward) call this first.  ``sync=False`` does not wait
        for the current stream (for pipelined callers that have already waited on an event)."""
        dev = device if device is not None else self.gate_layer.linear_layer.weight.device
        eng = self._engine(torch.device(dev))
        with torch.cuda.device(eng.device):
            if not sync:     # caller already knows the call of interest has finished (event wait)
                _cabi.check(eng.lib.taco2dec_poll_abort(eng.handle))
                return
            stream = torch.cuda.current_stream(eng.device).cuda_stream
            _cabi.check(eng.lib.taco2dec_check(eng.handle, C.c_void_p(stream)))

    def _bind_weights(self, eng: _Engine) -> None:
        eng.set_mode(self.decoder_path, self.weight_dtype, self.batched_precision)
        key = tuple((t.data_ptr(), t._version) for t in self._weight_tensors())
        # training mode: optimizers may write through .data (no version bump), and a re-pack (<0.5 ms) is noise next
        # to a training step -- always rebuild; eval mode: rebuild on pointer / version change or invalidate_weights()
        if key == eng.weights_key and not self.training:
            return
        eng.weights_key = key
        w = _cabi.Weights()
        for i, s in enumerate(self._sfx()):
            pre, rnn, att = getattr(self, "prenet" + s), getattr(self, "attention_rnn" + s), getattr(self, "attention_layer" + s)
            sw = w.stream[i]
            sw.prenet_w0 = _ptr(self._w(pre.layers[0].linear_layer.weight))
            sw.prenet_w1 = _ptr(self._w(pre.layers[1].linear_layer.weight))
            sw.arnn_w_ih, sw.arnn_w_hh = _ptr(self._w(rnn.weight_ih)), _ptr(self._w(rnn.weight_hh))
            sw.arnn_b_ih, sw.arnn_b_hh = _ptr(self._w(rnn.bias_ih)), _ptr(self._w(rnn.bias_hh))
            sw.query_w = _ptr(self._w(att.query_layer.linear_layer.weight))
            sw.memory_w = _ptr(self._w(att.memory_layer.linear_layer.weight))
            sw.v = _ptr(self._w(att.v_weight()))
            if self.attention_kind == LSA:
                sw.loc_conv_w = _ptr(self._w(att.location_layer.location_conv.conv.weight))
                sw.loc_dense_w = _ptr(self._w(att.location_layer.location_dense.linear_layer.weight))
        w.drnn_w_ih, w.drnn_w_hh = _ptr(self._w(self.decoder_rnn.weight_ih)), _ptr(self._w(self.decoder_rnn.weight_hh))
        w.drnn_b_ih, w.drnn_b_hh = _ptr(self._w(self.decoder_rnn.bias_ih)), _ptr(self._w(self.decoder_rnn.bias_hh))
        w.proj_w = _ptr(self._w(self.linear_projection.linear_layer.weight))
        w.proj_b = _ptr(self._w(self.linear_projection.linear_layer.bias))
        w.gate_w = _ptr(self._w(self.gate_layer.linear_layer.weight))
        w.gate_b = _ptr(self._w(self.gate_layer.linear_layer.bias))
        stream = torch.cuda.current_stream(eng.device).cuda_stream
        _cabi.check(eng.lib.taco2dec_set_weights(eng.handle, C.byref(w), C.c_void_p(stream)))

    def _rng(self, device, keepalive: list) -> _cabi.Rng:
        r = _cabi.Rng()
        seed = self.rng_seed
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())  # follows torch.manual_seed
        r.seed = seed
        rp = self.dropout_replay
        if rp is not None:
            rp = rp.to(device)
            keepalive.append(rp)
            if rp.prenet_keep is not None:
                for s in range(self.n_streams):
                    for l in range(2):
                        r.prenet_keep[s][l] = rp.prenet_keep[s][l].data_ptr()
            if rp.lstm_keep is not None:
                r.lstm_keep = rp.lstm_keep.data_ptr()
            if rp.sma_noise is not None:
                for s in range(self.n_streams):
                    r.sma_noise[s] = rp.sma_noise[s].data_ptr()
        return r

    def _prep(self, t: Optional[torch.Tensor], device, dtype=torch.float32) -> Optional[torch.Tensor]:
        if t is None:
            return None
        t = t.detach().to(device=device, dtype=dtype, non_blocking=True).contiguous()
        if t.data_ptr() % 16:
            t = t.clone()
        return t

    # ------------------------------------------------------------------------------------
    def forward(self, memory, embeddings, decoder_inputs, memory_lengths, bert_lengths=None, independent: bool = False):
        """Teacher-forced pass (model.py:392-428).

        memory [B,T_in,E], embeddings [B,T_sub,E], decoder_inputs [B,n_mel,T], lengths int64 [B]
        -> mel [B,n_mel,T], gate [B,T], alignments [B,T,T_in], alignments_bert [B,T,T_sub].

        ``independent=False`` is the reference's batched behaviour (with SMA, alignment mass leaks onto padded memory
        positions).  ``independent=True`` (extension, used by ``gta.py``) treats every utterance as its own sequence:
        row b equals the batch-1 run on ``memory[b, :memory_lengths[b]]``, which is what the reference's GTA.py computes
        one utterance at a time.

        With autograd enabled the outputs carry a hand-written backward (``_DecoderTF``, tensor path, default dims):
        2 <= B <= 128 in one BPTT call, B == 1 and B > 128 through ``_tf_resized``; other layer sizes raise at
        backward()."""
        wants_grad = torch.is_grad_enabled() and (
            any(p.requires_grad for p in self.parameters()) or memory.requires_grad or
            (embeddings is not None and embeddings.requires_grad))
        if wants_grad and self.decoder_path == "auto" and self.batched_precision == "fp32":
            raise NotImplementedError(
                "tacotron2_subword_b200: batched_precision='fp32' has no backward (the BPTT kernels use fp16 forward / "
                "bf16 backward operands); train with batched_precision='fp16' or evaluate under torch.no_grad()")
        if wants_grad and self._backward_supported(memory):
            B = memory.shape[0]
            if B == 1 or B > self.max_backward_rows:
                return self._tf_resized(self._tf_autograd, memory, embeddings, decoder_inputs, memory_lengths, bert_lengths, independent)
            return self._tf_autograd(memory, embeddings, decoder_inputs, memory_lengths, bert_lengths, independent)
        if self._wants_sub_batches(memory):
            # more utterances than one tensor-path launch takes (tcgen05 N limit): balanced sub-batches instead of the ~15x slower
            # generic kernel
            return self._tf_resized(self._tf_plain, memory, embeddings, decoder_inputs, memory_lengths, bert_lengths, independent)
        outs = self._tf_plain(memory, embeddings, decoder_inputs, memory_lengths, bert_lengths, independent)
        if wants_grad:
            live = [o for o in outs if o is not None]
            marked = list(_NoBackward.apply(self.gate_layer.linear_layer.weight, *live))
            outs = tuple(marked.pop(0) if o is not None else None for o in outs)
        return outs

    def _backward_supported(self, memory) -> bool:
        if self.attention_kind == LSA and (self.loc_filters, self.loc_kernel) != (32, 31):
            return False
        return (memory.shape[0] >= 1 and self.decoder_path in ("auto", "tensor", "tensor_graph")
                and (self.attention_rnn_dim, self.decoder_rnn_dim, self.encoder_embedding_dim, self.prenet_dim,
                     self.attention_dim, self.n_mel_channels) == (1024, 1024, 512, 256, 128, 80))

    def _wants_sub_batches(self, memory) -> bool:
        return (memory.shape[0] > self.max_backward_rows and self._backward_supported(memory) and
                (self.decoder_path != "auto" or self.batched_precision == "fp16"))

    def _tf_plain(self, memory, embeddings, decoder_inputs, memory_lengths, bert_lengths, independent):
        outs, _ = self._run_tf(memory, embeddings, decoder_inputs, memory_lengths, bert_lengths, save=False,
                               independent=independent)
        mel, gate, align, align_b = outs
        return mel.transpose(1, 2), gate, align, align_b

    def _tf_autograd(self, memory, embeddings, decoder_inputs, memory_lengths, bert_lengths, independent):
        """One differentiable teacher-forced pass on the tensor path (2 <= B <= max_backward_rows)."""
        two = self.n_streams == 2
        params = self._weight_tensors()
        outs = _DecoderTF.apply(self, bool(independent), memory, embeddings if two else None, decoder_inputs,
                                memory_lengths, bert_lengths if two else None, *params)
        return outs[0].transpose(1, 2), outs[1], outs[2], (outs[3] if two else None)

    def _tf_resized(self, run, memory, embeddings, decoder_inputs, memory_lengths, bert_lengths, independent):
        """Batches one tensor-path call does not take directly (train.py:330 trains whatever ``collate_fn`` emits, including
        size-1 remainders, data_utils.py:146-160).  Utterances never interact inside the decoder, so

        * B == 1 runs as a batch of two whose second row is a detached copy of the first: its upstream gradients are
          exactly zero, hence so is everything it adds to the weight gradients;
        * B > max_backward_rows (128, the tcgen05 N limit) runs as consecutive sub-batches over the SAME padded memory
          (so the SMA leakage onto padded positions is what the full batch would give); autograd sums their weight
          gradients.  Each sub-batch draws from its own Philox stream; replayed masks are sliced along the batch."""
        B = memory.shape[0]
        two = self.n_streams == 2
        saved = (self.dropout_replay, self.rng_seed, self.validate_lengths)
        rp = saved[0]

        def rows(t, dim, sl, dup):
            if t is None:
                return None
            if dup:
                return torch.cat([t, t], dim)
            idx = [slice(None)] * t.dim()
            idx[dim] = sl
            return t[tuple(idx)]

        def replay_for(sl, dup):
            if rp is None:
                return None
            pk = None if rp.prenet_keep is None else [[rows(m, 1, sl, dup) for m in row] for row in rp.prenet_keep]
            nz = None if rp.sma_noise is None else [rows(n, 1, sl, dup) for n in rp.sma_noise]
            return DropoutReplay(pk, rows(rp.lstm_keep, 2, sl, dup), nz)

        if self.validate_lengths:            # once, on the whole batch (a sub-batch need not contain the longest utterance)
            if memory_lengths is not None and int(memory_lengths.max()) != memory.shape[1]:
                raise ValueError("memory.size(1) must equal max(memory_lengths) (utils.py:11, model.py:414)")
            if two and bert_lengths is not None and int(bert_lengths.max()) != embeddings.shape[1]:
                raise ValueError("embeddings.size(1) must equal max(bert_lengths)")
        try:
            self.validate_lengths = False
            if B == 1:
                dd = lambda t: None if t is None else torch.cat([t, t.detach()], 0)
                self.dropout_replay = replay_for(None, True)
                outs = run(dd(memory), dd(embeddings) if two else None, dd(decoder_inputs), dd(memory_lengths),
                           dd(bert_lengths) if two else None, independent)
                return tuple(None if o is None else o[:1] for o in outs)
            n_sub = -(-B // self.max_backward_rows)           # balanced sub-batches: sizes differ by at most one, none is 1
            bounds = [B * k // n_sub for k in range(n_sub + 1)]
            parts = []
            for k in range(n_sub):
                sl = slice(bounds[k], bounds[k + 1])
                cut = lambda t: None if t is None else t[sl]
                self.dropout_replay = replay_for(sl, False)
                if saved[1] is not None:
                    self.rng_seed = (int(saved[1]) + 0x9E3779B97F4A7C15 * k) % (1 << 62)
                parts.append(run(cut(memory), cut(embeddings) if two else None, cut(decoder_inputs),
                                 cut(memory_lengths), cut(bert_lengths) if two else None, independent))
            return tuple(None if parts[0][i] is None else torch.cat([p_[i] for p_ in parts], 0) for i in range(4))
        finally:
            self.dropout_replay, self.rng_seed, self.validate_lengths = saved

    def _run_tf(self, memory, embeddings, decoder_inputs, memory_lengths, bert_lengths, save: bool, independent: bool = False):
        """One ``taco2dec_forward_teacher_forced`` call; mel is returned in its storage layout [B,T,n_mel]."""
        dev = self.gate_layer.linear_layer.weight.device
        eng = self._engine(dev)
        two = self.n_streams == 2
        mem = self._prep(memory, dev)
        emb = self._prep(embeddings, dev) if two else None
        dec_in = self._prep(decoder_inputs, dev)
        mlen = self._prep(memory_lengths, dev, torch.int64)
        blen = self._prep(bert_lengths, dev, torch.int64) if two else None
        B, T_in, _ = mem.shape
        T_sub = emb.shape[1] if two else 0
        T = dec_in.shape[2]
        if self.validate_lengths:
            # same host sync as the reference's get_mask_from_lengths (utils.py:11)
            if mlen is not None and int(mlen.max()) != T_in:
                raise ValueError("memory.size(1) must equal max(memory_lengths) (utils.py:11, model.py:414)")
            if blen is not None and int(blen.max()) != T_sub:
                raise ValueError("embeddings.size(1) must equal max(bert_lengths)")
        mel = torch.empty(B, T, self.n_mel_channels, device=dev)
        gate = torch.empty(B, T, device=dev)
        align = torch.empty(B, T, T_in, device=dev)
        align_b = torch.empty(B, T, T_sub, device=dev) if two else None
        ws = eng.get_workspace(B, T_in, T_sub, T, True)
        keep = []
        a = _cabi.TFArgs()
        a.B, a.T, a.T_in, a.T_sub = B, T, T_in, T_sub
        a.memory, a.embeddings, a.decoder_inputs = _ptr(mem), _ptr(emb), _ptr(dec_in)
        a.memory_lengths, a.bert_lengths = _ptr(mlen), _ptr(blen)
        a.training = int(self.training)
        a.independent = int(bool(independent))
        a.rng = self._rng(dev, keep)
        a.mel, a.gate, a.align, a.align_bert = _ptr(mel), _ptr(gate), _ptr(align), _ptr(align_b)
        a.workspace, a.workspace_bytes = _ptr(ws), ws.numel()
        pm_given = self._take_pm_hint(mem, emb)
        for i in range(2):
            a.processed_memory[i] = None if pm_given[i] is None else pm_given[i].data_ptr()
        keep.append(pm_given)
        state = None
        if save:
            SL = _cabi.SavedLayout()
            _cabi.check(eng.lib.taco2dec_saved_layout_query(eng.handle, B, T_in, T_sub, T, C.byref(SL)))
            saved = torch.empty(int(SL.total), dtype=torch.uint8, device=dev)
            a.saved, a.saved_bytes = _ptr(saved), saved.numel()
            state = dict(eng=eng, dev=dev, B=B, T=T, T_in=T_in, T_sub=T_sub, mem=mem, emb=emb, dec_in=dec_in, mlen=mlen,
                         blen=blen, training=int(self.training), independent=int(bool(independent)), rng=a.rng, keep=keep, align=align, align_b=align_b,
                         saved=saved, SL=SL, params=self._weight_tensors())
        with torch.cuda.device(dev):
            self._bind_weights(eng)
            stream = torch.cuda.current_stream(dev)
            _cabi.check(eng.lib.taco2dec_forward_teacher_forced(eng.handle, C.byref(a), C.c_void_p(stream.cuda_stream)))
        return (mel, gate, align, align_b), state

    def inference_batched(self, memory, embeddings, memory_lengths=None, bert_lengths=None,
                          max_decoder_steps: Optional[int] = None):
        """B independent free-running utterances in one persistent launch (extension; the
        reference is batch-1 only, model.py:461,480).  Utterance b is *defined* as its own
        batch-1 run on ``memory[b, :memory_lengths[b]]``.

        Returns mel [B,n_mel,Tmax], gate [B,Tmax,1], align [B,Tmax,T_in], align_bert, n_frames [B] (int32),
        reached_max [B] (int32); frames >= n_frames[b] are zeroed (gate there = 1e3 as in parse_output)."""
        if self._wants_sub_batches(memory):
            return self._inference_sub_batches(memory, embeddings, memory_lengths, bert_lengths, max_decoder_steps)
        return self._inference_one_launch(memory, embeddings, memory_lengths, bert_lengths, max_decoder_steps)[:6]

    def _inference_one_launch(self, memory, embeddings, memory_lengths=None, bert_lengths=None, max_decoder_steps=None):
        """``inference_batched`` for at most one launch's worth of utterances; also returns the host copy of
        (n_frames, reached_max) that the call reads back anyway, so that callers need no second device->host read."""
        dev = self.gate_layer.linear_layer.weight.device
        eng = self._engine(dev)
        two = self.n_streams == 2
        mem = self._prep(memory, dev)
        emb = self._prep(embeddings, dev) if two else None
        mlen = self._prep(memory_lengths, dev, torch.int64)
        blen = self._prep(bert_lengths, dev, torch.int64) if two else None
        B, T_in, _ = mem.shape
        T_sub = emb.shape[1] if two else 0
        steps = int(max_decoder_steps if max_decoder_steps is not None else self.max_decoder_steps)
        mel = torch.empty(B, steps, self.n_mel_channels, device=dev)
        gate = torch.empty(B, steps, device=dev)
        align = torch.empty(B, steps, T_in, device=dev)
        align_b = torch.empty(B, steps, T_sub, device=dev) if two else None
        counters = torch.zeros(2, B, dtype=torch.int32, device=dev)      # one fill for both
        n_frames, reached = counters[0], counters[1]
        ws = eng.get_workspace(B, T_in, T_sub, steps, False)
        keep = []
        a = _cabi.InferArgs()
        a.B, a.T_in, a.T_sub, a.max_decoder_steps = B, T_in, T_sub, steps
        a.gate_threshold = float(self.gate_threshold)
        a.memory, a.embeddings = _ptr(mem), _ptr(emb)
        a.memory_lengths, a.bert_lengths = _ptr(mlen), _ptr(blen)
        a.rng = self._rng(dev, keep)
        a.mel, a.gate, a.align, a.align_bert = _ptr(mel), _ptr(gate), _ptr(align), _ptr(align_b)
        a.n_frames, a.reached_max = _ptr(n_frames), _ptr(reached)
        a.workspace, a.workspace_bytes = _ptr(ws), ws.numel()
        pm_given = self._take_pm_hint(mem, emb)
        for i in range(2):
            a.processed_memory[i] = None if pm_given[i] is None else pm_given[i].data_ptr()
        with torch.cuda.device(dev):
            self._bind_weights(eng)
            stream = torch.cuda.current_stream(dev)
            _cabi.check(eng.lib.taco2dec_infer(eng.handle, C.byref(a), C.c_void_p(stream.cuda_stream)))
        # one device->host read per call (the reference does one per FRAME, model.py:480)
        counters_host = counters.cpu()
        nf = counters_host[0]
        _cabi.check(eng.lib.taco2dec_check(eng.handle, C.c_void_p(stream.cuda_stream)))
        Tmax = int(nf.max())
        if int(nf.min()) == Tmax:       # every utterance ran to the same frame (always so for B = 1): nothing to blank
            return (mel[:, :Tmax].transpose(1, 2), gate[:, :Tmax].unsqueeze(2), align[:, :Tmax],
                    align_b[:, :Tmax] if two else None, n_frames, reached, counters_host)
        frame_ids = torch.arange(Tmax, device=dev).unsqueeze(0)
        dead = frame_ids >= n_frames.unsqueeze(1).to(torch.int64)                # [B,Tmax]
        mel = mel[:, :Tmax].masked_fill(dead.unsqueeze(2), 0.0).transpose(1, 2)
        gate = gate[:, :Tmax].masked_fill(dead, 1e3).unsqueeze(2)
        align = align[:, :Tmax].masked_fill(dead.unsqueeze(2), 0.0)
        if two:
            align_b = align_b[:, :Tmax].masked_fill(dead.unsqueeze(2), 0.0)
        return mel, gate, align, align_b, n_frames, reached, counters_host

    def _inference_sub_batches(self, memory, embeddings, memory_lengths, bert_lengths, max_decoder_steps):
        """More utterances than one persistent launch takes: balanced sub-batches (utterances are independent), outputs padded
        to the longest utterance exactly as one call would (mel / alignments 0, gate 1e3 beyond an utterance's last frame)."""
        B = memory.shape[0]
        two = self.n_streams == 2
        n_sub = -(-B // self.max_backward_rows)
        bounds = [B * k // n_sub for k in range(n_sub + 1)]
        saved = (self.dropout_replay, self.rng_seed)
        rp = saved[0]
        parts = []
        try:
            for k in range(n_sub):
                sl = slice(bounds[k], bounds[k + 1])
                cut = lambda t: None if t is None else t[sl]
                if rp is not None:
                    pk = None if rp.prenet_keep is None else [[m[:, sl] for m in row] for row in rp.prenet_keep]
                    self.dropout_replay = DropoutReplay(pk, None, None)
                if saved[1] is not None:
                    self.rng_seed = (int(saved[1]) + 0x9E3779B97F4A7C15 * k) % (1 << 62)
                parts.append(self.inference_batched(cut(memory), cut(embeddings) if two else None, cut(memory_lengths),
                                                    cut(bert_lengths) if two else None, max_decoder_steps))
        finally:
            self.dropout_replay, self.rng_seed = saved
        Tmax = max(p_[0].shape[2] for p_ in parts)
        grow = lambda t, dim, val: t if t.shape[dim] == Tmax else F.pad(
            t, [0, 0] * (t.dim() - 1 - dim) + [0, Tmax - t.shape[dim]], value=val)
        cat = lambda i, dim, val: torch.cat([grow(p_[i], dim, val) for p_ in parts], 0)
        return (cat(0, 2, 0.0), cat(1, 1, 1e3), cat(2, 1, 0.0), cat(3, 1, 0.0) if two else None,
                torch.cat([p_[4] for p_ in parts]), torch.cat([p_[5] for p_ in parts]))

    def inference(self, memory, embeddings=None):
        """Free-running decode (model.py:430-492): mel [B,n_mel,T], gate [B,T,1], alignments,
        alignments_bert, INFER_FLAG (False when max_decoder_steps was reached)."""
        if memory.shape[0] != 1:
            raise ValueError("Decoder.inference is batch-1 as in the reference (model.py:461,480); "
                             "use inference_batched for B > 1")
        mel, gate, align, align_b, n_frames, reached, counters_host = self._inference_one_launch(memory, embeddings)
        flag = not bool(int(counters_host[1, 0]))
        if not flag:
            print("Warning! Reached max decoder steps")
        return mel, gate, align, align_b, flag


class _PostnetTrain(torch.autograd.Function):
    """Training-mode Postnet (model.py:27-70 under model.train()) with a hand-written backward on the repo's own kernels
    (csrc/postnet_train.cuh): per layer one tcgen05 "rows x weights^T" contraction over the channel-last, halo-padded
    activations + fused BatchNorm(batch statistics) / tanh / dropout; backward = the element-wise derivatives, one tcgen05
    weight-gradient product per tap and the transposed-convolution contraction.  The reference relies on autograd."""

    @staticmethod
    def forward(ctx, net, x, *params):
        lib, dev = _cabi.load_library(), x.device
        h = net._handle(dev)
        B, C0, T = x.shape
        N, L = B * T, net.n_layers
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        seed = net.rng_seed
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        replay = net.dropout_replay
        def padded(c):                      # [B][T + 4][c], only the two-frame halos need zeroing
            t_ = torch.empty(B, T + 4, c, device=dev)
            t_[:, :2].zero_()
            t_[:, T + 2:].zero_()
            return t_

        xpad = padded(C0)
        xpad[:, 2:T + 2] = x.detach().transpose(1, 2)
        saved, out = [], None
        with torch.cuda.device(dev):
            for l in range(L):
                w, b, gamma, beta = (t.detach() for t in params[4 * l: 4 * l + 4])
                bn = net.convolutions[l][1]
                Cout, Cin, K = w.shape
                wt = w.permute(0, 2, 1).reshape(Cout, K * Cin).contiguous()          # [co][(k, ci)]
                y = torch.empty(N, Cout, device=dev)
                stats = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
                ws = net._workspace(dev, int(lib.taco2dec_postnet_rows_gemm_workspace_bytes(h, Cout, K * Cin, N)))
                _cabi.check(lib.taco2dec_postnet_rows_gemm(h, _ptr(xpad), (T + 4) * Cin, Cin, B, T, K * Cin, _ptr(wt), Cout, _ptr(b.contiguous()),
                                                           0, _ptr(y), Cout, _ptr(stats), _ptr(ws), ws.numel(), stream))
                mean64 = stats[:Cout] / N
                var64 = (stats[Cout:] / N - mean64 * mean64).clamp_min(0.0)                  # biased, as BatchNorm normalises with
                mean, rstd = mean64.float(), (var64 + bn.eps).rsqrt().float()
                if bn.track_running_stats and bn.running_mean is not None:
                    mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked + 1)
                    bn.running_mean.mul_(1 - mom).add_(mean, alpha=mom)
                    bn.running_var.mul_(1 - mom).add_((var64 * (N / max(N - 1, 1))).float(), alpha=mom)
                    bn.num_batches_tracked += 1
                last = l == L - 1
                keep = None if replay is None else replay[l].to(device=dev, dtype=torch.uint8).contiguous()
                g32, b32 = gamma.contiguous(), beta.contiguous()
                if last:
                    out = torch.empty(B, Cout, T, device=dev)
                    dst, osb, ost, osc = out, Cout * T, 1, T
                    nxt = None
                else:
                    nxt = padded(Cout)
                    dst, osb, ost, osc = nxt[:, 2:], (T + 4) * Cout, Cout, 1
                _cabi.check(lib.taco2dec_postnet_bn_act_forward(h, _ptr(y), B, T, Cout, _ptr(mean), _ptr(rstd), _ptr(g32), _ptr(b32),
                                                                int(not last), seed, l, net.p_dropout, _ptr(keep), _ptr(dst), osb, ost, osc,
                                                                stream))
                saved.append((xpad, y, mean, rstd, keep))
                xpad = nxt
        ctx.net, ctx.saved, ctx.seed, ctx.shape = net, saved, seed, (B, C0, T)
        ctx.params = params
        return out

    @staticmethod
    def backward(ctx, d_out):
        net, saved, seed = ctx.net, ctx.saved, ctx.seed
        if saved is None:
            raise RuntimeError("postnet backward called twice (the saved activations are released after the first pass)")
        ctx.saved = None
        B, C0, T = ctx.shape
        N, L = B * T, net.n_layers
        lib, dev = _cabi.load_library(), d_out.device
        h = net._handle(dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        d = d_out.detach().to(torch.float32).transpose(1, 2).reshape(N, -1).contiguous()
        if d.data_ptr() == d_out.data_ptr():
            d = d.clone()                       # the element-wise backward works in place
        grads = [None] * (4 * L)
        need_dx = ctx.needs_input_grad[1]
        with torch.cuda.device(dev):
            for l in reversed(range(L)):
                w, b, gamma, beta = (t.detach() for t in ctx.params[4 * l: 4 * l + 4])
                xpad, y, mean, rstd, keep = saved[l]
                saved[l] = None
                Cout, Cin, K = w.shape
                g32, b32 = gamma.contiguous(), beta.contiguous()
                sums = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
                _cabi.check(lib.taco2dec_postnet_bn_act_backward(h, _ptr(d), _ptr(y), N, Cout, _ptr(mean), _ptr(rstd), _ptr(g32), _ptr(b32),
                                                                 int(l != L - 1), seed, l, net.p_dropout, _ptr(keep), _ptr(sums), stream))
                grads[4 * l + 3] = sums[:Cout].float()                                  # d beta  = sum dz
                grads[4 * l + 2] = sums[Cout:].float()                                  # d gamma = sum dz . zhat
                m_dz, m_dzz = (sums[:Cout] / N).float(), (sums[Cout:] / N).float()
                dypad = torch.empty(B, T + 4, Cout, device=dev)
                dypad[:, :2].zero_()
                dypad[:, T + 2:].zero_()
                _cabi.check(lib.taco2dec_postnet_bn_backward_input(h, _ptr(d), _ptr(y), B, T, Cout, _ptr(mean), _ptr(rstd), _ptr(g32),
                                                                   _ptr(m_dz), _ptr(m_dzz), _ptr(dypad), stream))
                # d W[co][(k, ci)] = sum_(b,t) dy[b][t][co] . x_pad[b][t + k][ci]: ONE tcgen05 product over the overlapping im2col rows
                dw = torch.empty(Cout, K * Cin, device=dev)
                ws = net._workspace(dev, int(lib.taco2dec_postnet_wgrad_workspace_bytes(h, Cout, K * Cin, T, B)))
                y3 = dypad[:, 2:T + 2]
                _cabi.check(lib.taco2dec_postnet_wgrad(h, _ptr(y3), Cout, (T + 4) * Cout, Cout, _ptr(xpad), Cin, (T + 4) * Cin, K * Cin, T, B,
                                                       _ptr(dw), K * Cin, 0, 0, _ptr(ws), ws.numel(), stream))
                grads[4 * l] = dw.view(Cout, K, Cin).permute(0, 2, 1).contiguous()
                grads[4 * l + 1] = torch.zeros(Cout, device=dev)      # a bias in front of a batch-statistics BatchNorm has exactly zero gradient
                if l > 0 or need_dx:
                    # d x[n][ci] = sum_{k,co} dy_pad[n + k][co] . W[co][ci][K-1-k]   (transposed convolution as rows x weights^T)
                    wd = w.flip(2).permute(1, 2, 0).reshape(Cin, K * Cout).contiguous()
                    dn = torch.empty(N, Cin, device=dev)
                    ws = net._workspace(dev, int(lib.taco2dec_postnet_rows_gemm_workspace_bytes(h, Cin, K * Cout, N)))
                    _cabi.check(lib.taco2dec_postnet_rows_gemm(h, _ptr(dypad), (T + 4) * Cout, Cout, B, T, K * Cout, _ptr(wd), Cin, None, 1,
                                                               _ptr(dn), Cin, None, _ptr(ws), ws.numel(), stream))
                    d = dn
        dx = d.view(B, T, C0).transpose(1, 2) if need_dx else None
        pg = tuple(g if p_.requires_grad else None for g, p_ in zip(grads, ctx.params))
        return (None, dx) + pg


class Postnet(nn.Module):
    """model.py:27-70: five conv1d(k=5)+BatchNorm, tanh on all but the last, dropout 0.5 in training.

    ``forward`` in training mode runs on the repo's own kernels with a hand-written backward (``_PostnetTrain``); in eval mode it
    is the reference module (PyTorch ops).  ``mel_postnet`` is what the model classes
    call: mel + postnet(mel) with the output mask, which in eval mode on a B200 runs as five tcgen05 GEMMs with the
    BatchNorm folded in (``taco2dec_postnet_*``, csrc/postnet.cuh) instead of ten cuDNN/elementwise launches."""

    def __init__(self, hparams):
        super().__init__()
        hp = hparams
        n, k, c = hp.postnet_n_convolutions, hp.postnet_kernel_size, hp.postnet_embedding_dim
        chans = [hp.n_mel_channels] + [c] * (n - 1) + [hp.n_mel_channels]
        self.convolutions = nn.ModuleList(
            nn.Sequential(ConvNorm(chans[i], chans[i + 1], kernel_size=k, stride=1, padding=(k - 1) // 2, dilation=1,
                                   w_init_gain="tanh" if i < n - 1 else "linear"),
                          nn.BatchNorm1d(chans[i + 1]))
            for i in range(n))
        self.n_mel, self.dim, self.kernel, self.n_layers = hp.n_mel_channels, c, k, n
        self.fused_eval = True          # extension: use the CUDA postnet when the module is in eval mode
        # "fp32": operands as split fp16 pairs (hi.hi + hi.lo + lo.hi, fp32-grade mel_postnet, default);
        # "fp16": plain fp16 operands (~8e-4 of the output scale, 2.5x less tensor work)
        self.fused_precision = "fp32"
        self._fused = {}                # device index -> (handle, weights key, workspace)
        # training mode (batch statistics, dropout, backward) on the repo's own kernels (csrc/postnet_train.cuh, fp16 operands /
        # fp32 accumulation: the grade of cuDNN's default TF32 convolutions); False = the reference's PyTorch ops
        self.fused_train = True
        self.p_dropout = 0.5            # model.py:60, 67
        self.rng_seed: Optional[int] = None           # fixed Philox seed for the dropout masks; None = fresh per call
        self.dropout_replay = None      # parity runs: one uint8 keep mask [B*T, C_out] per layer
        self._train_ws = {}

    def _handle(self, dev):
        key = dev.index if dev.index is not None else torch.cuda.current_device()
        ent = self._fused.get(key)
        if ent is None:
            h = C.c_void_p()
            _cabi.check(_cabi.load_library().taco2dec_postnet_create(self.n_mel, self.dim, self.kernel, self.n_layers, key, C.byref(h)))
            ent = self._fused[key] = {"h": h, "key": None, "ws": None}
        return ent["h"]

    def _workspace(self, dev, need: int) -> torch.Tensor:
        key = dev.index if dev.index is not None else torch.cuda.current_device()
        ws = self._train_ws.get(key)
        if ws is None or ws.numel() < need:
            ws = self._train_ws[key] = torch.empty(need, dtype=torch.uint8, device=dev)
        return ws

    def _train_fusable(self, x: torch.Tensor) -> bool:
        return (self.fused_train and self.training and x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and self.kernel == 5
                and self.n_mel <= 128 and self.n_mel % 4 == 0 and self.dim % 128 == 0 and self.dim <= 1024 and 2 <= self.n_layers <= 8)

    def forward(self, x):
        if self._train_fusable(x):
            params = []
            for seq in self.convolutions:
                params += [seq[0].conv.weight, seq[0].conv.bias, seq[1].weight, seq[1].bias]
            return _PostnetTrain.apply(self, x, *params)
        last = len(self.convolutions) - 1
        for i, conv in enumerate(self.convolutions):
            x = conv(x)
            x = F.dropout(torch.tanh(x) if i < last else x, 0.5, self.training)
        return x

    def _fusable(self, mel: torch.Tensor) -> bool:
        return (self.fused_eval and not self.training and mel.is_cuda and mel.dtype == torch.float32 and self.kernel == 5
                and self.n_mel <= 128 and self.dim % 128 == 0 and 2 <= self.n_layers <= 8
                and not (torch.is_grad_enabled() and (mel.requires_grad or any(p.requires_grad for p in self.parameters()))))

    def mel_postnet(self, mel: torch.Tensor, output_lengths: Optional[torch.Tensor] = None,
                    independent: bool = False) -> torch.Tensor:
        """mel [B, n_mel, T] (any strides) -> mel + postnet(mel) (model.py:557-558); with ``output_lengths`` the frames
        beyond each utterance's length are zero (model.py:531-541).  ``independent=True`` (extension, batched synthesis):
        frames beyond the length do not exist at all, so row b equals the batch-1 result on its own frames."""
        if independent and output_lengths is None:
            raise ValueError("independent=True needs output_lengths")
        if not self._fusable(mel):
            if independent:
                out = torch.zeros_like(mel)
                for b in range(mel.shape[0]):
                    n = int(output_lengths[b])
                    out[b:b + 1, :, :n] = mel[b:b + 1, :, :n] + self.forward(mel[b:b + 1, :, :n])
                return out
            out = mel + self.forward(mel)
            if output_lengths is not None:
                invalid = torch.arange(mel.shape[2], device=mel.device)[None, :] >= output_lengths.to(mel.device)[:, None]
                out = out.masked_fill(invalid[:, None, :], 0.0)
            return out
        lib = _cabi.load_library()
        dev = mel.device
        key = dev.index if dev.index is not None else torch.cuda.current_device()
        self._handle(dev)
        ent = self._fused[key]
        tensors = []
        for seq in self.convolutions:
            conv, bn = seq[0].conv, seq[1]
            tensors += [conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        wkey = tuple((t.data_ptr(), t._version) for t in tensors) + (self.fused_precision,)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            if wkey != ent["key"]:
                _cabi.check(lib.taco2dec_postnet_set_precision(ent["h"], int(self.fused_precision == "fp16")))
                w = _cabi.PostnetWeights()
                w.n_layers, w.bn_eps = self.n_layers, float(self.convolutions[0][1].eps)
                for i in range(self.n_layers):
                    for name, t in zip(("conv_w", "conv_b", "bn_weight", "bn_bias", "bn_mean", "bn_var"), tensors[6 * i: 6 * i + 6]):
                        if t.dtype != torch.float32 or not t.is_contiguous():
                            raise _cabi.Taco2DecError("postnet parameters must be contiguous fp32")
                        setattr(w.layer[i], name, t.data_ptr())
                _cabi.check(lib.taco2dec_postnet_set_weights(ent["h"], C.byref(w), stream))
                ent["key"] = wkey
            B, _, T = mel.shape
            need = int(lib.taco2dec_postnet_workspace_bytes(ent["h"], B, T))
            if ent["ws"] is None or ent["ws"].numel() < need:
                ent["ws"] = torch.empty(need, dtype=torch.uint8, device=dev)
            out = torch.empty(B, self.n_mel, T, device=dev)
            lens = None if output_lengths is None else output_lengths.to(device=dev, dtype=torch.int64).contiguous()
            sb, sc, st = mel.stride()
            _cabi.check(lib.taco2dec_postnet_forward(ent["h"], _ptr(mel.detach()), sb, sc, st, B, T, _ptr(lens), int(independent),
                                                     _ptr(out), _ptr(ent["ws"]), ent["ws"].numel(), stream))
        return out

    def __del__(self):
        try:
            lib = _cabi.load_library()
            for ent in self._fused.values():
                lib.taco2dec_postnet_destroy(ent["h"])
        except Exception:
            pass


class MemoryPrep:
    """Decoder inputs on the GPU path: ``memory = linear_converter(cat(encoder_outputs, cls))``
    (model.py:548-549 / 553-554) and ``processed_memory = memory_layer(memory)`` (model.py:258-261) as two chained tcgen05
    GEMMs with split-fp16 operands (fp32-grade results, ``taco2dec_memprep_*``).  Inference / no-grad only; one instance
    per stream.  The modules keep owning the parameters (state_dict layout unchanged)."""

    def __init__(self, converter: LinearNorm, memory_layer: LinearNorm):
        self.converter, self.memory_layer = converter, memory_layer
        self._h = {}           # device index -> {"h": handle, "key": weights key, "ws": workspace}

    def usable(self, enc: torch.Tensor, cls: torch.Tensor) -> bool:
        w = self.converter.linear_layer.weight
        e, c, a = enc.shape[-1], cls.shape[-1], self.memory_layer.linear_layer.weight.shape[0]
        return (enc.is_cuda and enc.dtype == torch.float32 and cls.dtype == torch.float32 and e % 128 == 0 and a % 128 == 0
                and (e + c) % 64 == 0 and w.shape == (e, e + c) and self.converter.linear_layer.bias is not None
                and not (torch.is_grad_enabled() and (enc.requires_grad or cls.requires_grad or w.requires_grad)))

    def __call__(self, enc: torch.Tensor, cls: torch.Tensor):
        """enc [B,T,enc_dim], cls [B,T,cls_dim] -> (memory [B,T,enc_dim], processed_memory [B,T,attn_dim])."""
        lib = _cabi.load_library()
        dev = enc.device
        key = dev.index if dev.index is not None else torch.cuda.current_device()
        wc, bc, wm = self.converter.linear_layer.weight, self.converter.linear_layer.bias, self.memory_layer.linear_layer.weight
        B, T, E = enc.shape
        ent = self._h.get(key)
        if ent is None:
            h = C.c_void_p()
            _cabi.check(lib.taco2dec_memprep_create(E, cls.shape[-1], wm.shape[0], key, C.byref(h)))
            ent = self._h[key] = {"h": h, "key": None, "ws": None}
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            wkey = tuple((t.data_ptr(), t._version) for t in (wc, bc, wm))
            if wkey != ent["key"]:
                for t in (wc, bc, wm):
                    if t.dtype != torch.float32 or not t.is_contiguous():
                        raise _cabi.Taco2DecError("converter / memory-layer parameters must be contiguous fp32")
                _cabi.check(lib.taco2dec_memprep_set_weights(ent["h"], _ptr(wc.detach()), _ptr(bc.detach()), _ptr(wm.detach()), stream))
                ent["key"] = wkey
            n = B * T
            need = int(lib.taco2dec_memprep_workspace_bytes(ent["h"], n))
            if ent["ws"] is None or ent["ws"].numel() < need:
                ent["ws"] = torch.empty(need, dtype=torch.uint8, device=dev)
            enc_c, cls_c = enc.detach().contiguous(), cls.detach().contiguous()
            memory = torch.empty(B, T, E, device=dev)
            pm = torch.empty(B, T, wm.shape[0], device=dev)
            _cabi.check(lib.taco2dec_memprep_forward(ent["h"], _ptr(enc_c), _ptr(cls_c), n, _ptr(memory), _ptr(pm), _ptr(ent["ws"]),
                                                     ent["ws"].numel(), stream))
        return memory, pm

    def __del__(self):
        try:
            lib = _cabi.load_library()
            for ent in self._h.values():
                lib.taco2dec_memprep_destroy(ent["h"])
        except Exception:
            pass


class Encoder(nn.Module):
    """model.py:73-125: 3x (conv1d k=5 + BatchNorm + ReLU + dropout) then a BiLSTM."""

    def __init__(self, hparams):
        super().__init__()
        hp = hparams
        d, k = hp.encoder_embedding_dim, hp.encoder_kernel_size
        self.convolutions = nn.ModuleList(
            nn.Sequential(ConvNorm(d, d, kernel_size=k, stride=1, padding=(k - 1) // 2, dilation=1, w_init_gain="relu"),
                          nn.BatchNorm1d(d))
            for _ in range(hp.encoder_n_convolutions))
        self.lstm = nn.LSTM(d, d // 2, 1, batch_first=True, bidirectional=True)

    def _convs(self, x):
        for conv in self.convolutions:
            x = F.dropout(F.relu(conv(x)), 0.5, self.training)
        return x.transpose(1, 2)

    def forward(self, x, input_lengths):
        x = self._convs(x)
        packed = nn.utils.rnn.pack_padded_sequence(x, input_lengths.cpu().numpy(), batch_first=True,
                                                   enforce_sorted=False)
        self.lstm.flatten_parameters()
        out, _ = self.lstm(packed)
        out, _ = nn.utils.rnn.pad_packed_sequence(out, batch_first=True)
        return out

    def inference(self, x):
        x = self._convs(x)
        self.lstm.flatten_parameters()
        out, _ = self.lstm(x)
        return out

    def inference_independent(self, x, input_lengths):
        """Eval-mode batch in which every row equals ``inference`` on that utterance alone (extension).  The reference's
        batched ``forward`` lets the convolutions see the padding symbols' embeddings; here positions beyond a row's
        length are held at zero before every convolution -- exactly the zero padding a batch-1 call sees -- and the
        BiLSTM runs on the packed sequences."""
        valid = (torch.arange(x.shape[2], device=x.device)[None, :] < input_lengths.to(x.device)[:, None]).unsqueeze(1)
        x = x * valid
        for conv in self.convolutions:
            x = F.relu(conv(x)) * valid
        packed = nn.utils.rnn.pack_padded_sequence(x.transpose(1, 2), input_lengths.cpu(), batch_first=True, enforce_sorted=False)
        self.lstm.flatten_parameters()
        out, _ = self.lstm(packed)
        out, _ = nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=x.shape[2])
        return out


def _uniform_embedding(n, dim, n_symbols_for_std):
    emb = nn.Embedding(n, dim)
    val = sqrt(3.0) * sqrt(2.0 / (n_symbols_for_std + dim))  # model.py:503-506
    emb.weight.data.uniform_(-val, val)
    return emb


def _mask_outputs(outputs, output_lengths, n_mel, enabled):
    """model.py:531-541: beyond output_lengths mel(s) <- 0, gate <- 1e3."""
    if enabled and output_lengths is not None:
        dead = ~get_mask_from_lengths(output_lengths)
        outputs[0].data.masked_fill_(dead.unsqueeze(1), 0.0)
        outputs[1].data.masked_fill_(dead.unsqueeze(1), 0.0)
        outputs[2].data.masked_fill_(dead, 1e3)
    return outputs


class BERT_Tacotron2(nn.Module):
    """model.py:494-582 -- dual-stream Tacotron2 (phoneme stream + sub-word stream)."""

    def __init__(self, hparams):
        super().__init__()
        hp = hparams
        self.mask_padding = hp.mask_padding
        self.fp16_run = hp.fp16_run
        self.n_mel_channels = hp.n_mel_channels
        self.n_frames_per_step = hp.n_frames_per_step
        self.embedding = _uniform_embedding(hp.n_symbols, hp.symbols_embedding_dim, hp.n_symbols)
        self.embedding_sub = _uniform_embedding(hp.sub_n_symbols, hp.symbols_embedding_dim, hp.n_symbols)
        self.encoder = Encoder(hp)
        self.encoder_sub = Encoder(hp)
        self.linear_converter = LinearNorm(hp.encoder_embedding_dim + hp.BERT_embedding_dim, hp.encoder_embedding_dim)
        self.linear_converter_sub = LinearNorm(hp.encoder_embedding_dim + hp.BERT_embedding_dim,
                                               hp.encoder_embedding_dim)
        self.decoder = Decoder(hp)
        self.postnet = Postnet(hp)
        # extension: decoder inputs through the CUDA path when no gradient is needed (plain attributes, not sub-modules)
        self.fused_memory = True
        object.__setattr__(self, "_memprep", None)

    def _convert(self, enc, pcls, enc_s, bcls):
        """(memory, memory_sub) = the two linear converters (model.py:548-549, 553-554).  Without autograd on a CUDA device the
        converter and the attention layers' memory_layer run as chained tcgen05 GEMMs and the processed memories are handed
        to the decoder's next call; otherwise the reference's PyTorch modules run."""
        if self._memprep is None:
            object.__setattr__(self, "_memprep", (MemoryPrep(self.linear_converter, self.decoder.attention_layer.memory_layer),
                                                  MemoryPrep(self.linear_converter_sub, self.decoder.attention_layer_bert.memory_layer)))
        mp0, mp1 = self._memprep
        if self.fused_memory and mp0.usable(enc, pcls) and mp1.usable(enc_s, bcls):
            mem, pm = mp0(enc, pcls)
            mem_s, pm_s = mp1(enc_s, bcls)
            self.decoder.set_processed_memory(mem, pm, mem_s, pm_s)
            return mem, mem_s
        return self.linear_converter(torch.cat([enc, pcls], 2)), self.linear_converter_sub(torch.cat([enc_s, bcls], 2))

    def parse_batch(self, batch):
        (text_padded, input_lengths, input_lengths_bert, mel_padded, gate_padded, output_lengths, embeddings,
         phoneme_embeddings_cls, bert_embeddings_cls, align_padded) = batch
        text_padded = to_gpu(text_padded).long()
        input_lengths = to_gpu(input_lengths).long()
        input_lengths_bert = to_gpu(input_lengths_bert).long()
        max_input_len = torch.max(torch.cat((input_lengths, input_lengths_bert), 0).data).item()
        max_output_len = torch.max(output_lengths.data).item()
        mel_padded = to_gpu(mel_padded).float()
        gate_padded = to_gpu(gate_padded).float()
        output_lengths = to_gpu(output_lengths).long()
        align_padded = to_gpu(align_padded).float()
        x = (text_padded, input_lengths, input_lengths_bert, mel_padded, (max_input_len, max_output_len),
             output_lengths, embeddings, phoneme_embeddings_cls, bert_embeddings_cls)
        return x, (mel_padded, gate_padded, align_padded)

    def parse_output(self, outputs, output_lengths=None):
        return _mask_outputs(outputs, output_lengths, self.n_mel_channels, self.mask_padding)

    def _memories(self, text, sub, pcls, bcls, text_lengths=None, bert_lengths=None):
        e = self.embedding(text).transpose(1, 2)
        es = self.embedding_sub(sub).transpose(1, 2)
        if text_lengths is None:
            enc, enc_s = self.encoder.inference(e), self.encoder_sub.inference(es)
        else:
            enc, enc_s = self.encoder(e, text_lengths), self.encoder_sub(es, bert_lengths)
        return self._convert(enc, pcls, enc_s, bcls)                    # model.py:548-549, 553-554

    def forward(self, inputs):
        (text_inputs, text_lengths, bert_lengths, mels, _max_lens, output_lengths, embeddings,
         phoneme_embeddings_cls, bert_embeddings_cls) = inputs
        text_lengths, bert_lengths, output_lengths = text_lengths.data, bert_lengths.data, output_lengths.data
        mem, mem_s = self._memories(text_inputs, embeddings, phoneme_embeddings_cls, bert_embeddings_cls,
                                    text_lengths, bert_lengths)
        mel, gate, align, align_b = self.decoder(mem, mem_s, mels, text_lengths, bert_lengths)   # model.py:556
        mel_post = self.postnet.mel_postnet(mel)
        return self.parse_output([mel, mel_post, gate, align, align_b], output_lengths)

    def inference(self, inputs, embeddings, phoneme_embeddings_cls, bert_embeddings_cls):
        mem, mem_s = self._memories(inputs, embeddings, phoneme_embeddings_cls, bert_embeddings_cls)
        mel, gate, align, align_b, flag = self.decoder.inference(mem, mem_s)                      # model.py:574-575
        mel_post = self.postnet.mel_postnet(mel)
        return self.parse_output([mel, mel_post, gate, align, align_b, flag])


    @torch.no_grad()
    def inference_batch(self, inputs, embeddings, phoneme_embeddings_cls, bert_embeddings_cls,
                        max_decoder_steps: Optional[int] = None):
        """Batched form of ``inference`` (extension; the reference synthesises one utterance per call, inference.py:361-375).

        Every argument is a list with one entry per utterance, each entry exactly what ``inference`` takes ([1, T_in] ids,
        [1, T_sub] ids, [1, T_in, 768] and [1, T_sub, 768] BERT embeddings).  Utterances are encoded together with
        ``Encoder.inference_independent`` (rows equal batch-1 encoding), decoded together (``Decoder.inference_batched``: every row is its own
        batch-1 run, stop frames per utterance) and post-processed together (``Postnet.mel_postnet(independent=True)``).
        Returns one ``inference``-shaped list [mel, mel_postnet, gate, align, align_bert, INFER_FLAG] per utterance."""
        n = len(inputs)
        dev = self.embedding.weight.device

        def pad(items, dtype=None):
            L = max(int(t.shape[1]) for t in items)
            out = torch.zeros((n, L) + tuple(items[0].shape[2:]), dtype=dtype or items[0].dtype, device=dev)
            for i, t in enumerate(items):
                out[i, : t.shape[1]] = t[0].to(dev)
            return out

        mlen = torch.tensor([int(t.shape[1]) for t in inputs], dtype=torch.int64, device=dev)
        blen = torch.tensor([int(t.shape[1]) for t in embeddings], dtype=torch.int64, device=dev)
        text, sub = pad(inputs, torch.int64), pad(embeddings, torch.int64)
        enc = self.encoder.inference_independent(self.embedding(text).transpose(1, 2), mlen)
        enc_s = self.encoder_sub.inference_independent(self.embedding_sub(sub).transpose(1, 2), blen)
        mem, mem_s = self._convert(enc, pad(phoneme_embeddings_cls), enc_s, pad(bert_embeddings_cls))   # model.py:548-549, 553-554
        mel, gate, align, align_b, n_frames, reached = self.decoder.inference_batched(mem, mem_s, mlen, blen, max_decoder_steps)
        mel_post = self.postnet.mel_postnet(mel, n_frames.to(torch.int64), independent=True)
        nf, rm = n_frames.tolist(), reached.tolist()
        outs = []
        for i in range(n):
            t, li, lb = int(nf[i]), int(mlen[i]), int(blen[i])
            outs.append([mel[i:i + 1, :, :t], mel_post[i:i + 1, :, :t], gate[i:i + 1, :t], align[i:i + 1, :t, :li],
                         align_b[i:i + 1, :t, :lb], not bool(rm[i])])
        return outs


class Tacotron2(nn.Module):
    """Single-stream compat class for the reference's stale callers (GTA.py:21,57-59;
    inference.py:302,334; streamlitNews.py:121,150), with the upstream-NVIDIA signature they use:
    ``forward((text, input_lengths, mel, max_len, output_lengths)) -> (mel, mel_postnet, gate, align)`` and
    ``inference(sequence) -> (mel, mel_postnet, gate, align)``.  The reference itself defines no such
    class, so its parity is pinned against the oracle only."""

    def __init__(self, hparams):
        super().__init__()
        hp = hparams
        self.mask_padding = hp.mask_padding
        self.n_mel_channels = hp.n_mel_channels
        self.embedding = _uniform_embedding(hp.n_symbols, hp.symbols_embedding_dim, hp.n_symbols)
        self.encoder = Encoder(hp)
        self.decoder = Decoder(hp, n_streams=1)
        self.postnet = Postnet(hp)

    def forward(self, inputs):
        text, input_lengths, mels, _max_len, output_lengths = inputs
        input_lengths, output_lengths = input_lengths.data, output_lengths.data
        mem = self.encoder(self.embedding(text).transpose(1, 2), input_lengths)
        mel, gate, align, _ = self.decoder(mem, None, mels, input_lengths, None)
        mel_post = self.postnet.mel_postnet(mel)
        out = _mask_outputs([mel, mel_post, gate, align], output_lengths, self.n_mel_channels, self.mask_padding)
        return tuple(out)

    def inference(self, sequence):
        mem = self.encoder.inference(self.embedding(sequence).transpose(1, 2))
        mel, gate, align, _, _flag = self.decoder.inference(mem, None)
        mel_post = self.postnet.mel_postnet(mel)
        return mel, mel_post, gate, align
