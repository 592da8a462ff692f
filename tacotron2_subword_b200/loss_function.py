"""Drop-in ``Tacotron2Loss`` (/root/reference/loss_function.py:7-66): same constructor, same ``forward(model_output,
targets, x, iters)`` and the same 5-tuple ``(loss, mel_loss, gate_loss, align_loss, align_bert_loss)``.

On a B200 the default variant (``alignloss == ""``) and the L2 alignment variant run as ONE sweep over the outputs
(``taco2dec_loss_forward``, csrc/loss.cuh) that also writes d loss / d output for every output, the mel gradient directly in
the [B, T, n_mel] storage order the decoder's BPTT consumes; ``backward()`` only scales those buffers.  The KL variant
(loss_function.py:34-57) and CPU tensors take the reference's PyTorch formulation.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _cabi


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mel, mel_post, gate, mel_target, gate_target, align, align_bert, align_target):
        lib = _cabi.load_library()
        dev = mel.device
        B, M, T = mel.shape
        post_c, gate_c = mel_post.contiguous(), gate.contiguous()
        tgt_c, gtgt_c = mel_target.contiguous().float(), gate_target.contiguous().float()
        d_mel = torch.empty(B, T, M, device=dev)
        d_post = torch.empty(B, M, T, device=dev)
        d_gate = torch.empty(B, T, device=dev)
        losses = torch.empty(5, device=dev)
        a = _cabi.LossArgs()
        a.B, a.n_mel, a.T = B, M, T
        a.mel = _ptr(mel)
        a.mel_stride_b, a.mel_stride_c, a.mel_stride_t = mel.stride()
        a.mel_postnet, a.gate, a.mel_target, a.gate_target = _ptr(post_c), _ptr(gate_c), _ptr(tgt_c), _ptr(gtgt_c)
        d_al = [None, None]
        keep = [post_c, gate_c, tgt_c, gtgt_c]
        for s, al in enumerate((align, align_bert)):
            if al is None:
                continue
            al_c, at_c = al.contiguous(), align_target.contiguous().float()
            if al_c.shape != at_c.shape:
                raise ValueError("alignment and alignment target shapes differ (the reference's nn.MSELoss would reject this too)")
            d_al[s] = torch.empty_like(al_c)
            a.align[s], a.align_target[s], a.T_align[s], a.d_align[s] = al_c.data_ptr(), at_c.data_ptr(), al_c.shape[2], d_al[s].data_ptr()
            keep += [al_c, at_c]
        need = int(lib.taco2dec_loss_workspace_bytes(B, M, T))
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        a.d_mel, a.d_mel_postnet, a.d_gate, a.losses = _ptr(d_mel), _ptr(d_post), _ptr(d_gate), _ptr(losses)
        a.workspace, a.workspace_bytes = _ptr(ws), need
        with torch.cuda.device(dev):
            _cabi.check(lib.taco2dec_loss_forward(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        ctx.grads = (d_mel, d_post, d_gate, d_al[0], d_al[1])
        ctx.gate_shape = gate.shape
        ctx.mark_non_differentiable(losses)
        return losses[0], losses

    @staticmethod
    def backward(ctx, g_total, _g_parts):
        d_mel, d_post, d_gate, d_al, d_alb = ctx.grads
        ctx.grads = None
        sc = lambda t: None if t is None else t * g_total
        # d_mel lives in [B, T, n_mel]; the forward input was the transposed view, so hand the matching view back
        return (sc(d_mel).transpose(1, 2), sc(d_post), sc(d_gate).view(ctx.gate_shape), None, None, sc(d_al), sc(d_alb), None)


class Tacotron2Loss(nn.Module):
    def __init__(self, alignloss=""):
        super().__init__()
        self.alignloss = alignloss
        self.fused = True          # extension: one-sweep CUDA path when the inputs are CUDA tensors

    def forward(self, model_output, targets, x, iters=0):
        mel_target, gate_target, align_target = targets[0], targets[1], targets[2]
        mel_out, mel_out_postnet, gate_out, align_out, align_bert_out = model_output
        use_align = self.alignloss == "L2" and iters < 40000
        kl = self.alignloss == "KL" and iters < 40000
        if self.fused and mel_out.is_cuda and mel_out.dtype == torch.float32 and not kl:
            total, parts = _FusedLoss.apply(mel_out, mel_out_postnet, gate_out, mel_target, gate_target,
                                            align_out if use_align else None, align_bert_out if use_align else None,
                                            align_target if use_align else None)
            al = parts[3] if use_align else None
            alb = parts[4] if use_align else None
            return total, parts[1], parts[2], al, alb
        return self._reference_formulation(model_output, targets, x, iters)

    def _reference_formulation(self, model_output, targets, x, iters):
        """loss_function.py:12-66 with PyTorch ops (CPU tensors, KL variant)."""
        mel_target, gate_target, align_target = targets[0], targets[1], targets[2]
        gate_target = gate_target.view(-1, 1)
        mel_out, mel_out_postnet, gate_out, align_out, align_bert_out = model_output
        gate_out = gate_out.view(-1, 1)
        mel_loss = nn.MSELoss()(mel_out, mel_target) + nn.MSELoss()(mel_out_postnet, mel_target)
        gate_loss = nn.BCEWithLogitsLoss()(gate_out, gate_target)
        align_loss = align_bert_loss = None
        if self.alignloss == "L2" and iters < 40000:
            align_loss = nn.MSELoss()(align_out, align_target)
            align_bert_loss = nn.MSELoss()(align_bert_out, align_target)
        elif self.alignloss == "KL" and iters < 40000:
            eps = 0.000001
            ao = torch.where(align_out == 0, torch.full_like(align_out, eps), align_out)
            abo = torch.where(align_bert_out == 0, torch.full_like(align_bert_out, eps), align_bert_out)
            at = torch.where(align_target == 0, torch.full_like(align_target, eps), align_target)
            text_len, mel_len = x[1], x[4]
            align_loss = align_bert_loss = 0
            for b in range(at.size(0)):
                n = min(int(mel_len[b]) - 1, int(text_len[b]) - 1)     # the reference slices the FRAME axis twice (loss_function.py:49-51)
                a_, ab_, t_ = ao[b][:n], abo[b][:n], at[b][:n]
                align_loss = align_loss + torch.mean(torch.sum(t_ * (torch.log(t_) - torch.log(a_)), dim=-1))
                align_bert_loss = align_bert_loss + torch.mean(torch.sum(t_ * (torch.log(t_) - torch.log(ab_)), dim=-1))
        total = mel_loss + gate_loss
        if align_loss is not None:
            total = total + align_loss
        if align_bert_loss is not None:
            total = total + align_bert_loss
        return total, mel_loss, gate_loss, align_loss, align_bert_loss
