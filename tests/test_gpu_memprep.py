"""-m gpu: decoder inputs on the CUDA path: linear_converter + memory_layer as chained tcgen05 GEMMs
with split-fp16 operands, through the C ABI (taco2dec_memprep_*), against the CPU oracle and the reference golden.

Stated bound: max|got - want| <= 1e-5 * max(1, max|want|) for memory AND processed memory (fp32-grade: these tensors are
inputs of the whole recurrence); measured ~1e-6."""
import os

import numpy as np
import pytest
import torch

from oracle.memprep_oracle import make_memprep_inputs, make_memprep_weights, memory_prepare
from tacotron2_subword_b200.layers import LinearNorm
from tacotron2_subword_b200.model import MemoryPrep

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _modules(w):
    conv = LinearNorm(512 + 768, 512)
    meml = LinearNorm(512, 128, bias=False)
    conv.load_state_dict({"linear_layer.weight": w["linear_converter.linear_layer.weight"],
                          "linear_layer.bias": w["linear_converter.linear_layer.bias"]})
    meml.load_state_dict({"linear_layer.weight": w["memory_layer.linear_layer.weight"]})
    return conv.cuda(), meml.cuda()


def _check(got, want, tag):
    for n, g, w_ in zip(("memory", "processed_memory"), got, want):
        err = float((g.cpu().double() - w_.double()).abs().max()) / max(1.0, float(w_.abs().max()))
        assert err <= TOL, (tag, n, err)


def test_memprep_matches_reference_golden():
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "memprep.npz"))
    seed, B, T = int(z["seed"]), int(z["B"]), int(z["T"])
    w = make_memprep_weights(seed)
    enc, cls = make_memprep_inputs(B, T, seed + 1)
    mp = MemoryPrep(*_modules(w))
    with torch.no_grad():
        assert mp.usable(enc.cuda(), cls.cuda())
        got = mp(enc.cuda(), cls.cuda())
    _check(got, (torch.from_numpy(z["memory"]), torch.from_numpy(z["processed_memory"])), "golden")


@pytest.mark.parametrize("B,T", [(1, 150), (1, 1), (3, 43), (64, 120), (128, 160)])
def test_memprep_vs_oracle(B, T):
    """One utterance (2 row groups, split-K), a single row, rows not a multiple of 128, and the batched BASELINE shapes."""
    w = make_memprep_weights(7 + B)
    enc, cls = make_memprep_inputs(B, T, 11 + T)
    want = memory_prepare(w, enc, cls, dtype=torch.float64)
    mp = MemoryPrep(*_modules(w))
    with torch.no_grad():
        got = mp(enc.cuda(), cls.cuda())
    assert got[0].shape == (B, T, 512) and got[1].shape == (B, T, 128)
    _check(got, want, f"B={B} T={T}")


def test_memprep_picks_up_weight_updates_and_refuses_autograd():
    w = make_memprep_weights(3)
    conv, meml = _modules(w)
    mp = MemoryPrep(conv, meml)
    enc, cls = make_memprep_inputs(2, 20, 5)
    with torch.no_grad():
        a = mp(enc.cuda(), cls.cuda())[0].clone()
        conv.linear_layer.weight.mul_(0.5)          # version bump -> re-pack
        b = mp(enc.cuda(), cls.cuda())[0]
    w2 = dict(w); w2["linear_converter.linear_layer.weight"] = w["linear_converter.linear_layer.weight"] * 0.5
    _check((b, mp(enc.cuda(), cls.cuda())[1]), memory_prepare(w2, enc, cls, dtype=torch.float64), "after update")
    assert float((a - b).abs().max()) > 1e-3
    assert not mp.usable(enc.cuda().requires_grad_(True), cls.cuda())       # training goes through the PyTorch modules


def test_model_inference_with_and_without_fused_memory():
    """BERT_Tacotron2.inference: the fused converter + the processed memory handed to the decoder give the same utterance as
    the PyTorch converter + the decoder's own memory projection."""
    from tacotron2_subword_b200 import BERT_Tacotron2, create_hparams
    torch.manual_seed(1234)
    hp = create_hparams()
    model = BERT_Tacotron2(hp).cuda().eval()
    model.decoder.rng_seed = 5
    model.decoder.max_decoder_steps = 40
    with torch.no_grad():
        model.decoder.gate_layer.linear_layer.bias.fill_(-20.0)
    T_in, T_sub = 37, 12
    text = torch.randint(0, hp.n_symbols, (1, T_in)).cuda()
    sub = torch.randint(0, hp.sub_n_symbols, (1, T_sub)).cuda()
    pcls = torch.randn(1, T_in, hp.BERT_embedding_dim).cuda()
    bcls = torch.randn(1, T_sub, hp.BERT_embedding_dim).cuda()
    outs = {}
    for fused in (True, False):
        model.fused_memory = fused
        with torch.no_grad():
            outs[fused] = model.inference(text, sub, pcls, bcls)
    assert outs[True][0].shape == outs[False][0].shape == (1, 80, 40)
    assert float((outs[True][0] - outs[False][0]).abs().max()) <= 1e-4
    assert float((outs[True][3] - outs[False][3]).abs().max()) <= 1e-5
