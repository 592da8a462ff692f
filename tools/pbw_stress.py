"""Persistent BPTT kernel vs the per-frame graph over a sweep of shapes (same operands, replayed masks): worst relative gradient
difference per shape.  usage: python tools/pbw_stress.py"""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.synth import SMA, DecoderDims, make_decoder_weights, make_inputs, make_dropout_plan
from tests.gpu_util import make_decoder, replay_of

worst_all = 0.0
for S in (2, 1):
    dims = DecoderDims(streams=S)
    w = make_decoder_weights(SMA, seed=5, dims=dims)
    for B, T_in, T in itertools.product((2, 5, 17, 32, 33, 48, 64), (9, 160, 301), (1, 7)):
        T_sub = max(1, T_in // 3)
        inp = make_inputs(B, T_in, T_sub if S == 2 else 1, T, seed=B + T_in, ragged=True, dims=dims)
        plan = make_dropout_plan(B, T + 1, T, T_in, T_sub if S == 2 else 1, True, seed=7, dims=dims)
        res = []
        for no_persist in (False, True):
            os.environ.pop("TACO2DEC_NO_PERSIST_BWD", None)
            if no_persist:
                os.environ["TACO2DEC_NO_PERSIST_BWD"] = "1"
            dec = make_decoder(w, SMA, n_streams=S, exact=False).train()
            dec.dropout_replay = replay_of(plan)
            mem = inp["memory"].cuda().requires_grad_(True)
            emb = inp["embeddings"].cuda().requires_grad_(True) if S == 2 else None
            outs = dec(mem, emb, inp["mels"].cuda(), inp["memory_lengths"].cuda(), inp["bert_lengths"].cuda() if S == 2 else None)
            g = torch.Generator().manual_seed(1)
            loss = sum((o * torch.randn(o.shape, generator=g).cuda()).sum() for o in outs if o is not None)
            loss.backward()
            torch.cuda.synchronize()
            dec.check()
            gr = {n: p.grad.clone() for n, p in dec.named_parameters() if p.grad is not None}
            gr["memory"] = mem.grad.clone()
            res.append(gr)
        worst = max(float((res[0][n] - res[1][n]).abs().max()) / max(float(res[1][n].abs().max()), 1e-30) for n in res[0])
        worst_all = max(worst_all, worst)
        flag = "" if worst < 2e-3 else "   <-- CHECK"
        print(f"S={S} B={B:3d} T_in={T_in:3d} T={T}: worst relative difference {worst:.2e}{flag}", flush=True)
print(f"worst over the sweep: {worst_all:.2e}")
