"""Persistent batched forward kernel vs the per-frame graph (same fp16-operand arithmetic) over a sweep of batch sizes, memory lengths
and stop patterns, free-running and teacher-forced: stop frames must agree, outputs within 2e-4.  usage: python tools/pb_stress.py"""
import os, sys, itertools, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.synth import SMA, DecoderDims, make_decoder_weights, make_inputs
from tacotron2_subword_b200 import Decoder, create_hparams

worst_all, bad = 0.0, 0
for S in (2, 1):
    dims = DecoderDims(streams=S)
    for B, T_in, bias in itertools.product((2, 3, 5, 8, 13, 16, 31, 33, 64, 100, 128), (7, 120, 290), (-20.0, -0.5)):
        T_sub, steps = max(1, T_in // 3), 24
        w = make_decoder_weights(SMA, seed=11, gate_bias=bias, dims=dims)
        inp = make_inputs(B, T_in, T_sub if S == 2 else 1, steps, seed=B + T_in, ragged=True, dims=dims)
        res = []
        for path in ("tensor", "tensor_graph"):
            hp = create_hparams()
            dec = Decoder(hp, n_streams=S); dec.load_state_dict(w); dec = dec.cuda().eval(); dec.rng_seed = 5
            dec.decoder_path = path
            args = (inp["memory"].cuda(), inp["embeddings"].cuda() if S == 2 else None, inp["memory_lengths"].cuda(),
                    inp["bert_lengths"].cuda() if S == 2 else None)
            with torch.no_grad():
                fr = dec.inference_batched(*args, max_decoder_steps=steps)
                tf = dec(args[0], args[1], inp["mels"].cuda(), args[2], args[3])
            dec.check()
            res.append((fr, tf))
        (fa, ta), (fb, tb) = res
        same_stop = torch.equal(fa[4], fb[4]) and torch.equal(fa[5], fb[5])
        n = min(fa[0].shape[2], fb[0].shape[2])
        d = max(float((fa[0][:, :, :n] - fb[0][:, :, :n]).abs().max()), float((ta[0] - tb[0]).abs().max()),
                float((ta[2] - tb[2]).abs().max()))
        worst_all = max(worst_all, d)
        flag = "" if (same_stop and d < 2e-4) else "   <-- CHECK"
        bad += bool(flag)
        print(f"S={S} B={B:3d} T_in={T_in:3d} gate bias {bias:5.1f}: stop frames equal {same_stop} ({int(fa[4].min())}..{int(fa[4].max())}), max diff {d:.2e}{flag}", flush=True)
print(f"worst difference over the sweep: {worst_all:.2e}; cases to check: {bad}")
