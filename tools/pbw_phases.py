"""Per-phase SM-clock shares of the persistent backward kernel (CTA 0).  usage: python tools/pbw_phases.py [B] [T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.synth import SMA, make_decoder_weights, make_inputs
from tacotron2_subword_b200 import Decoder, create_hparams

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
w = make_decoder_weights(SMA, seed=1234)
dec = Decoder(create_hparams()); dec.load_state_dict(w); dec = dec.cuda().train()
eng = dec._engine(torch.device("cuda", 0)); eng.set_profiling(True)
inp = make_inputs(B, 160, 53, T, seed=3, ragged=True)
mem, emb, mels = inp["memory"].cuda(), inp["embeddings"].cuda(), inp["mels"].cuda()
ml, bl = inp["memory_lengths"].cuda(), inp["bert_lengths"].cuda()
for _ in range(3):
    dec.zero_grad(set_to_none=True)
    outs = dec(mem, emb, mels, ml, bl)
    loss = outs[0].square().mean() + outs[1].square().mean()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loss.backward(); e1.record(); e1.synchronize()
ms, kms, pc = e0.elapsed_time(e1), eng.last_kernel_ms(), eng.phase_clocks()
names = ["pw2 (step 0) / loop top", "wait acc2", "epilogue 2 + signal", "wait X2 (all)", "pw2 (frame ahead) + signal", "wait X1 (stream)", "d prenet save",
         "attention tasks", "wait dq (stream)", "pw1 (dq load, Wq^T dq, cells, stores) + signal", "wait acc1", "epilogue 1 + signal",
         "  attention: early loads, waits, d ctx", "  attention: d alpha' (memory rows)", "  attention: recurrence + energies (pm, dpm rows)",
         "  attention: reduce, dq / dv, signal"]
tot = sum(pc)
print(f"B={B} T={T}: backward {ms:.2f} ms, persistent kernel {kms:.2f} ms = {1e3 * kms / T:.1f} us/frame; CTA 0: {tot / T / 1e3:.1f} kcyc/frame")
for n, v in zip(names, pc):
    print(f"  {n:50s} {v / T / 1e3:7.2f} kcyc/frame ({100 * v / max(tot, 1):5.1f}%)")
print("(the four attention rows are parts of what the 'attention tasks' row reported before they were split out: that row now holds only the loop overhead)")

if os.environ.get("TACO2DEC_PBW_DEBUG"):
    import ctypes as C
    from tacotron2_subword_b200 import _cabi
    buf = (C.c_longlong * 256)()
    _cabi.check(eng.lib.taco2dec_read_debug_stamps(eng.handle, C.c_void_p(torch.cuda.current_stream().cuda_stream), buf))
    v = list(buf)
    t0 = v[0]
    print(f"timeline of the attention-LSTM product, CTA 0, step {os.environ['TACO2DEC_PBW_DEBUG']} (kcyc after the CTA published its own gate gradients):")
    print(f"  counter complete (all CTAs of the stream published): {(v[1] - t0) / 1e3:7.2f}")
    for i in range(16):
        print(f"  tile {i:2d}: requested {(v[16 + i] - t0) / 1e3:7.2f}   operands landed {(v[32 + i] - t0) / 1e3:7.2f}   issued {(v[48 + i] - t0) / 1e3:7.2f}")
    print(f"  accumulator complete: {(v[2] - t0) / 1e3:7.2f}")
