"""Per-role cycle sums of the batch-1 latency kernel (LSTM CTA 0 per-step detail, attention CTAs (0,0) / (1,0), aux CTA 0).
usage: TACO2DEC_LAT_DEBUG=1 python tools/lat_stamps.py [fp32|fp16]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TACO2DEC_LAT_DEBUG", "1")
import torch
from bench import make_problem, CFG
from tacotron2_subword_b200 import Decoder, create_hparams, _cabi

wdt = sys.argv[1] if len(sys.argv) > 1 else "fp32"
w, inp = make_problem()
hp = create_hparams(); hp.max_decoder_steps = CFG["max_steps"]
dec = Decoder(hp); dec.load_state_dict(w); dec = dec.cuda().eval(); dec.rng_seed = 1
dec.decoder_path, dec.weight_dtype = "latency", wdt
eng = dec._engine(torch.device("cuda", 0)); eng.set_profiling(True)
mem, emb = inp["memory"].cuda(), inp["embeddings"].cuda()
for it in range(3):
    with torch.no_grad():
        out = dec.inference_batched(mem, emb)
ms = eng.last_kernel_ms()
n = out[0].shape[-1]
buf = (C.c_longlong * 256)()
_cabi.check(eng.lib.taco2dec_read_debug_stamps(eng.handle, C.c_void_p(torch.cuda.current_stream().cuda_stream), buf))
v = list(buf)
print(f"weights={wdt} kernel {ms:.2f} ms, {n} frames, {1e3 * ms / n:.2f} us/frame")
print("LSTM CTA 0, warp 0, per step (kcyc/frame): dot product | reduce + accumulate | calls/frame")
for si, name in enumerate(["a W_hh.h1", "b W_ih.ctx", "c Wd_hh.h2", "d W_ih.prenet", "e0 Wd_ih.h1(0)", "e1 Wd_ih.h1(1)", "f Wd_ih.ctx"]):
    print(f"  {name:16s} {v[si*4]/n/1e3:7.2f} | {v[si*4+1]/n/1e3:7.2f} | {v[si*4+2]/n:5.2f}")
AN = ["wait stop word + barrier", "q gather (poll)", "barrier", "q combine + barrier", "energies: barrier after the exchange",
      "alignment + barrier", "context partial + barrier", "context reduce + publish", "barrier",
      "energies: own partials + store", "energies: exchange (poll, thread 0)"]
for s in range(2):
    print(f"attention CTA (stream {s}, slice 0), kcyc/frame:")
    for k, nm in enumerate(AN):
        print(f"  {nm:28s} {v[64 + 16*s + k]/n/1e3:7.2f}")
XN = ["loop top", "poll ctx", "poll h2 (wait)", "barrier", "projection", "barrier", "L0 partial + store", "L0 gather (poll)", "barrier",
      "L1 + publish", "barrier"]
print("aux CTA 0, kcyc/frame:")
for k, nm in enumerate(XN):
    print(f"  {nm:28s} {v[96 + k]/n/1e3:7.2f}")
