"""Timeline of one frame of one CTA of the persistent batched kernel (SM-clock stamps, kcyc relative to the first event).
usage: TACO2DEC_PB_DEBUG=<frame> [TACO2DEC_PB_DEBUG_CTA=<cta>] python tools/pb_timeline.py fr|tf B T_in T_sub T"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TACO2DEC_PB_DEBUG", "20")
import torch
from oracle.synth import SMA, make_decoder_weights, make_inputs
from tacotron2_subword_b200 import Decoder, create_hparams, _cabi

mode, B, T_in, T_sub, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
w = make_decoder_weights(SMA, seed=1234, gate_bias=-20.0)
dec = Decoder(create_hparams()); dec.load_state_dict(w); dec = dec.cuda().eval(); dec.rng_seed = 7
dec.decoder_path = "tensor"
eng = dec._engine(torch.device("cuda", 0))
inp = make_inputs(B, T_in, T_sub, T if mode == "tf" else 1, seed=3, ragged=B > 1)
mem, emb = inp["memory"].cuda(), inp["embeddings"].cuda()
ml, bl = inp["memory_lengths"].cuda(), inp["bert_lengths"].cuda()
with torch.no_grad():
    for _ in range(2):
        if mode == "tf":
            dec(mem, emb, inp["mels"].cuda(), ml, bl)
        else:
            dec.inference_batched(mem, emb, ml, bl, max_decoder_steps=T)
buf = (C.c_longlong * 256)()
_cabi.check(eng.lib.taco2dec_read_debug_stamps(eng.handle, C.c_void_p(torch.cuda.current_stream().cuda_stream), buf))
v = list(buf)
ev = []
PH = ["acc1 ready", "epi1 signalled", "partials1 ready", "pointwise1 done (h1 signalled)", "h1 (all) ready", "attention done (ctx signalled)",
      "ctx projection done", "acc2 ready", "epi2 signalled", "partials2 ready", "pointwise2 done (h2 signalled)", "h2 (all) ready",
      "mel sum done", "prenet L0 done", "prenet L1 done"]
for i in range(32):
    if v[i]: ev.append((v[i], f"producer: X tile {i} requested"))
    if v[32 + i]: ev.append((v[32 + i], f"mma: operands of tile {i} landed"))
    if v[64 + i]: ev.append((v[64 + i], f"mma: tile {i} issued"))
    if v[96 + i]: ev.append((v[96 + i], f"producer: weight tile {i} requested"))
for i, n in enumerate(PH):
    if v[128 + i]: ev.append((v[128 + i], f"COMPUTE: {n}"))
for j in range(2):
    for k, n in enumerate(["loop top", "after weight cursor", "after dependency check", "ring slot free", "TMA issued"]):
        if v[160 + 8 * j + k]: ev.append((v[160 + 8 * j + k], f"   producer detail tile {19 + j}: {n} (last pass)"))
for k, n in enumerate(["pw1: start", "pw1: partials loaded", "pw1: cells done", "pw1: barrier passed", "pw1: chunk stores issued",
                       "pw1: q partials issued", "pw1: barrier passed", "pw1: signalled"]):
    if v[176 + k]: ev.append((v[176 + k], f"   compute detail {n}"))
ev.sort()
t0 = ev[0][0]
for t, n in ev:
    print(f"{(t - t0) / 1e3:9.2f} kcyc  {n}")
