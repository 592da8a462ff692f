"""Latency-path determinism / parity probe: python tools/repro_latency.py [fp32|fp16] [steps] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.synth import SMA, make_decoder_weights, make_inputs
from tacotron2_subword_b200 import Decoder, create_hparams
wdt = sys.argv[1] if len(sys.argv) > 1 else "fp32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
w = make_decoder_weights(SMA, seed=1, gate_bias=-20.0)
inp = make_inputs(1, 21, 7, 1, seed=1)
hp = create_hparams(); hp.max_decoder_steps = steps
dec = Decoder(hp); dec.load_state_dict(w); dec = dec.cuda().eval(); dec.rng_seed = 1
mem, emb = inp["memory"].cuda(), inp["embeddings"].cuda()
dec.decoder_path = "generic"
with torch.no_grad():
    ref = dec.inference_batched(mem, emb)
torch.cuda.synchronize()
dec.decoder_path, dec.weight_dtype = "latency", wdt
for r in range(reps):
    try:
        with torch.no_grad():
            out = dec.inference_batched(mem, emb)
        torch.cuda.synchronize()
    except Exception as e:
        print("EXC", str(e).splitlines()[0]); break
    d = (out[0] - ref[0]).abs().amax(dim=(0, 1))       # per-frame max-abs mel diff vs generic
    da = (out[2] - ref[2]).abs().amax(dim=(0, 2))
    bad = [i for i, v in enumerate(d.tolist()) if v > (1e-4 if wdt == "fp32" else 1e-3)]
    print(f"rep {r} {wdt} frames {int(out[4][0])} mel-diff max {float(d.max()):.2e} first-bad-frames {bad[:6]} align-diff {float(da.max()):.2e}")
