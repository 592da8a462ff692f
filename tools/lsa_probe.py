"""Batch-1 free-running decode with location-sensitive attention (cfg-2 shape): latency kernel vs the generic kernel.
usage: python tools/lsa_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.synth import LSA, make_decoder_weights, make_inputs
from tacotron2_subword_b200 import Decoder, create_hparams

hp = create_hparams(); hp.attention = LSA; hp.max_decoder_steps = 1000
w = make_decoder_weights(LSA, seed=1234, gate_bias=-20.0)
dec = Decoder(hp); dec.load_state_dict(w); dec = dec.cuda().eval(); dec.rng_seed = 1
eng = dec._engine(torch.device("cuda", 0)); eng.set_profiling(True)
inp = make_inputs(1, 150, 50, 1, seed=3)
mem, emb = inp["memory"].cuda(), inp["embeddings"].cuda()
rows = []
for path, wdt in (("latency", "fp32"), ("latency", "fp16"), ("generic", "fp32")):
    dec.decoder_path, dec.weight_dtype = path, wdt
    ms = []
    with torch.no_grad():
        for i in range(4):
            out = dec.inference_batched(mem, emb)
            torch.cuda.synchronize()
            if i:
                ms.append(eng.last_kernel_ms())
    n = out[0].shape[-1]
    rows.append(dict(config="B=1 LSA 150/50 free-running", path=eng.last_path(), weights=wdt, frames=n,
                     us_per_frame=round(1e3 * min(ms) / n, 2), frames_per_s=round(n / (min(ms) * 1e-3))))
    print(json.dumps(rows[-1]), flush=True)
json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "lsa_probe.json"), "w"), indent=1)
