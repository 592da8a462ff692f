"""Print per-phase SM-clock shares of the persistent decoder kernel (CTA 0's view).
usage: python tools/phase_profile.py [fr|tf] [generic|latency] [fp32|fp16]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_problem, CFG
from tacotron2_subword_b200 import Decoder, create_hparams

mode = sys.argv[1] if len(sys.argv) > 1 else "fr"
path = sys.argv[2] if len(sys.argv) > 2 else "latency"
wdt = sys.argv[3] if len(sys.argv) > 3 else "fp32"
w, inp = make_problem()
hp = create_hparams(); hp.max_decoder_steps = CFG["max_steps"]
dec = Decoder(hp); dec.load_state_dict(w); dec = dec.cuda().eval(); dec.rng_seed = 1
dec.decoder_path, dec.weight_dtype = path, wdt
eng = dec._engine(torch.device("cuda", 0)); eng.set_profiling(True)
mem, emb = inp["memory"].cuda(), inp["embeddings"].cuda()
for it in range(3):
    with torch.no_grad():
        if mode == "fr":
            out = dec.inference_batched(mem, emb)
        else:
            dec(mem, emb, torch.randn(1, 80, 1000).cuda(), torch.tensor([150]).cuda(), torch.tensor([50]).cuda())
    ms = eng.last_kernel_ms(); pc = eng.phase_clocks()
tot = sum(pc)
print(f"mode={mode} path={eng.last_path()} weights={wdt} kernel {ms:.2f} ms ({ms:.1f} us/frame); cycles {tot/1e6:.1f}M -> {tot/ms/1e3:.0f} MHz")
if eng.last_path() == "generic":
    names = ["P0a", "P0b", "A", "Q", "B", "C", "D"]
    for k, n in enumerate(names):
        print(f"  {n:4s} work {pc[2*k]/1e6:8.2f} kcyc/frame ({100*pc[2*k]/tot:5.1f}%)   barrier {pc[2*k+1]/1e6:8.2f} kcyc/frame ({100*pc[2*k+1]/tot:5.1f}%)")
else:
    names = ["b compute", "wait h2[t-1]", "c compute", "wait prenet[t] (aux chain)", "d: pointwise + sync + q partials",
             "wait h1[t]", "e compute", "wait ctx[t] (attention)", "f: last sync + zero", "a(t+1) compute", "f: consume (warp 0)", "f: barrier after consume",
             "f: pointwise + publish", "d: consume (warp 0)", "d: barrier after consume"]
    for k, n in enumerate(names):
        print(f"  {n:32s} {pc[k]/1e6:8.2f} kcyc/frame ({100*pc[k]/tot:5.1f}%)")
