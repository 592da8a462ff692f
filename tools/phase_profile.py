"""Print per-phase SM-clock shares of the persistent decoder kernel (CTA 0's view)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_problem, CFG
from tacotron2_subword_b200 import Decoder, create_hparams

mode = sys.argv[1] if len(sys.argv) > 1 else "fr"
w, inp = make_problem()
hp = create_hparams(); hp.max_decoder_steps = CFG["max_steps"]
dec = Decoder(hp); dec.load_state_dict(w); dec = dec.cuda().eval(); dec.rng_seed = 1
eng = dec._engine(torch.device("cuda", 0)); eng.set_profiling(True)
mem, emb = inp["memory"].cuda(), inp["embeddings"].cuda()
names = ["P0a", "P0b", "A", "Q", "B", "C", "D"]
for it in range(3):
    with torch.no_grad():
        if mode == "fr":
            dec.inference_batched(mem, emb)
        else:
            dec(mem, emb, torch.randn(1, 80, 1000).cuda(), torch.tensor([150]).cuda(), torch.tensor([50]).cuda())
    ms = eng.last_kernel_ms(); pc = eng.phase_clocks()
tot = sum(pc)
print(f"mode={mode} kernel {ms:.2f} ms; cycles total {tot/1e6:.1f}M -> {tot/ms/1e3:.0f} MHz")
for k, n in enumerate(names):
    print(f"  {n:4s} work {pc[2*k]/1000/1000:8.2f} kcyc/frame ({100*pc[2*k]/tot:5.1f}%)   barrier {pc[2*k+1]/1e6:8.2f} kcyc/frame ({100*pc[2*k+1]/tot:5.1f}%)")
print(f"  init {pc[15]/1e3:.1f} kcyc")
