"""Time the batched tensor-core paths (persistent kernel vs the per-frame graph) on the BASELINE cfg-3/4/5 shapes and print
the persistent kernel's per-phase clock shares (CTA 0).   usage: python tools/pb_probe.py [quick] [nograph]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.synth import SMA, make_decoder_weights, make_inputs
from tacotron2_subword_b200 import Decoder, create_hparams

quick = "quick" in sys.argv
w = make_decoder_weights(SMA, seed=1234, gate_bias=-20.0)
dec = Decoder(create_hparams()); dec.load_state_dict(w); dec = dec.cuda().eval(); dec.rng_seed = 7
eng = dec._engine(torch.device("cuda", 0)); eng.set_profiling(True)
PH = ["wait acc1", "epi1+signal", "wait partials1", "pointwise1+q", "wait h1", "attention", "ctx projection", "wait acc2",
      "epi2+signal", "wait partials2", "pointwise2", "wait h2 (all)", "mel sum+stop", "wait mel, prenet L0", "wait L0, prenet L1", "-"]


def run(name, mode, B, T_in, T_sub, T, path, train=False):
    inp = make_inputs(B, T_in, T_sub, T if mode == "tf" else 1, seed=3, ragged=True)
    mem, emb = inp["memory"].cuda(), inp["embeddings"].cuda()
    ml, bl = inp["memory_lengths"].cuda(), inp["bert_lengths"].cuda()
    dec.decoder_path = path
    dec.train(train)
    if mode == "tf":
        mels = inp["mels"].cuda()
        fn = lambda: dec(mem, emb, mels, ml, bl)
    else:
        fn = lambda: dec.inference_batched(mem, emb, ml, bl, max_decoder_steps=T)
    ts, ks = [], []
    with torch.no_grad():
        for i in range(3 if quick else 4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            if i:
                ts.append(e0.elapsed_time(e1))
                if eng.last_path() == "tensor":
                    ks.append(eng.last_kernel_ms())
    dec.check()
    ms = min(ts)
    r = dict(config=name, path=eng.last_path(), B=B, frames=T, ms=round(ms, 3), us_per_frame_step=round(1e3 * ms / T, 2),
             frames_per_s=round(B * T / (ms * 1e-3)))
    if ks:
        r["kernel_us_per_frame"] = round(1e3 * min(ks) / T, 2)
        pc = eng.phase_clocks(); tot = max(1, sum(pc))
        r["phases_kcyc_per_frame"] = {n: round(v / T / 1e3, 2) for n, v in zip(PH, pc) if v}
    print(json.dumps(r), flush=True)
    return r


rows = []
paths = ["tensor"] + ([] if "nograph" in sys.argv else ["tensor_graph"])
for path in paths:
    rows.append(run("cfg3 FR B=64 120/40", "fr", 64, 120, 40, 200 if quick else 1000, path))
    rows.append(run("cfg3-strong FR B=8", "fr", 8, 120, 40, 200 if quick else 1000, path))
    rows.append(run("cfg4 TF B=16 160/53", "tf", 16, 160, 53, 200 if quick else 800, path))
    rows.append(run("cfg4-weak TF B=128 160/53", "tf", 128, 160, 53, 200 if quick else 800, path))
    rows.append(run("cfg5 fwd TF B=64 160/53 train", "tf", 64, 160, 53, 200 if quick else 800, path, train=True))
json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "pb_probe.json"), "w"), indent=1)
