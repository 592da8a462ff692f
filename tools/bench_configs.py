"""Throughput of the other BASELINE.json configs (parity-test cases, not the bench headline) on one GPU.
usage: python tools/bench_configs.py [quick]"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.synth import SMA, make_decoder_weights, make_inputs
from tacotron2_subword_b200 import Decoder, create_hparams

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
w = make_decoder_weights(SMA, seed=1234, gate_bias=-20.0)
hp = create_hparams()
dec = Decoder(hp); dec.load_state_dict(w); dec = dec.cuda().eval(); dec.rng_seed = 7
eng = dec._engine(torch.device("cuda", 0))


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


rows = []
def run(name, mode, B, T_in, T_sub, T, path, wdt):
    inp = make_inputs(B, T_in, T_sub, T if mode == "tf" else 1, seed=3, ragged=B > 1)
    mem, emb = inp["memory"].cuda(), inp["embeddings"].cuda()
    ml, bl = inp["memory_lengths"].cuda(), inp["bert_lengths"].cuda()
    dec.decoder_path, dec.weight_dtype = path, wdt
    if mode == "tf":
        mels = inp["mels"].cuda()
        fn = lambda: dec(mem, emb, mels, ml, bl)
    else:
        fn = lambda: dec.inference_batched(mem, emb, ml, bl, max_decoder_steps=T)
    with torch.no_grad():
        ms = timed(fn, 2 if quick else 3)
    r = dict(config=name, mode=mode, B=B, T_in=T_in, T_sub=T_sub, frames=T, path=eng.last_path(), weights=wdt,
             ms=round(ms, 3), us_per_frame_step=round(1e3 * ms / T, 2), frames_per_s=round(B * T / (ms * 1e-3)))
    rows.append(r); print(json.dumps(r), flush=True)

Tq = 100 if quick else None
run("cfg1 TF B=1 80/26 T=400", "tf", 1, 80, 26, 400, "auto", "fp32")
run("cfg1 TF B=1 80/26 T=400", "tf", 1, 80, 26, 400, "auto", "fp16")
run("cfg3 FR B=64 120/40", "fr", 64, 120, 40, Tq or 1000, "tensor", "fp16")
run("cfg3 FR B=64 120/40", "fr", 64, 120, 40, 50, "generic", "fp32")
run("cfg3-strong FR B=8 120/40 (8 GPUs x 8)", "fr", 8, 120, 40, Tq or 1000, "tensor", "fp16")
run("cfg3-strong FR B=8 120/40 generic fp32", "fr", 8, 120, 40, 200, "generic", "fp32")
run("cfg3-strong FR B=16 (4 GPUs x 16)", "fr", 16, 120, 40, Tq or 1000, "tensor", "fp16")
run("cfg3-strong FR B=32 (2 GPUs x 32)", "fr", 32, 120, 40, Tq or 1000, "tensor", "fp16")
run("cfg4 GTA TF B=16/GPU 160/53 T=800", "tf", 16, 160, 53, Tq or 800, "tensor", "fp16")
run("cfg4 GTA TF B=128 160/53 T=800 (weak)", "tf", 128, 160, 53, Tq or 800, "tensor", "fp16")
run("cfg4 GTA TF B=16 generic fp32", "tf", 16, 160, 53, 40, "generic", "fp32")
json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "bench_configs.json"), "w"), indent=1)
