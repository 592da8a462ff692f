"""Short tensor-path run for ncu launch lists: python tools/tensor_probe.py [B] [T] [fr|tf]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.synth import SMA, make_decoder_weights, make_inputs
from tacotron2_subword_b200 import Decoder, create_hparams
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 12
mode = sys.argv[3] if len(sys.argv) > 3 else "fr"
w = make_decoder_weights(SMA, seed=1234, gate_bias=-20.0)
dec = Decoder(create_hparams()); dec.load_state_dict(w); dec = dec.cuda().eval(); dec.rng_seed = 7
dec.decoder_path, dec.weight_dtype = "tensor", "fp16"
inp = make_inputs(B, 120, 40, T, seed=3, ragged=True)
a = [inp[k].cuda() for k in ("memory", "embeddings", "mels", "memory_lengths", "bert_lengths")]
with torch.no_grad():
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if mode == "tf": dec(a[0], a[1], a[2], a[3], a[4])
        else: dec.inference_batched(a[0], a[1], a[3], a[4], max_decoder_steps=T)
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1)
print(f"ok B={B} T={T} {mode}: {ms:.2f} ms, {1e3 * ms / T:.1f} us per frame-step, {B * T / (ms * 1e-3):.0f} frames/s")
