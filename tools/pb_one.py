"""One batched decoder call, repeated (for ncu captures).  usage: python tools/pb_one.py fr|tf B T_in T_sub T [path] [reps] [train]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.synth import SMA, make_decoder_weights, make_inputs
from tacotron2_subword_b200 import Decoder, create_hparams

mode, B, T_in, T_sub, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
path = sys.argv[6] if len(sys.argv) > 6 else "tensor"
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
w = make_decoder_weights(SMA, seed=1234, gate_bias=-20.0)
dec = Decoder(create_hparams()); dec.load_state_dict(w); dec = dec.cuda().eval(); dec.rng_seed = 7
dec.decoder_path = path
dec.train("train" in sys.argv)
eng = dec._engine(torch.device("cuda", 0)); eng.set_profiling(True)
inp = make_inputs(B, T_in, T_sub, T if mode == "tf" else 1, seed=3, ragged=B > 1)
mem, emb = inp["memory"].cuda(), inp["embeddings"].cuda()
ml, bl = inp["memory_lengths"].cuda(), inp["bert_lengths"].cuda()
mels = inp["mels"].cuda() if mode == "tf" else None
with torch.no_grad():
    for _ in range(reps):
        if mode == "tf":
            dec(mem, emb, mels, ml, bl)
        else:
            dec.inference_batched(mem, emb, ml, bl, max_decoder_steps=T)
        torch.cuda.synchronize()
        print(f"{eng.last_path()} kernel {1e3 * eng.last_kernel_ms() / T:.2f} us/frame", flush=True)
dec.check()
