"""Whole-model inference latency (text ids -> mel_postnet) through the drop-in BERT_Tacotron2: reference-code encoder
(PyTorch) + CUDA decoder (latency path) + CUDA postnet.  usage: python tools/model_e2e.py"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tacotron2_subword_b200 import BERT_Tacotron2, create_hparams

torch.manual_seed(1234)
hp = create_hparams()
model = BERT_Tacotron2(hp).cuda().eval()
with torch.no_grad():
    model.decoder.gate_layer.linear_layer.bias.fill_(-20.0)      # never stops: exactly max_decoder_steps frames
model.decoder.max_decoder_steps = 1000
model.decoder.rng_seed = 1
T_in, T_sub = 150, 50
text = torch.randint(0, hp.n_symbols, (1, T_in)).cuda()
sub = torch.randint(0, hp.sub_n_symbols, (1, T_sub)).cuda()
pcls = torch.randn(1, T_in, hp.BERT_embedding_dim).cuda()
bcls = torch.randn(1, T_sub, hp.BERT_embedding_dim).cuda()


def run():
    with torch.no_grad():
        return model.inference(text, sub, pcls, bcls)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return sorted(ts)[len(ts) // 2]


import contextlib, io
with contextlib.redirect_stdout(io.StringIO()):
    total = timed(run)
    with torch.no_grad():
        mem, mem_s = model._memories(text, sub, pcls, bcls)
        enc = timed(lambda: model._memories(text, sub, pcls, bcls))
        dec = timed(lambda: model.decoder.inference(mem, mem_s))
        mel = model.decoder.inference(mem, mem_s)[0]
        post = timed(lambda: model.postnet.mel_postnet(mel))
        model.postnet.fused_eval = False
        post_torch = timed(lambda: model.postnet.mel_postnet(mel))
res = dict(workload="BERT_Tacotron2.inference, 150 phones + 50 sub-words, 1000 frames", total_ms_p50=round(total, 3),
           encoder_and_converters_ms=round(enc, 3), decoder_ms=round(dec, 3), postnet_cuda_ms=round(post, 3),
           postnet_pytorch_ms=round(post_torch, 3), frames_per_s=round(1000 / (total * 1e-3)))
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/model_e2e.json", "w"), indent=1)
