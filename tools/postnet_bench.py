"""Eval-mode Postnet: the tcgen05 path (csrc/postnet.cuh) next to the reference module's PyTorch/cuDNN ops on the same GPU,
and both against the fp32 CPU oracle.  usage: python tools/postnet_bench.py"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.postnet_oracle import make_postnet_weights, mel_postnet
from tacotron2_subword_b200 import create_hparams
from tacotron2_subword_b200.model import Postnet

w = make_postnet_weights(1234)
net = Postnet(create_hparams())
sd = net.state_dict()
for k in sd:
    if k in w:
        sd[k] = w[k]
net.load_state_dict(sd)
net = net.cuda().eval()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps


rows = []
for B, T in [(1, 1000), (16, 800), (64, 1000), (128, 800)]:
    storage = torch.randn(B, T, 80, generator=torch.Generator().manual_seed(B)).cuda()
    mel = storage.transpose(1, 2)
    r = dict(B=B, T=T)
    with torch.no_grad():
        net.fused_eval = True
        r["fused_ms"] = round(timed(lambda: net.mel_postnet(mel)), 4)
        got = net.mel_postnet(mel)
        net.fused_eval = False
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            r["torch_tf32_ms" if tf32 else "torch_fp32_ms"] = round(timed(lambda: net.mel_postnet(mel)), 4)
            if B * T <= 20000:
                ref = net.mel_postnet(mel)
                want = mel_postnet(w, mel.cpu().contiguous())
                r["torch_tf32_err" if tf32 else "torch_fp32_err"] = float((ref.cpu() - want).abs().max())
        if B * T <= 20000:
            r["fused_err"] = float((got.cpu() - want).abs().max())
            r["out_scale"] = float((want - mel.cpu()).abs().max())
    r["frames_per_s_fused"] = round(B * T / (r["fused_ms"] * 1e-3))
    r["speedup_vs_torch_tf32"] = round(r["torch_tf32_ms"] / r["fused_ms"], 2)
    r["speedup_vs_torch_fp32"] = round(r["torch_fp32_ms"] / r["fused_ms"], 2)
    rows.append(r); print(json.dumps(r), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/postnet_bench.json", "w"), indent=1)
