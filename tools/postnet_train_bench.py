"""Training-mode Postnet forward + backward: the repo's kernels (csrc/postnet_train.cuh) vs the reference formulation in PyTorch
(cuDNN, TF32 on = its default, and fp32).  usage: python tools/postnet_train_bench.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tacotron2_subword_b200 import create_hparams
from tacotron2_subword_b200.model import Postnet


def timed(fn, reps=5):
    fn(); fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


rows = []
for B, T in ((16, 800), (64, 800), (128, 800)):
    torch.manual_seed(1)
    net = Postnet(create_hparams()).cuda().train()
    ref = Postnet(create_hparams()).cuda().train(); ref.load_state_dict(net.state_dict()); ref.fused_train = False
    x = torch.randn(B, 80, T, device="cuda", requires_grad=True)
    wgt = torch.randn(B, 80, T, device="cuda") / (B * T)

    def step(m):
        m.zero_grad(set_to_none=True)
        x.grad = None
        (m(x) * wgt).sum().backward()

    r = dict(B=B, T=T, fused_ms=round(timed(lambda: step(net)), 3))
    torch.backends.cudnn.allow_tf32 = True
    r["torch_tf32_ms"] = round(timed(lambda: step(ref)), 3)
    torch.backends.cudnn.allow_tf32 = False
    r["torch_fp32_ms"] = round(timed(lambda: step(ref)), 3)
    torch.backends.cudnn.allow_tf32 = True
    r["frames_per_s_fused"] = round(B * T / (r["fused_ms"] * 1e-3))
    r["peak_mem_gb"] = round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)
    rows.append(r)
    print(json.dumps(r), flush=True)
json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "postnet_train_bench.json"), "w"), indent=1)
