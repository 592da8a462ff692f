"""BASELINE cfg 5, whole model: one BERT_Tacotron2 training step on one GPU: reference-code encoder in PyTorch; decoder forward +
backward, training-mode postnet forward + backward and the fused loss on the repo's kernels; SGD update.
usage: python tools/train_step_full.py [B] [T] [--torch-postnet-loss]   (the flag puts postnet and loss back on PyTorch ops)"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from tacotron2_subword_b200 import BERT_Tacotron2, create_hparams

from tacotron2_subword_b200.loss_function import Tacotron2Loss
args_ = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(args_[0]) if len(args_) > 0 else 64
T = int(args_[1]) if len(args_) > 1 else 800
torch_tail = "--torch-postnet-loss" in sys.argv
T_in, T_sub = 160, 53
torch.manual_seed(1234)
hp = create_hparams()
model = BERT_Tacotron2(hp).cuda().train()
model.decoder.weight_dtype = "fp16"
model.postnet.fused_train = not torch_tail
criterion = Tacotron2Loss()
criterion.fused = not torch_tail
g = torch.Generator().manual_seed(0)
in_len = torch.randint(T_in // 2, T_in + 1, (B,), generator=g); in_len[0] = T_in
in_len, _ = torch.sort(in_len, descending=True)                      # collate order (data_utils.py:146-160)
sub_len = torch.clamp(in_len // 3, min=2); sub_len[0] = T_sub
out_len = torch.randint(T // 2, T + 1, (B,), generator=g); out_len[0] = T
text = torch.randint(0, hp.n_symbols, (B, T_in), generator=g).cuda()
sub = torch.randint(0, hp.sub_n_symbols, (B, T_sub), generator=g).cuda()
mels = torch.randn(B, 80, T, generator=g).cuda()
gate_t = torch.zeros(B, T).cuda()
pcls = torch.randn(B, T_in, hp.BERT_embedding_dim, generator=g).cuda()
bcls = torch.randn(B, T_sub, hp.BERT_embedding_dim, generator=g).cuda()
x = (text, in_len.cuda(), sub_len.cuda(), mels, (T_in, T), out_len.cuda(), sub, pcls, bcls)
opt = torch.optim.SGD(model.parameters(), lr=1e-4)


def step():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    opt.zero_grad(set_to_none=True)
    ev[0].record()
    loss = criterion(model(x), (mels, gate_t, None), x, 50000)[0]
    ev[1].record()
    loss.backward()
    ev[2].record()
    opt.step()
    ev[3].record(); ev[3].synchronize()
    return ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), float(loss.detach())


step()
res = [step() for _ in range(3)]
fw, bw, up = (min(r[i] for r in res) for i in range(3))
out = dict(config=f"cfg5 whole-model training step B={B} T={T} {T_in}/{T_sub}", postnet_and_loss="PyTorch ops" if torch_tail else "repo kernels", forward_ms=round(fw, 2), backward_ms=round(bw, 2),
           optimizer_ms=round(up, 2), step_ms=round(fw + bw + up, 2), frames_per_s=round(B * T / ((fw + bw + up) * 1e-3)),
           loss_first=res[0][3], loss_last=res[-1][3], peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2**30, 2))
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/train_step_full%s.json" % ("_torch_tail" if torch_tail else ""), "w"), indent=1)
