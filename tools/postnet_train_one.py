"""A few training-mode Postnet steps (forward + backward) for ncu launch lists.  usage: python tools/postnet_train_one.py [B] [T] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tacotron2_subword_b200 import create_hparams
from tacotron2_subword_b200.model import Postnet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 800
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(1)
net = Postnet(create_hparams()).cuda().train()
x = torch.randn(B, 80, T, device="cuda", requires_grad=True)
wgt = torch.randn(B, 80, T, device="cuda") / (B * T)
for _ in range(steps):
    net.zero_grad(set_to_none=True)
    x.grad = None
    (net(x) * wgt).sum().backward()
torch.cuda.synchronize()
print("ok")
