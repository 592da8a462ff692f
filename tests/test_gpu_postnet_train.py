"""-m gpu: training-mode Postnet (csrc/postnet_train.cuh: tcgen05 contractions + fused BatchNorm / tanh / dropout, hand-written
backward) vs the reference formulation in PyTorch fp32 (/root/reference/model.py:27-70 under model.train(): Conv1d ->
BatchNorm1d with batch statistics -> tanh -> F.dropout(0.5), gradients from autograd) on identical dropout masks.

The CUDA path multiplies fp16 operands with fp32 accumulation (the grade of cuDNN's default TF32 convolutions), so the bounds are
relative to each tensor's own scale:  outputs <= 3e-3 * max|out|,  gradients <= 1e-2 * max|g|  (measured ~5e-4 / ~2e-3)."""
import pytest
import torch
import torch.nn.functional as F

from tacotron2_subword_b200 import create_hparams
from tacotron2_subword_b200.model import Postnet

pytestmark = pytest.mark.gpu

TOL_OUT, TOL_GRAD = 3e-3, 1e-2


@pytest.fixture(autouse=True)
def _fp32_reference():
    """The PyTorch side is the fp32 reference: no TF32 in cuDNN while these tests run."""
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = prev


def _nets(seed):
    torch.manual_seed(seed)
    net = Postnet(create_hparams()).cuda().train()
    with torch.no_grad():
        for seq in net.convolutions:            # non-trivial BatchNorm affine parameters
            seq[1].weight.uniform_(0.5, 1.5)
            seq[1].bias.uniform_(-0.3, 0.3)
    ref = Postnet(create_hparams()).cuda().train()
    ref.load_state_dict(net.state_dict())
    ref.fused_train = False
    return net, ref


def _reference_forward(ref, x, masks):
    """model.py:62-70 with the dropout masks made explicit ([B*T, C] uint8 per layer, channel-last rows)."""
    B, _, T = x.shape
    last = len(ref.convolutions) - 1
    for i, conv in enumerate(ref.convolutions):
        x = conv(x)
        if i < last:
            x = torch.tanh(x)
        m = masks[i].view(B, T, -1).transpose(1, 2).to(x.dtype)
        x = x * m * 2.0
    return x


def _masks(net, B, T, seed):
    g = torch.Generator().manual_seed(seed)
    chans = [seq[0].conv.weight.shape[0] for seq in net.convolutions]
    return [(torch.rand(B * T, c, generator=g) >= 0.5).to(torch.uint8).cuda() for c in chans]


@pytest.mark.parametrize("B,T", [(3, 37), (1, 5), (16, 129), (2, 300)])
def test_training_postnet_forward_backward_vs_pytorch(B, T):
    if True:
        net, ref = _nets(5 + B)
        masks = _masks(net, B, T, 11)
        net.dropout_replay = masks
        x = (torch.randn(B, 80, T, generator=torch.Generator().manual_seed(3)) * 2.0 - 3.0).cuda()
        wgt = torch.randn(B, 80, T, generator=torch.Generator().manual_seed(4)).cuda() / (B * T)   # tiny upstream gradients, as a mean loss gives
        x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        out = net(x1)
        want = _reference_forward(ref, x2, masks)
        assert out.shape == want.shape
        scale = float(want.abs().max())
        assert float((out - want).abs().max()) <= TOL_OUT * scale
        (out * wgt).sum().backward()
        (want * wgt).sum().backward()
        torch.cuda.synchronize()
        worst = {"input": float((x1.grad - x2.grad).abs().max() / x2.grad.abs().max())}
        wscale = 0.0
        for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
            assert p.grad is not None, n
            if n.endswith("conv.bias"):
                continue
            worst[n] = float((p.grad - q.grad).abs().max() / q.grad.abs().max())
            wscale = max(wscale, float(q.grad.abs().max()))
        bad = {k: v for k, v in worst.items() if not v < TOL_GRAD}
        assert not bad, bad
        for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
            if n.endswith("conv.bias"):     # exactly zero in exact arithmetic (BatchNorm removes the mean); autograd leaves rounding noise
                assert float(p.grad.abs().max()) == 0.0 and float(q.grad.abs().max()) <= 1e-3 * wscale, n
        # running statistics follow nn.BatchNorm1d's update (momentum 0.1, unbiased variance)
        for seq_n, seq_r in zip(net.convolutions, ref.convolutions):
            bn_n, bn_r = seq_n[1], seq_r[1]
            assert int(bn_n.num_batches_tracked) == int(bn_r.num_batches_tracked) == 1
            assert float((bn_n.running_mean - bn_r.running_mean).abs().max()) <= 2e-3 * max(1.0, float(bn_r.running_mean.abs().max()))
            assert float((bn_n.running_var - bn_r.running_var).abs().max()) <= 5e-3 * float(bn_r.running_var.abs().max())


def test_training_postnet_philox_masks_are_reproducible_and_backward_uses_them():
    """Production mode (no replay): masks are drawn in-kernel from Philox; backward must re-draw the SAME masks.  Checked with a
    directional derivative under a fixed seed (the function is then deterministic)."""
    net, _ = _nets(21)
    net.rng_seed = 777
    B, T = 4, 64
    x = torch.randn(B, 80, T, generator=torch.Generator().manual_seed(8)).cuda()
    wgt = torch.randn(B, 80, T, generator=torch.Generator().manual_seed(9)).cuda()
    f = lambda inp: float((net(inp) * wgt).sum())
    with torch.no_grad():
        o1, o2 = net(x), net(x)
    assert torch.equal(o1, o2)
    frac = float((o1 == 0).float().mean())
    assert 0.4 < frac < 0.6, frac                  # dropout(0.5) on the last layer
    net.rng_seed = 778
    with torch.no_grad():
        assert not torch.equal(net(x), o1)
    net.rng_seed = 777
    xg = x.clone().requires_grad_(True)
    (net(xg) * wgt).sum().backward()
    v = torch.randn(x.shape, generator=torch.Generator().manual_seed(10)).cuda()
    eps = 2e-2
    with torch.no_grad():
        fd = (f(x + eps * v) - f(x - eps * v)) / (2 * eps)
    an = float((xg.grad * v).sum())
    assert abs(fd - an) <= 0.05 * max(abs(an), 1.0), (fd, an)


def test_training_postnet_inside_the_model_and_eval_switch():
    """mel_postnet in train() goes through the CUDA training path; eval() switches to the folded-BatchNorm path."""
    net, ref = _nets(31)
    B, T = 2, 50
    masks = _masks(net, B, T, 5)
    net.dropout_replay = masks
    mel = torch.randn(B, 80, T, generator=torch.Generator().manual_seed(1)).cuda()
    got = net.mel_postnet(mel)
    want = mel + _reference_forward(ref, mel, masks)
    assert float((got - want).abs().max()) <= TOL_OUT * float(want.abs().max())
    net.eval(); ref.eval()
    net.dropout_replay = None
    with torch.no_grad():
        e1, e2 = net.mel_postnet(mel), mel + ref(mel)
    assert float((e1 - e2).abs().max()) <= 1e-3 * float(e2.abs().max())
