"""-m gpu: eval-mode Postnet on tcgen05 (csrc/postnet.cuh) vs the reference golden, the CPU oracle and the PyTorch module.

Default precision: operands as split fp16 pairs (hi.hi + hi.lo + lo.hi), fp32 accumulation -> stated bound
max|delta| <= 1e-4 * max(1, max|postnet output|), an order of magnitude inside the 1e-3 mel bar (the tensor the
vocoder consumes must not be looser than the decoder mel).  ``fused_precision = "fp16"`` (plain fp16 operands, five layers
deep, linear last layer) is held to 3e-3 of the output scale in its own test."""
import os

import numpy as np
import pytest
import torch

from oracle.postnet_oracle import make_postnet_weights, mel_postnet, postnet_eval
from tacotron2_subword_b200 import create_hparams
from tacotron2_subword_b200.model import Postnet
from tests.helpers import GOLDEN_DIR

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(autouse=True)
def _exact_torch_convolutions():
    """The PyTorch module is used as a second reference in some tests: cuDNN's default TF32 convolutions are off by ~7e-3 of the
    output scale, far outside the bound tested here."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _close(got, want, residual=None):
    """|got - want| within TOL of the scale of the postnet's own output (want - residual)."""
    y = want if residual is None else want - residual
    scale = max(1.0, float(y.abs().max()))
    err = float((got - want).abs().max())
    assert err <= TOL * scale, (err, scale)


def _module(w):
    net = Postnet(create_hparams())
    sd = net.state_dict()
    for k in sd:
        if k in w:
            sd[k] = w[k]
    net.load_state_dict(sd, strict=True)
    return net.cuda().eval()


def test_postnet_matches_reference_golden():
    z = np.load(os.path.join(GOLDEN_DIR, "postnet_eval.npz"))
    net = _module(make_postnet_weights(int(z["seed"])))
    x = torch.from_numpy(z["x"]).cuda()
    with torch.no_grad():
        got = net.mel_postnet(x) - x
    _close(got.cpu(), torch.from_numpy(z["y"]))


@pytest.mark.parametrize("B,T", [(1, 1000), (1, 5), (3, 129), (16, 257), (64, 40)])
def test_postnet_vs_oracle_shapes_strides_and_mask(B, T):
    """Tile boundaries (T not a multiple of 128, utterances sharing a 128-position tile), the decoder's transposed storage
    as input, and the output-length mask."""
    seed = 5 + B
    w = make_postnet_weights(seed)
    net = _module(w)
    g = torch.Generator().manual_seed(seed)
    storage = torch.randn(B, T, 80, generator=g)                 # the decoder writes [B, T, n_mel]
    mel = storage.transpose(1, 2)                                # ... and hands out this view
    lens = torch.randint(max(1, T // 2), T + 1, (B,), generator=g)
    lens[0] = T
    want = mel_postnet(w, mel.contiguous(), lens)
    with torch.no_grad():
        got = net.mel_postnet(storage.cuda().transpose(1, 2), lens.cuda())
    assert got.shape == (B, 80, T) and got.is_contiguous()
    _close(got.cpu(), want, (mel.contiguous() * (torch.arange(T)[None, None, :] < lens[:, None, None])))
    for b in range(B):
        assert float(got[b, :, int(lens[b]):].abs().max() if int(lens[b]) < T else 0.0) == 0.0
    # the unfused module path (PyTorch ops) agrees as well, and is what training mode uses
    net.fused_eval = False
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # cuDNN's default TF32 convolutions are off by ~8e-3 here
    try:
        with torch.no_grad():
            ref = net.mel_postnet(mel.cuda(), lens.cuda())
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert float((ref.cpu() - want).abs().max()) <= 2e-4


def test_postnet_weights_are_repacked_after_an_update():
    w = make_postnet_weights(9)
    net = _module(w)
    x = torch.randn(2, 80, 50, generator=torch.Generator().manual_seed(1)).cuda()
    with torch.no_grad():
        a = net.mel_postnet(x)
        net.convolutions[2][0].conv.weight.mul_(0.5)             # in-place update bumps the tensor version
        b = net.mel_postnet(x)
        want = x + net.forward(x)
    assert float((a - b).abs().max()) > 1e-3
    _close(b, want, x)


def test_model_inference_uses_fused_postnet():
    from tacotron2_subword_b200 import BERT_Tacotron2
    torch.manual_seed(3)
    hp = create_hparams()
    model = BERT_Tacotron2(hp).cuda().eval()
    model.decoder.max_decoder_steps = 9
    model.decoder.rng_seed = 4
    text = torch.randint(0, hp.n_symbols, (1, 11)).cuda()
    sub = torch.randint(0, hp.sub_n_symbols, (1, 4)).cuda()
    pcls = torch.randn(1, 11, hp.BERT_embedding_dim).cuda()
    bcls = torch.randn(1, 4, hp.BERT_embedding_dim).cuda()
    with torch.no_grad():
        fused = model.inference(text, sub, pcls, bcls)
        model.postnet.fused_eval = False
        plain = model.inference(text, sub, pcls, bcls)
    assert torch.equal(fused[0], plain[0])                        # same decoder output (same seed)
    _close(fused[1], plain[1], plain[0])


def test_postnet_independent_rows_equal_batch1_on_truncated_input():
    """independent=True: frames beyond an utterance's length do not exist in ANY layer, so row b equals the oracle on
    mel[b, :, :len_b] alone -- whatever the padded frames contain (a free-running batch keeps decoding after a stop)."""
    seed, B, T = 17, 6, 150
    w = make_postnet_weights(seed)
    net = _module(w)
    g = torch.Generator().manual_seed(seed)
    mel = torch.randn(B, 80, T, generator=g)
    lens = torch.tensor([150, 149, 131, 128, 64, 3])
    with torch.no_grad():
        got = net.mel_postnet(mel.cuda(), lens.cuda(), independent=True).cpu()
        garbage = mel.clone()
        for b in range(B):
            garbage[b, :, int(lens[b]):] = 100.0 * torch.randn(80, T - int(lens[b]), generator=g)
        got2 = net.mel_postnet(garbage.cuda(), lens.cuda(), independent=True).cpu()
    assert torch.equal(got, got2)                                   # padded frames are never read
    for b in range(B):
        n = int(lens[b])
        want = mel_postnet(w, mel[b:b + 1, :, :n])
        _close(got[b:b + 1, :, :n], want, mel[b:b + 1, :, :n])
        assert float(got[b, :, n:].abs().max() if n < T else 0.0) == 0.0


def test_model_inference_batch_equals_per_utterance_inference():
    """BERT_Tacotron2.inference_batch vs one inference() call per utterance with the same replayed prenet masks:
    stop frames identical, mel / mel_postnet within the 16-bit-path bounds."""
    from tacotron2_subword_b200 import BERT_Tacotron2, DropoutReplay
    torch.manual_seed(7)
    hp = create_hparams()
    model = BERT_Tacotron2(hp).cuda().eval()
    steps, n = 12, 5
    with torch.no_grad():
        model.decoder.gate_layer.linear_layer.bias.fill_(-7.3)       # random-init gate logits straddle the stop threshold
    model.decoder.max_decoder_steps = steps
    model.decoder.weight_dtype = "fp16"
    g = torch.Generator().manual_seed(2)
    T_ins = [17, 9, 13, 17, 6]
    seqs = [torch.randint(0, hp.n_symbols, (1, t), generator=g).cuda() for t in T_ins]
    subs = [torch.randint(0, hp.sub_n_symbols, (1, max(2, t // 3)), generator=g).cuda() for t in T_ins]
    pcls = [torch.randn(1, t, hp.BERT_embedding_dim, generator=g).cuda() for t in T_ins]
    bcls = [torch.randn(1, s.shape[1], hp.BERT_embedding_dim, generator=g).cuda() for s in subs]
    keep = [[(torch.rand(steps, n, 256, generator=g) < 0.5).to(torch.uint8) for _ in range(2)] for _ in range(2)]
    model.decoder.dropout_replay = DropoutReplay(prenet_keep=keep)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        batch = model.inference_batch(seqs, subs, pcls, bcls)
        assert model.decoder._engine(torch.device("cuda", 0)).last_path() == "tensor"
        for i in range(n):
            model.decoder.dropout_replay = DropoutReplay(prenet_keep=[[m[:, i:i + 1] for m in row] for row in keep])
            with torch.no_grad():
                one = model.inference(seqs[i], subs[i], pcls[i], bcls[i])
            assert batch[i][0].shape == one[0].shape and batch[i][5] == one[5]
            assert float((batch[i][0] - one[0]).abs().max()) <= 1e-3
            _close(batch[i][1], one[1], one[0])
            assert float((batch[i][3] - one[3]).abs().max()) <= 2e-4 and float((batch[i][4] - one[4]).abs().max()) <= 2e-4
            assert batch[i][2].shape == one[2].shape
        lengths = [b_[0].shape[2] for b_ in batch]
        print("frames per utterance:", lengths)


def test_postnet_plain_fp16_mode_states_its_looser_bound():
    """fused_precision = "fp16": 2.5x less tensor work, error ~8e-4 of the output scale (bound 3e-3); switching modes re-packs."""
    w = make_postnet_weights(9)
    net = _module(w)
    x = torch.randn(3, 80, 129, generator=torch.Generator().manual_seed(4))
    want = mel_postnet(w, x)
    scale = max(1.0, float((want - x).abs().max()))
    with torch.no_grad():
        exact = net.mel_postnet(x.cuda()).cpu()
        net.fused_precision = "fp16"
        fast = net.mel_postnet(x.cuda()).cpu()
        net.fused_precision = "fp32"
        again = net.mel_postnet(x.cuda()).cpu()
    e_exact, e_fast = float((exact - want).abs().max()) / scale, float((fast - want).abs().max()) / scale
    print(f"postnet error / output scale: split-fp16 {e_exact:.1e}, plain fp16 {e_fast:.1e}")
    assert e_exact <= TOL and e_fast <= 3e-3 and e_fast > e_exact
    assert torch.equal(exact, again)
