"""-m gpu: decoder backward (BPTT kernels, tensor path) vs autograd through the CPU oracle.

The reference obtains these gradients from torch.autograd over its per-frame loop (model.py:392-428); the oracle
is the same arithmetic in fp32 on the CPU (pinned to the reference in tests/test_oracle_golden.py and
tests/test_host_cpu.py).  The CUDA path runs the recurrent products with fp16 (forward) / bf16 (backward) operands and
fp32 accumulation, so gradients are compared relative to each tensor's own scale:
    max|got - want| <= TOL_GRAD * max|want|        TOL_GRAD = 1e-2 (bf16 operands: 8-bit mantissa; measured <= 3e-3)
"""
import pytest
import torch

from oracle.decoder_oracle import DecoderOracle
from oracle.synth import SMA, DecoderDims, make_decoder_weights, make_dropout_plan, make_inputs
from tests.gpu_util import make_decoder, replay_of

pytestmark = pytest.mark.gpu

TOL_GRAD = 1e-2


def _loss(outs, seed):
    g = torch.Generator().manual_seed(seed)
    total = 0.0
    for o in outs:
        if o is None:
            continue
        wgt = torch.randn(o.shape, generator=g).to(o.device)
        total = total + (o * wgt).sum()
    return total


def _oracle_grads(w, inp, plan, training, dims=DecoderDims(), attention=SMA):
    orc = DecoderOracle(w, attention, dims=dims)
    orc.w = {k: v.clone().requires_grad_(True) for k, v in orc.w.items()}
    mem = inp["memory"].clone().requires_grad_(True)
    emb = inp["embeddings"].clone().requires_grad_(True) if dims.streams == 2 else None
    outs = orc.forward(mem, emb, inp["mels"], inp["memory_lengths"], inp["bert_lengths"] if emb is not None else None,
                       plan, training=training)
    _loss(outs, 5).backward()
    grads = {k: v.grad for k, v in orc.w.items()}
    return grads, mem.grad, (emb.grad if emb is not None else None), outs


@pytest.mark.parametrize("B,T,training,tf32,T_in,T_sub", [(3, 6, True, False, 24, 8), (16, 5, False, False, 24, 8),
                                                           (24, 7, True, False, 24, 8), (64, 4, True, True, 24, 8),
                                                           (100, 3, True, False, 300, 70),    # 128-column tiles, long memory
                                                           (17, 1, True, False, 9, 3)])       # a single frame
def test_backward_vs_oracle_autograd(B, T, training, tf32, T_in, T_sub):
    seed = 500 + B
    w = make_decoder_weights(SMA, seed=seed)
    inp = make_inputs(B, T_in, T_sub, T, seed=seed, ragged=True)
    plan = make_dropout_plan(B, T + 1, T, T_in, T_sub, training, seed=seed + 1)
    want, want_dmem, want_demb, want_outs = _oracle_grads(w, inp, plan, training)

    dec = make_decoder(w, SMA)
    dec.decoder_path, dec.weight_dtype = "tensor", "fp16"
    dec.grad_gemm_tf32 = tf32
    dec.dropout_replay = replay_of(plan)
    dec.train(training)
    mem = inp["memory"].cuda().requires_grad_(True)
    emb = inp["embeddings"].cuda().requires_grad_(True)
    outs = dec(mem, emb, inp["mels"].cuda(), inp["memory_lengths"].cuda(), inp["bert_lengths"].cuda())
    for o, wo in zip(outs, want_outs):
        assert (o.detach().cpu() - wo.detach()).abs().max() < 1e-3
    _loss(outs, 5).backward()
    torch.cuda.synchronize()
    sd = dict(dec.named_parameters())
    worst = {}
    for name, gw in want.items():
        got = sd[name].grad
        if gw is None:
            assert got is None, f"{name}: dead parameter must not receive a gradient"
            continue
        assert got is not None, f"{name}: no gradient"
        scale = float(gw.abs().max())
        if scale == 0.0:        # e.g. a single frame: the recurrent inputs are the zero initial state
            assert float(got.abs().max()) == 0.0, f"{name}: gradient must be exactly zero"
            continue
        worst[name] = float((got.cpu() - gw).abs().max()) / scale
    for name, gw, got in (("memory", want_dmem, mem.grad), ("embeddings", want_demb, emb.grad)):
        worst[name] = float((got.cpu() - gw).abs().max() / gw.abs().max())
    bad = {k: v for k, v in worst.items() if not v < TOL_GRAD}
    print({k: f"{v:.2e}" for k, v in worst.items()})
    assert not bad, f"relative gradient error above {TOL_GRAD}: {bad}"

@pytest.mark.parametrize("B,rows,training", [(1, 128, True), (1, 128, False), (11, 4, True), (130, 128, False)])
def test_backward_batch1_and_sub_batches(B, rows, training):
    """Batches outside one BPTT call (train.py:330 trains whatever collate_fn emits, data_utils.py:146-160): B = 1 runs as a
    padded pair, B > max_backward_rows as balanced sub-batches over the same padded memory; outputs and every gradient must
    equal autograd through the oracle on the whole batch (replayed masks sliced along the batch)."""
    T, T_in, T_sub, seed = 4, 22, 7, 700 + B
    w = make_decoder_weights(SMA, seed=seed)
    inp = make_inputs(B, T_in, T_sub, T, seed=seed, ragged=B > 1)
    plan = make_dropout_plan(B, T + 1, T, T_in, T_sub, training, seed=seed + 1)
    want, want_dmem, want_demb, want_outs = _oracle_grads(w, inp, plan, training)
    dec = make_decoder(w, SMA, exact=False)
    dec.max_backward_rows = rows
    dec.dropout_replay = replay_of(plan)
    dec.train(training)
    mem = inp["memory"].cuda().requires_grad_(True)
    emb = inp["embeddings"].cuda().requires_grad_(True)
    outs = dec(mem, emb, inp["mels"].cuda(), inp["memory_lengths"].cuda(), inp["bert_lengths"].cuda())
    for o, wo in zip(outs, want_outs):
        assert o.shape == wo.shape
        assert (o.detach().cpu() - wo.detach()).abs().max() < 1e-3
    _loss(outs, 5).backward()
    torch.cuda.synchronize()
    sd = dict(dec.named_parameters())
    worst = {}
    for name, gw in want.items():
        got = sd[name].grad
        if gw is None:
            assert got is None, name
            continue
        assert got is not None, f"{name}: no gradient"
        worst[name] = float((got.cpu() - gw).abs().max()) / float(gw.abs().max())
    for name, gw, got in (("memory", want_dmem, mem.grad), ("embeddings", want_demb, emb.grad)):
        assert got.shape == gw.shape
        worst[name] = float((got.cpu() - gw).abs().max() / gw.abs().max())
    bad = {k: v for k, v in worst.items() if not v < TOL_GRAD}
    assert not bad, f"relative gradient error above {TOL_GRAD}: {bad}"
    assert dec.dropout_replay is not None and dec.validate_lengths and dec.max_backward_rows == rows     # state restored


def test_backward_philox_masks_match_forward():
    """Production mode (no replay): backward re-draws the LSTM-state dropout masks from Philox.  Materialising the
    same masks through taco2dec_philox_keep_mask and replaying them must give the same outputs and gradients."""
    import ctypes as C
    from tacotron2_subword_b200 import DropoutReplay, _cabi
    B, T, T_in, T_sub, seed, H = 16, 4, 20, 7, 91, 1024
    w = make_decoder_weights(SMA, seed=seed)
    inp = make_inputs(B, T_in, T_sub, T, seed=seed, ragged=True)
    dec = make_decoder(w, SMA).train()
    dec.decoder_path, dec.weight_dtype, dec.rng_seed = "tensor", "fp16", 1234
    args = (inp["memory"].cuda(), inp["embeddings"].cuda(), inp["mels"].cuda(), inp["memory_lengths"].cuda(),
            inp["bert_lengths"].cuda())
    outs = dec(*args)
    _loss(outs, 3).backward()
    g1 = {n: p.grad.clone() for n, p in dec.named_parameters() if p.grad is not None}
    assert len(g1) == 26
    dec.zero_grad(set_to_none=True)

    lib = _cabi.load_library()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    pk = [[torch.empty(T + 1, B, 256, dtype=torch.uint8, device="cuda") for _ in range(2)] for _ in range(2)]
    for s_ in range(2):
        for l in range(2):
            _cabi.check(lib.taco2dec_philox_keep_mask(1234, s_ * 2 + l, T + 1, B * 256, 0.5, C.c_void_p(pk[s_][l].data_ptr()), st))
    lk = torch.empty(6, T, B, H, dtype=torch.uint8, device="cuda")
    for k in range(6):
        _cabi.check(lib.taco2dec_philox_keep_mask(1234, 4 + k, T, B * H, 0.1, C.c_void_p(lk[k].data_ptr()), st))
    dec.dropout_replay = DropoutReplay(prenet_keep=pk, lstm_keep=lk.permute(1, 0, 2, 3).contiguous())
    outs2 = dec(*args)
    for o, o2 in zip(outs, outs2):
        assert torch.equal(o, o2)
    _loss(outs2, 3).backward()
    for n, p in dec.named_parameters():
        if p.grad is not None:
            assert float((p.grad - g1[n]).abs().max()) <= 1e-3 * float(g1[n].abs().max()), n   # atomics reorder fp32 sums; bf16 rounding of the gate gradients amplifies it


def test_backward_unsupported_shape_raises():
    """A forced latency / generic path has no backward kernel: the output must refuse backward() loudly."""
    for B, path in ((1, "latency"), (4, "generic")):
        T, T_in, T_sub, seed = 3, 12, 4, 7
        w = make_decoder_weights(SMA, seed=seed)
        inp = make_inputs(B, T_in, T_sub, T, seed=seed)
        dec = make_decoder(w, SMA).train()
        dec.decoder_path = path
        outs = dec(inp["memory"].cuda(), inp["embeddings"].cuda(), inp["mels"].cuda(), inp["memory_lengths"].cuda(),
                   inp["bert_lengths"].cuda())
        with pytest.raises(NotImplementedError):
            outs[0].sum().backward()


@pytest.mark.parametrize("B,T,training,T_in,T_sub", [(4, 5, False, 24, 8), (16, 4, True, 40, 13), (33, 3, True, 70, 20)])
def test_lsa_backward_vs_oracle_autograd(B, T, training, T_in, T_sub):
    """Location-sensitive attention (attention.py:7-85): softmax, the location conv/dense layers and the cumulative
    weights are differentiated by bw_attention_lsa; all 30 live parameter gradients + d memory / d embeddings vs autograd
    through the oracle."""
    from oracle.synth import LSA
    seed = 700 + B
    w = make_decoder_weights(LSA, seed=seed)
    inp = make_inputs(B, T_in, T_sub, T, seed=seed, ragged=True)
    plan = make_dropout_plan(B, T + 1, T, T_in, T_sub, training, seed=seed + 1)
    want, want_dmem, want_demb, want_outs = _oracle_grads(w, inp, plan, training, attention=LSA)
    dec = make_decoder(w, LSA)
    dec.decoder_path, dec.weight_dtype = "tensor", "fp16"
    dec.dropout_replay = replay_of(plan)
    dec.train(training)
    mem = inp["memory"].cuda().requires_grad_(True)
    emb = inp["embeddings"].cuda().requires_grad_(True)
    outs = dec(mem, emb, inp["mels"].cuda(), inp["memory_lengths"].cuda(), inp["bert_lengths"].cuda())
    for o, wo in zip(outs, want_outs):
        assert (o.detach().cpu() - wo.detach()).abs().max() < 1e-3
    _loss(outs, 5).backward()
    sd = dict(dec.named_parameters())
    worst = {}
    n_live = 0
    for name, gw in want.items():
        got = sd[name].grad
        if gw is None:
            assert got is None, name
            continue
        n_live += 1
        assert got is not None, f"{name}: no gradient"
        worst[name] = float((got.cpu() - gw).abs().max() / gw.abs().max())
    for name, gw, got in (("memory", want_dmem, mem.grad), ("embeddings", want_demb, emb.grad)):
        worst[name] = float((got.cpu() - gw).abs().max() / gw.abs().max())
    print({k: f"{v:.1e}" for k, v in worst.items() if "location" in k or "attention_layer" in k})
    assert n_live == 30
    bad = {k: v for k, v in worst.items() if not v < TOL_GRAD}
    assert not bad, f"relative gradient error above {TOL_GRAD}: {bad}"


def test_model_training_step_end_to_end():
    """BERT_Tacotron2 drop-in: forward -> Tacotron2Loss-style loss (loss_function.py:12-66, L2 alignment variant) ->
    backward -> optimizer step -> forward again (the engine must pick up the updated weights)."""
    from tacotron2_subword_b200 import BERT_Tacotron2, create_hparams
    torch.manual_seed(1234)
    hp = create_hparams()
    model = BERT_Tacotron2(hp).cuda().train()
    model.decoder.weight_dtype = "fp16"
    model.decoder.rng_seed = 11
    B, T_in, T_sub, T = 16, 14, 6, 9
    g = torch.Generator().manual_seed(3)
    text = torch.randint(0, hp.n_symbols, (B, T_in), generator=g).cuda()
    sub = torch.randint(0, hp.sub_n_symbols, (B, T_sub), generator=g).cuda()
    in_len = torch.randint(T_in // 2, T_in + 1, (B,), generator=g); in_len[0] = T_in
    sub_len = torch.randint(T_sub // 2, T_sub + 1, (B,), generator=g); sub_len[0] = T_sub
    out_len = torch.randint(T // 2, T + 1, (B,), generator=g); out_len[0] = T
    order = torch.argsort(in_len, descending=True)          # the encoder packs by length (data_utils.py:146-160)
    in_len, sub_len, out_len = in_len[order].cuda(), sub_len[order].cuda(), out_len[order].cuda()
    mels = torch.randn(B, 80, T, generator=g).cuda()
    gate_t = torch.zeros(B, T).cuda()
    align_t = torch.rand(B, T, T_in, generator=g).cuda()
    pcls = torch.randn(B, T_in, hp.BERT_embedding_dim, generator=g).cuda()
    bcls = torch.randn(B, T_sub, hp.BERT_embedding_dim, generator=g).cuda()
    x = (text, in_len, sub_len, mels, (T_in, T), out_len, sub, pcls, bcls)
    opt = torch.optim.SGD(model.parameters(), lr=1e-2)

    def loss_of(out):
        mel, mel_post, gate, al, alb = out
        F = torch.nn.functional
        return (F.mse_loss(mel, mels) + F.mse_loss(mel_post, mels) +
                F.binary_cross_entropy_with_logits(gate.reshape(-1, 1), gate_t.reshape(-1, 1)) + F.mse_loss(al, align_t))

    loss0 = loss_of(model(x))
    opt.zero_grad(set_to_none=True)
    loss0.backward()
    dead = 0
    for n, p in model.named_parameters():
        if "decoder_rnn_bert" in n:
            assert p.grad is None
            dead += 1
            continue
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    assert dead == 4
    assert float(model.decoder.attention_rnn.weight_hh.grad.abs().max()) > 0
    assert float(model.embedding.weight.grad.abs().max()) > 0          # gradient reaches the encoder through d memory
    opt.step()
    loss1 = loss_of(model(x))
    assert torch.isfinite(loss1)
    assert float(loss1.detach()) < float(loss0.detach())       # same dropout seed, small SGD step: the loss goes down


@pytest.mark.parametrize("name", ["grad_sma_train_B16", "grad_sma_eval_B3", "grad_lsa_train_B4"])
def test_backward_vs_reference_gradient_digests(name):
    """CUDA backward against digests of the REFERENCE's own autograd (tests/golden/grad_*.npz, made by
    oracle/make_golden.py from the unmodified reference): max|g|, a seeded random projection, the sum and the first 32
    elements of every gradient tensor, each within TOL_GRAD of max|g_ref|."""
    from oracle.synth import seeded_loss
    from tests.helpers import check_grad_digest, load_grad_golden, materialise
    recipe, digests, _ = load_grad_golden(name)
    w, inp, plan = materialise(recipe)
    dec = make_decoder(w, recipe["attention"])
    dec.decoder_path, dec.weight_dtype = "tensor", "fp16"
    dec.dropout_replay = replay_of(plan)
    dec.train(recipe["training"])
    mem = inp["memory"].cuda().requires_grad_(True)
    emb = inp["embeddings"].cuda().requires_grad_(True)
    outs = dec(mem, emb, inp["mels"].cuda(), inp["memory_lengths"].cuda(), inp["bert_lengths"].cuda())
    seeded_loss(outs, recipe["loss_seed"]).backward()
    grads = {n: p.grad for n, p in dec.named_parameters()}
    grads["memory"], grads["embeddings"] = mem.grad, emb.grad
    worst = {}
    for n, want in digests.items():
        if want is None:
            assert grads[n] is None, n
        else:
            worst[n] = check_grad_digest(n, grads[n], want, rtol=TOL_GRAD)
    print({k: f"{v:.1e}" for k, v in worst.items()})


@pytest.mark.parametrize("B,T,T_in,T_sub", [(16, 12, 40, 13), (33, 5, 150, 50), (64, 3, 24, 8), (2, 9, 17, 5)])
def test_persistent_backward_equals_per_frame_graph(B, T, T_in, T_sub, monkeypatch):
    """The persistent BPTT kernel (csrc/persist_bwd.cuh: one launch for the whole reverse-time loop, SMA, B <= 64) and the per-frame
    graph of csrc/backward.cuh compute the same arithmetic on the same operands: every gradient must agree to fp32 summation
    order (the two differ only in the order of split-K and atomic additions)."""
    seed = 900 + B
    w = make_decoder_weights(SMA, seed=seed)
    inp = make_inputs(B, T_in, T_sub, T, seed=seed, ragged=True)
    plan = make_dropout_plan(B, T + 1, T, T_in, T_sub, True, seed=seed + 1)
    grads, counts = [], []
    for no_persist in (False, True):
        if no_persist:
            monkeypatch.setenv("TACO2DEC_NO_PERSIST_BWD", "1")
        else:
            monkeypatch.delenv("TACO2DEC_NO_PERSIST_BWD", raising=False)
        dec = make_decoder(w, SMA, exact=False).train()
        dec.dropout_replay = replay_of(plan)
        mem = inp["memory"].cuda().requires_grad_(True)
        emb = inp["embeddings"].cuda().requires_grad_(True)
        eng = dec._engine(torch.device("cuda", 0))
        n0 = eng.launch_count()
        outs = dec(mem, emb, inp["mels"].cuda(), inp["memory_lengths"].cuda(), inp["bert_lengths"].cuda())
        n1 = eng.launch_count()
        _loss(outs, 5).backward()
        torch.cuda.synchronize()
        dec.check()
        counts.append(eng.launch_count() - n1)
        g = {n: p.grad.clone() for n, p in dec.named_parameters() if p.grad is not None}
        g["memory"], g["embeddings"] = mem.grad.clone(), emb.grad.clone()
        grads.append(g)
    # the per-frame graph replays 5 kernels per frame (+ 3 around the loop); the persistent path is the projection gradient, four
    # fp16 copies of the attention operands and ONE launch for the whole reverse-time loop
    assert counts[1] - counts[0] == 5 * T + 3 - 6, counts
    assert grads[0].keys() == grads[1].keys() and len(grads[0]) == 28
    for n in grads[0]:
        scale = float(grads[1][n].abs().max())
        assert float((grads[0][n] - grads[1][n]).abs().max()) <= 2e-3 * scale, n
