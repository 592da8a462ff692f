"""-m gpu: parity at the horizons of the BASELINE.json configurations (the other GPU tests use 3-14 frames).

  cfg 2  free-running B=1, 150 phones / 50 sub-words, ALL 1000 frames against the CPU oracle (latency fp32 / fp16, generic)
  cfg 3  free-running B=64, 120 / 40, 400 frames, per-utterance stop frames, tensor path (fp16 operands)
  cfg 4  teacher-forced B=16 (= 128 utterances over 8 GPUs), 160 / 53, 800 frames, tensor path
  cfg 5  decoder backward (bf16 operands) at T = 50 / 200 frames, per-tensor relative error printed against T

Stated bounds: fp32 kernels mel / gate <= 1e-4, alignments <= 1e-5; 16-bit modes mel / gate <= 1e-3, alignments <= 2e-4
(north_star: mel <= 1e-3); stop frames / flags exact whenever the oracle's gate logits are at least 2e-3 away from the
threshold ("safe" utterances), and the fraction of ALL utterances with a matching stop frame is reported (>= 90 %).  Gradients:
max|got - want| <= GRAD_BOUND(T) * max|want| per tensor, GRAD_BOUND(T) = 1e-2 for T <= 200 (measured on a B200: 3.0e-3 at
B=16/T=50, 3.5e-3 at B=16/T=200, 5.3e-3 at B=64/T=200 -- the error does NOT grow with the horizon: the bf16 rounding of the gate
gradients is re-drawn every frame and averages out in the sums over frames), 1e-2 * sqrt(T / 200) beyond that.

The oracle for the batched free-running case is a batched loop over ``DecoderOracle._decode(..., truncate=True)`` (the
per-utterance definition costs 64 x 400 batch-1 frames); it is checked against the per-utterance definition on two
utterances inside the test."""
import math

import pytest
import torch

from oracle.decoder_oracle import DecoderOracle, lengths_to_mask
from oracle.synth import SMA, make_decoder_weights, make_dropout_plan, make_inputs
from tests.gpu_util import make_decoder, replay_of
from tests.helpers import maxabs

pytestmark = pytest.mark.gpu

LOGIT_THR = math.log(0.001 / 0.999)        # sigmoid(g) > 0.001  <=>  g > LOGIT_THR  (hparams.gate_threshold, model.py:480)


def grad_bound(T: int) -> float:
    return 1e-2 * max(1.0, math.sqrt(T / 200.0))


def _cmp(got, want, tol_mel, tol_align, tag):
    worst = {}
    for n, t, g, w in zip(("mel", "gate", "align", "align_bert"), (tol_mel, tol_mel, tol_align, tol_align), got, want):
        g = g.detach().float().cpu()
        assert g.shape == w.shape, (tag, n, g.shape, w.shape)
        assert torch.isfinite(g).all(), (tag, n)
        worst[n] = maxabs(g, w)
        assert worst[n] <= t, (tag, n, worst[n])
    return worst


def _first_divergence(a, b, tol):
    """First frame (dim 1 of [B, T, ...]) at which any element differs by more than tol; None if never."""
    d = (a.double() - b.double()).abs()
    while d.dim() > 2:
        d = d.amax(-1)
    bad = (d > tol).any(0).nonzero()
    return int(bad[0]) if bad.numel() else None


# ---------------------------------------------------------------------------------------------------------------
# cfg 2: the bench's own configuration, every one of its 1000 frames against the oracle
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cfg2_oracle():
    T_in, T_sub, steps, seed = 150, 50, 1000, 1234
    w = make_decoder_weights(SMA, seed=seed, gate_bias=-20.0)
    inp = make_inputs(1, T_in, T_sub, 1, seed=seed)
    plan = make_dropout_plan(1, steps + 1, steps, T_in, T_sub, False, seed=seed + 1)
    with torch.no_grad():
        want = DecoderOracle(w, SMA).inference(inp["memory"], inp["embeddings"], plan, max_decoder_steps=steps)
    return w, inp, plan, want


@pytest.mark.parametrize("path,wdtype,tol_mel,tol_align", [("latency", "fp32", 1e-4, 1e-5), ("latency", "fp16", 1e-3, 2e-4),
                                                           ("generic", "fp32", 1e-4, 1e-5)])
def test_cfg2_all_1000_frames_vs_oracle(cfg2_oracle, path, wdtype, tol_mel, tol_align):
    w, inp, plan, (omel, ogate, oal, oalb, oflag) = cfg2_oracle
    dec = make_decoder(w, SMA).eval()
    dec.decoder_path, dec.weight_dtype = path, wdtype
    dec.max_decoder_steps = 1000
    dec.dropout_replay = replay_of(plan)
    with torch.no_grad():
        mel, gate, al, alb, flag = dec.inference(inp["memory"].cuda(), inp["embeddings"].cuda())
    assert dec._engine(torch.device("cuda", 0)).last_path() == path
    assert mel.shape == (1, 80, 1000) == tuple(omel.shape) and flag is False and oflag is False
    worst = _cmp((mel, gate, al, alb), (omel, ogate, oal, oalb), tol_mel, tol_align, f"cfg2 {path}/{wdtype}")
    # the error must not grow along the utterance: last 100 frames no worse than 4x the first 100 (+ float noise)
    head = maxabs(mel[:, :, :100].cpu(), omel[:, :, :100])
    tail = maxabs(mel[:, :, 900:].cpu(), omel[:, :, 900:])
    print(f"cfg2 {path}/{wdtype}: worst {worst}, mel error frames 0-99 {head:.2e}, frames 900-999 {tail:.2e}")
    assert tail <= 4 * head + 0.1 * tol_mel


# ---------------------------------------------------------------------------------------------------------------
# cfg 3: batched free-running, B=64, 120 phones / 40 sub-words, 400 frames, mixed stop frames
# ---------------------------------------------------------------------------------------------------------------
def _oracle_free_running_batched(orc, inp, plan, steps):
    """Never-stopping batched free run of the oracle with the independent-utterance rule (positions >= length do not
    exist: truncate=True).  Returns mel [B, steps, 80], gate [B, steps], align [B, steps, T_in], align_bert."""
    mems = [inp["memory"], inp["embeddings"]]
    lens = [inp["memory_lengths"], inp["bert_lengths"]]
    B = mems[0].shape[0]
    streams = [orc._init_stream(s, m, ~lengths_to_mask(l)) for s, m, l in zip(orc.sfx, mems, lens)]
    h2 = torch.zeros(B, orc.d.drnn)
    c2 = torch.zeros(B, orc.d.drnn)
    x = torch.zeros(B, orc.d.n_mel)
    mels, gates, aligns = [], [], [[], []]
    for t in range(steps):
        pre = [orc._prenet(s, x, plan.prenet_keep[i][0][t], plan.prenet_keep[i][1][t]) for i, s in enumerate(orc.sfx)]
        mel, gate, h2, c2 = orc._decode(streams, pre, h2, c2, None, None, True)
        mels.append(mel)
        gates.append(gate.squeeze(1))
        for a, s in zip(aligns, streams):
            a.append(s.a_prev)
        x = mel
    st = lambda xs: torch.stack(xs).transpose(0, 1).contiguous()
    return st(mels), st(gates), st(aligns[0]), st(aligns[1])


def _stops(G, c, steps):
    """Per-utterance frame count for gate threshold c on zero-bias logits G [B, steps] (stop frame included)."""
    out = []
    for b in range(G.shape[0]):
        hit = (G[b] > c).nonzero()
        out.append(int(hit[0]) + 1 if hit.numel() else steps)
    return out


def test_cfg3_batched_free_running_400_frames():
    B, T_in, T_sub, steps, seed = 64, 120, 40, 400, 4321
    w = make_decoder_weights(SMA, seed=seed, gate_bias=0.0)
    inp = make_inputs(B, T_in, T_sub, 1, seed=seed, ragged=True)
    plan = make_dropout_plan(B, steps, steps, T_in, T_sub, False, seed=seed + 1)
    orc = DecoderOracle(w, SMA)
    with torch.no_grad():
        omel, ogate0, oal, oalb = _oracle_free_running_batched(orc, inp, plan, steps)   # gate logits with zero bias
        # the batched helper equals the per-utterance definition (two utterances, 40 frames)
        for b in (1, B - 1):
            Lm, Lb = int(inp["memory_lengths"][b]), int(inp["bert_lengths"][b])
            w20 = dict(w); w20["gate_layer.linear_layer.bias"] = torch.full_like(w["gate_layer.linear_layer.bias"], -20.0)
            m1, g1, a1, ab1, _ = DecoderOracle(w20, SMA).inference(inp["memory"][b:b + 1, :Lm], inp["embeddings"][b:b + 1, :Lb],
                                                                   plan, max_decoder_steps=40, plan_batch_index=b)
            assert maxabs(m1[0].t(), omel[b, :40]) <= 2e-5 and maxabs(a1[0], oal[b, :40, :Lm]) <= 2e-6
            assert maxabs(g1.reshape(-1) + 20.0, ogate0[b, :40]) <= 2e-5

    # threshold = median over utterances of the largest logit: about half of the utterances stop somewhere, the others run
    # into max_decoder_steps.  25,600 logits are dense around any threshold, so an utterance is "safe" when every logit the
    # stop test looks at is at least MARGIN away from it -- safe utterances must stop on exactly the oracle's frame; the
    # others may flip on a 1e-4 rounding difference and are only counted.
    MARGIN = 2e-3
    c = float(ogate0.max(1).values.median())
    want_n = _stops(ogate0, c, steps)
    safe = [bool(((ogate0[b, :n] - c).abs() >= MARGIN).all()) for b, n in enumerate(want_n)]
    assert len(set(want_n)) >= 8 and steps in want_n and sum(safe) >= B // 2, (want_n, sum(safe))
    bias = LOGIT_THR - c
    w_run = dict(w); w_run["gate_layer.linear_layer.bias"] = torch.full_like(w["gate_layer.linear_layer.bias"], bias)
    dec = make_decoder(w_run, SMA).eval()
    dec.decoder_path = "tensor"
    dec.dropout_replay = replay_of(plan)
    with torch.no_grad():
        mel, gate, al, alb, nf, reached = dec.inference_batched(inp["memory"].cuda(), inp["embeddings"].cuda(),
                                                                inp["memory_lengths"].cuda(), inp["bert_lengths"].cuda(),
                                                                max_decoder_steps=steps)
    assert dec._engine(torch.device("cuda", 0)).last_path() in ("tensor", "tensor_graph")
    got_n = [int(x) for x in nf]
    same = [g == w_ for g, w_ in zip(got_n, want_n)]
    assert all(sm for sm, sf in zip(same, safe) if sf), [(b, got_n[b], want_n[b]) for b in range(B) if safe[b] and not same[b]]
    worst = {"mel": 0.0, "gate": 0.0, "align": 0.0, "align_bert": 0.0}
    for b in range(B):
        n = min(got_n[b], want_n[b])
        if same[b]:
            assert bool(reached[b]) == (not bool((ogate0[b] > c).any()))
        Lm, Lb = int(inp["memory_lengths"][b]), int(inp["bert_lengths"][b])
        r = _cmp((mel[b, :, :n].t(), gate[b, :n, 0], al[b, :n, :Lm], alb[b, :n, :Lb]),
                 (omel[b, :n], ogate0[b, :n] + bias, oal[b, :n, :Lm], oalb[b, :n, :Lb]), 1e-3, 2e-4, f"cfg3 utt {b}")
        worst = {k: max(worst[k], r[k]) for k in worst}
        if got_n[b] < mel.shape[2]:
            assert float(mel[b, :, got_n[b]:].abs().max()) == 0.0           # frames after the stop are zeroed
    n_all = min(min(got_n), min(want_n))
    div = _first_divergence(mel.transpose(1, 2).cpu()[:, :n_all], omel[:, :n_all], 1e-3)
    print(f"cfg3 B=64 x {steps} frames, tensor path: worst {worst}; {sum(same)}/{B} utterances stop on the oracle's frame "
          f"({sum(safe)} of them with every logit >= {MARGIN} from the threshold: all match); distinct stop frames "
          f"{len(set(want_n))}; first frame with |d mel| > 1e-3 over the common prefix: {div}")
    assert sum(same) >= int(0.9 * B)      # stated fraction: >= 90 %


# ---------------------------------------------------------------------------------------------------------------
# cfg 4: teacher-forced GTA shape, B=16 per GPU, 160 phones / 53 sub-words, 800 frames
# ---------------------------------------------------------------------------------------------------------------
def test_cfg4_teacher_forced_800_frames():
    B, T_in, T_sub, T, seed = 16, 160, 53, 800, 808
    w = make_decoder_weights(SMA, seed=seed)
    inp = make_inputs(B, T_in, T_sub, T, seed=seed, ragged=True)
    plan = make_dropout_plan(B, T + 1, T, T_in, T_sub, False, seed=seed + 1)
    orc = DecoderOracle(w, SMA)
    with torch.no_grad():
        want = orc.forward(inp["memory"], inp["embeddings"], inp["mels"], inp["memory_lengths"], inp["bert_lengths"], plan)
    dec = make_decoder(w, SMA).eval()
    dec.decoder_path = "tensor"
    dec.dropout_replay = replay_of(plan)
    args = (inp["memory"].cuda(), inp["embeddings"].cuda(), inp["mels"].cuda(), inp["memory_lengths"].cuda(),
            inp["bert_lengths"].cuda())
    with torch.no_grad():
        got = dec(*args)
    worst = _cmp(got, want, 1e-3, 2e-4, "cfg4 batched semantics")
    head, tail = maxabs(got[0][:, :, :100].cpu(), want[0][:, :, :100]), maxabs(got[0][:, :, 700:].cpu(), want[0][:, :, 700:])
    print(f"cfg4 B=16 x 800 frames, tensor path: worst {worst}; mel error frames 0-99 {head:.2e}, 700-799 {tail:.2e}")
    assert tail <= 4 * head + 1e-4
    # independent-utterance mode (what gta.py runs): two rows against their own batch-1 oracle runs over all 800 frames
    from tests.test_gpu_parity import _sub_plan
    with torch.no_grad():
        mel, gate, al, alb = dec(*args, independent=True)
        for b in (0, 5):
            Lm, Lb = int(inp["memory_lengths"][b]), int(inp["bert_lengths"][b])
            w1 = orc.forward(inp["memory"][b:b + 1, :Lm], inp["embeddings"][b:b + 1, :Lb], inp["mels"][b:b + 1],
                             torch.tensor([Lm]), torch.tensor([Lb]), _sub_plan(plan, b, T))
            _cmp((mel[b:b + 1], gate[b:b + 1], al[b:b + 1, :, :Lm], alb[b:b + 1, :, :Lb]), w1, 1e-3, 2e-4, f"cfg4 independent utt {b}")


# ---------------------------------------------------------------------------------------------------------------
# cfg 5: backward through long horizons; relative error per tensor against T
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,T", [(16, 50), (16, 200), (64, 200)])
def test_cfg5_backward_long_horizon(B, T):
    from tests.test_gpu_backward import _loss, _oracle_grads
    T_in, T_sub, seed = 160, 53, 900 + B + T
    w = make_decoder_weights(SMA, seed=seed)
    inp = make_inputs(B, T_in, T_sub, T, seed=seed, ragged=True)
    plan = make_dropout_plan(B, T + 1, T, T_in, T_sub, True, seed=seed + 1)
    want, want_dmem, want_demb, want_outs = _oracle_grads(w, inp, plan, True)
    dec = make_decoder(w, SMA).train()
    dec.decoder_path = "tensor"
    dec.dropout_replay = replay_of(plan)
    mem = inp["memory"].cuda().requires_grad_(True)
    emb = inp["embeddings"].cuda().requires_grad_(True)
    outs = dec(mem, emb, inp["mels"].cuda(), inp["memory_lengths"].cuda(), inp["bert_lengths"].cuda())
    for o, wo, tol in zip(outs, want_outs, (1e-3, 1e-3, 2e-4, 2e-4)):
        assert maxabs(o.detach().cpu(), wo.detach()) <= tol
    _loss(outs, 5).backward()
    torch.cuda.synchronize()
    sd = dict(dec.named_parameters())
    worst = {}
    for name, gw in want.items():
        if gw is None:
            assert sd[name].grad is None, name
            continue
        worst[name] = float((sd[name].grad.cpu() - gw).abs().max() / gw.abs().max())
    worst["memory"] = float((mem.grad.cpu() - want_dmem).abs().max() / want_dmem.abs().max())
    worst["embeddings"] = float((emb.grad.cpu() - want_demb).abs().max() / want_demb.abs().max())
    bound = grad_bound(T)
    top = sorted(worst.items(), key=lambda kv: -kv[1])[:5]
    print(f"cfg5 backward B={B} T={T}: bound {bound:.1e}; worst tensors " + ", ".join(f"{k} {v:.1e}" for k, v in top))
    bad = {k: v for k, v in worst.items() if not v <= bound}
    assert not bad, f"relative gradient error above {bound:.1e} at T={T}: {bad}"
