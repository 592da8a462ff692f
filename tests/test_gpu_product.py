"""-m gpu: product-behaviour regressions."""
import pytest
import torch

from oracle.decoder_oracle import DecoderOracle
from oracle.synth import SMA, make_decoder_weights, make_dropout_plan, make_inputs
from tests.gpu_util import make_decoder, replay_of
from tests.helpers import maxabs

pytestmark = pytest.mark.gpu


def _args(inp):
    return (inp["memory"].cuda(), inp["embeddings"].cuda(), inp["mels"].cuda(), inp["memory_lengths"].cuda(),
            inp["bert_lengths"].cuda())


def test_training_steps_do_not_leak_saved_activations():
    """Five forward/backward steps: allocated memory after each step must stay flat (the saved-activation buffer of a
    step, ~27 MB here, is released by backward; round 1 leaked it through a ctx reference cycle)."""
    B, T, T_in, T_sub, seed = 16, 12, 30, 10, 3
    w = make_decoder_weights(SMA, seed=seed)
    inp = make_inputs(B, T_in, T_sub, T, seed=seed, ragged=True)
    dec = make_decoder(w, SMA, exact=False).train()
    dec.rng_seed = 5
    args = _args(inp)
    used = []
    import gc
    gc.disable()
    try:
        for _ in range(5):
            outs = dec(*args)
            outs[0].square().mean().backward()
            del outs
            dec.zero_grad(set_to_none=True)
            torch.cuda.synchronize()
            used.append(torch.cuda.memory_allocated())
    finally:
        gc.enable()
    assert max(used[1:]) - min(used[1:]) < (1 << 20), used


def test_data_inplace_weight_update_is_picked_up():
    """`p.data.add_()` does not bump `_version`: eval-mode callers must call invalidate_weights(); training mode re-packs on
    every call.  Latency path (packed streams) and tensor path (fp16 tiles) must both follow the new weights."""
    T_in, T_sub, T, seed = 20, 7, 5, 8
    w = make_decoder_weights(SMA, seed=seed)
    for B, path in ((1, "latency"), (4, "tensor")):
        inp = make_inputs(B, T_in, T_sub, T, seed=seed)
        plan = make_dropout_plan(B, T + 1, T, T_in, T_sub, False, seed=seed + 1)
        dec = make_decoder(w, SMA).eval()
        dec.decoder_path = path
        dec.dropout_replay = replay_of(plan)
        with torch.no_grad():
            a = dec(*_args(inp))[0].clone()
            dec.decoder_rnn.weight_hh.data.mul_(0.5)
            dec.attention_rnn.weight_ih.data.mul_(0.5)
            stale = dec(*_args(inp))[0].clone()
            dec.invalidate_weights()
            fresh = dec(*_args(inp))[0].clone()
        w2 = {k: v.clone() for k, v in w.items()}
        w2["decoder_rnn.weight_hh"] *= 0.5
        w2["attention_rnn.weight_ih"] *= 0.5
        want = DecoderOracle(w2, SMA).forward(inp["memory"], inp["embeddings"], inp["mels"], inp["memory_lengths"],
                                              inp["bert_lengths"], plan)[0]
        assert maxabs(fresh.cpu(), want) <= 1e-3, path
        assert maxabs(stale, a) == 0.0 and maxabs(fresh, a) > 1e-3, path       # documents the blind spot + the fix
        # training mode: no explicit invalidation needed
        dec.train()
        dec.decoder_rnn.weight_hh.data.mul_(2.0)
        dec.attention_rnn.weight_ih.data.mul_(2.0)
        with torch.no_grad():
            back = dec(*_args(inp))[0]
        # (train mode adds LSTM-state dropout, so compare against a train-mode call with the original weights)
        dec2 = make_decoder(w, SMA).train()
        dec2.decoder_path = path
        dec2.dropout_replay = replay_of(plan)
        dec2.rng_seed = dec.rng_seed = 17
        with torch.no_grad():
            back = dec(*_args(inp))[0]
            ref = dec2(*_args(inp))[0]
        assert maxabs(back, ref) <= 1e-5, path


def test_auto_path_is_fast_by_default_and_fp32_is_honoured():
    """decoder_path='auto' sends 2 <= B <= 128 to the tensor path by default; batched_precision='fp32'
    keeps it on the fp32-exact generic kernel and refuses to train (there is no fp32 backward) instead of silently
    switching precision."""
    B, T_in, T_sub, T, seed = 4, 18, 6, 4, 12
    w = make_decoder_weights(SMA, seed=seed)
    inp = make_inputs(B, T_in, T_sub, T, seed=seed)
    dec = make_decoder(w, SMA, exact=False).eval()
    dec.rng_seed = 1
    eng = dec._engine(torch.device("cuda", 0))
    with torch.no_grad():
        dec(*_args(inp))
    assert eng.last_path() in ("tensor", "tensor_graph")
    dec.batched_precision = "fp32"
    with torch.no_grad():
        dec(*_args(inp))
    assert eng.last_path() == "generic"
    dec.train()
    with pytest.raises(NotImplementedError):
        dec(*_args(inp))
    dec.batched_precision = "fp16"
    outs = dec(*_args(inp))
    outs[0].sum().backward()
    assert dec.decoder_rnn.weight_hh.grad is not None


def test_check_reports_and_clears_nothing_on_a_healthy_run():
    B, T_in, T_sub, T, seed = 3, 18, 6, 4, 13
    w = make_decoder_weights(SMA, seed=seed)
    inp = make_inputs(B, T_in, T_sub, T, seed=seed)
    dec = make_decoder(w, SMA).eval()
    with torch.no_grad():
        dec(*_args(inp))
    dec.check()
    dec.check(sync=False)
