"""CPU: host-side logic of the drop-in boundary (no GPU compute calls)."""
import ctypes
import os
import re

import pytest
import torch

from tacotron2_subword_b200 import BERT_Tacotron2, Decoder, Tacotron2, _cabi, create_hparams
from tacotron2_subword_b200.utils import get_mask_from_lengths

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """The C-ABI shared library loads and exports exactly what include/taco2dec.h declares."""
    hdr = open(os.path.join(ROOT, "include", "taco2dec.h")).read()
    declared = set(re.findall(r"\b(taco2dec_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "header declares no functions?"
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(_cabi.EXPORTED_SYMBOLS), declared ^ set(_cabi.EXPORTED_SYMBOLS)
    lib.taco2dec_abi_version.restype = ctypes.c_int
    assert lib.taco2dec_abi_version() == _cabi.ABI_VERSION


def test_ctypes_structs_match_header_field_order():
    hdr = open(os.path.join(ROOT, "include", "taco2dec.h")).read()

    def fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                m = re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*(\[[^\]]*\])*\s*$", part.strip())
                names.append(m.group(1))
        return names

    assert fields("taco2dec_config") == [f[0] for f in _cabi.Config._fields_]
    assert fields("taco2dec_stream_weights") == [f[0] for f in _cabi.StreamWeights._fields_]
    assert fields("taco2dec_weights") == [f[0] for f in _cabi.Weights._fields_]
    assert fields("taco2dec_rng") == [f[0] for f in _cabi.Rng._fields_]
    assert fields("taco2dec_tf_args") == [f[0] for f in _cabi.TFArgs._fields_]
    assert fields("taco2dec_infer_args") == [f[0] for f in _cabi.InferArgs._fields_]
    assert fields("taco2dec_saved_layout") == [f[0] for f in _cabi.SavedLayout._fields_]
    assert fields("taco2dec_grad_layout") == [f[0] for f in _cabi.GradLayout._fields_]
    assert fields("taco2dec_bwd_args") == [f[0] for f in _cabi.BwdArgs._fields_]
    assert fields("taco2dec_postnet_layer") == [f[0] for f in _cabi.PostnetLayer._fields_]
    assert fields("taco2dec_postnet_weights") == [f[0] for f in _cabi.PostnetWeights._fields_]


def test_create_refuses_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    lib = _cabi.load_library()
    cfg = _cabi.Config(80, 512, 1024, 1024, 256, 128, 32, 31, 0, 2, 0.1, 0.1)
    h = ctypes.c_void_p()
    rc = lib.taco2dec_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc != 0 and lib.taco2dec_last_error()          # no device -> loud error, never a CPU fallback


def test_create_rejects_bad_config():
    lib = _cabi.load_library()
    cfg = _cabi.Config(80, 512, 1024, 1024, 256, 128, 32, 31, 7, 2, 0.1, 0.1)   # attention kind 7
    h = ctypes.c_void_p()
    assert lib.taco2dec_create(ctypes.byref(cfg), 0, ctypes.byref(h)) == -1
    assert b"attention" in lib.taco2dec_last_error()


def test_decoder_has_no_cpu_path():
    dec = Decoder(create_hparams())
    with pytest.raises(_cabi.Taco2DecError):
        dec(torch.zeros(1, 3, 512), torch.zeros(1, 2, 512), torch.zeros(1, 80, 4), torch.tensor([3]), torch.tensor([2]))
    with pytest.raises(ValueError):
        dec.inference(torch.zeros(2, 3, 512), torch.zeros(2, 2, 512))          # batch-1 only, model.py:461


def test_get_mask_from_lengths_exact():
    lens = torch.tensor([3, 5, 1])
    m = get_mask_from_lengths(lens)
    assert m.dtype == torch.bool and m.shape == (3, 5)
    assert m.tolist() == [[1, 1, 1, 0, 0], [1, 1, 1, 1, 1], [1, 0, 0, 0, 0]]


def test_hparams_string_override_keeps_reference_semantics():
    hp = create_hparams("{attention:LocationSensitiveAttention-batch_size:16}}")
    assert hp.attention == "LocationSensitiveAttention"
    assert hp.batch_size == "16"            # values stay strings (hparams.py:108-114)
    assert hp.max_decoder_steps == 1000 and hp.gate_threshold == 0.001 and hp.p_attention_dropout == 0.1


def test_attention_choice_and_stream_variants():
    hp = create_hparams()
    sma = Decoder(hp).state_dict()
    assert "attention_layer_bert.v.weight" in sma and sma["decoder_rnn.weight_ih"].shape == (4096, 3072)
    assert "decoder_rnn_bert.weight_ih" in sma            # dead cell stays in the checkpoint layout
    hp.attention = "LocationSensitiveAttention"
    lsa = Decoder(hp).state_dict()
    assert lsa["attention_layer_bert.location_layer.location_conv.conv.weight"].shape == (32, 2, 31)
    assert "attention_layer.v.linear_layer.weight" in lsa
    one = Decoder(hp, n_streams=1).state_dict()
    assert one["decoder_rnn.weight_ih"].shape == (4096, 1536) and one["linear_projection.linear_layer.weight"].shape == (80, 1536)
    hp.attention = "GMMAttention"
    with pytest.raises(ValueError):
        Decoder(hp)
    assert sum(p.numel() for p in Tacotron2(create_hparams()).parameters()) > 2e7


@pytest.mark.reference
def test_state_dict_layout_identical_to_reference():
    import contextlib, io
    from oracle.ref_shim import import_reference
    ref_model, _, ref_hp = import_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        ref = ref_model.BERT_Tacotron2(ref_hp.create_hparams())
    mine = BERT_Tacotron2(create_hparams())
    a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    assert list(a) == list(b) and a == b
    mine.load_state_dict(ref.state_dict(), strict=True)
    assert dict(ref_hp.create_hparams()) == dict(create_hparams())


@pytest.mark.reference
def test_oracle_vs_live_reference():
    """Re-run the unmodified reference live (build container only) against the oracle on a fresh case."""
    from oracle.decoder_oracle import DecoderOracle
    from oracle.make_golden import run_reference
    from oracle.synth import SMA
    from tests.helpers import materialise, maxabs
    case = dict(mode="tf", attention=SMA, B=2, T_in=14, T_sub=5, T=5, ragged=True, training=True, seed=77)
    ref = run_reference(case)
    w, inp, plan = materialise(case)
    out = DecoderOracle(w, SMA).forward(inp["memory"], inp["embeddings"], inp["mels"], inp["memory_lengths"],
                                        inp["bert_lengths"], plan, training=True)
    for k, v in zip(("mel", "gate", "align", "align_bert"), out):
        assert maxabs(v, torch.from_numpy(ref[k])) <= 2e-6, k


def test_decoder_tf_function_does_not_pin_saved_activations():
    """_DecoderTF must not keep its own outputs reachable from ctx -- output -> grad_fn -> ctx -> output
    is a cycle that pins the multi-GB saved-activation buffer.  With the C call mocked out, dropping the outputs must free
    the buffer by reference counting alone (no gc.collect())."""
    import gc
    import weakref
    from tacotron2_subword_b200.model import _DecoderTF
    dec = Decoder(create_hparams())
    refs = []

    def fake_run_tf(memory, embeddings, dec_in, mlen, blen, save, independent=False):
        B, T = memory.shape[0], dec_in.shape[2]
        outs = (torch.zeros(B, T, 80), torch.zeros(B, T), torch.zeros(B, T, memory.shape[1]), torch.zeros(B, T, embeddings.shape[1]))
        saved = torch.zeros(1 << 16, dtype=torch.uint8)
        refs.append(weakref.ref(saved))
        return outs, dict(saved=saved, align=outs[2], align_b=outs[3], params=dec._weight_tensors())

    dec._run_tf = fake_run_tf
    gc.disable()
    try:
        for _ in range(3):
            outs = _DecoderTF.apply(dec, False, torch.zeros(2, 5, 512, requires_grad=True), torch.zeros(2, 3, 512), torch.zeros(2, 80, 4),
                                    None, None, *dec._weight_tensors())
            assert outs[0].grad_fn is not None
            del outs
        assert all(r() is None for r in refs), "saved-activation buffers are still alive after their outputs were dropped"
    finally:
        gc.enable()


def test_invalidate_weights_and_training_mode_force_a_rebind():
    """`.data` updates do not bump ``_version``; invalidate_weights() drops the cached key, and in
    training mode the key is never trusted."""
    dec = Decoder(create_hparams())

    class FakeEng:
        weights_key = None
        device = torch.device("cpu")

        def set_mode(self, *a):
            pass

    eng = FakeEng()
    dec._engines[0] = eng
    key = tuple((t.data_ptr(), t._version) for t in dec._weight_tensors())
    eng.weights_key = key
    dec.decoder_rnn.weight_hh.data.add_(1.0)
    assert tuple((t.data_ptr(), t._version) for t in dec._weight_tensors()) == key     # the blind spot itself
    dec.invalidate_weights()
    assert eng.weights_key is None
    eng.weights_key = key
    dec.train()
    with pytest.raises(Exception):      # training mode goes on to re-pack (and fails here only because there is no library handle)
        dec._bind_weights(eng)
    dec.eval()
    dec._bind_weights(eng)              # eval + unchanged key: early return, nothing touched


def test_tf_resized_splits_batches_and_slices_replayed_masks():
    """Decoder._tf_resized (batches outside one tensor-path call): B = 1 runs as a pair whose second row is a detached copy, B > rows as
    balanced sub-batches (sizes differ by at most one, none of size 1) over the same padded memory, replayed masks sliced along the
    batch, Philox seeds offset per sub-batch, module state restored afterwards.  The CUDA call is replaced by a recorder."""
    from tacotron2_subword_b200 import DropoutReplay
    dec = Decoder(create_hparams())
    calls = []

    def run(memory, embeddings, dec_in, mlen, blen, independent):
        rp = dec.dropout_replay
        calls.append(dict(B=memory.shape[0], T_in=memory.shape[1], seed=dec.rng_seed, validate=dec.validate_lengths,
                          pk=None if rp is None else tuple(rp.prenet_keep[0][0].shape), lk=None if rp is None else tuple(rp.lstm_keep.shape),
                          first=float(memory.detach()[0, 0, 0])))
        B, T = memory.shape[0], dec_in.shape[2]
        return (memory[:, :1, :80].expand(B, T, 80).transpose(1, 2) * 1.0, torch.zeros(B, T), torch.zeros(B, T, memory.shape[1]),
                torch.zeros(B, T, embeddings.shape[1]))

    for B, rows, want_sizes in ((1, 128, [2]), (11, 4, [3, 4, 4]), (130, 128, [65, 65]), (257, 128, [85, 86, 86])):
        calls.clear()
        T, T_in, T_sub = 3, 6, 2
        dec.max_backward_rows = rows
        dec.rng_seed = 99
        dec.dropout_replay = DropoutReplay(prenet_keep=[[torch.ones(T + 1, B, 256, dtype=torch.uint8) for _ in range(2)] for _ in range(2)],
                                           lstm_keep=torch.ones(T, 6, B, 1024, dtype=torch.uint8), sma_noise=[torch.zeros(T, B, T_in), torch.zeros(T, B, T_sub)])
        mem = torch.arange(B, dtype=torch.float32).view(B, 1, 1).expand(B, T_in, 512).clone().requires_grad_(True)
        outs = dec._tf_resized(run, mem, torch.zeros(B, T_sub, 512), torch.zeros(B, 80, T), torch.full((B,), T_in), torch.full((B,), T_sub), False)
        assert [c_["B"] for c_ in calls] == want_sizes
        assert all(c_["T_in"] == T_in and not c_["validate"] for c_ in calls)            # same padded memory, lengths validated once
        assert all(c_["pk"] == (T + 1, c_["B"], 256) and c_["lk"] == (T, 6, c_["B"], 1024) for c_ in calls)
        assert outs[0].shape == (B, 80, T) and outs[2].shape == (B, T, T_in)
        assert torch.equal(outs[0][:, 0, 0], torch.arange(B, dtype=torch.float32))       # rows come back in order
        if B > 1:
            assert len({c_["seed"] for c_ in calls}) == len(calls)                        # one Philox stream per sub-batch
            assert [c_["first"] for c_ in calls] == [float(sum(want_sizes[:k])) for k in range(len(want_sizes))]
        outs[0].sum().backward()
        assert mem.grad.shape == mem.shape
        assert dec.rng_seed == 99 and dec.validate_lengths and dec.dropout_replay.lstm_keep.shape[2] == B    # state restored
    with pytest.raises(ValueError):
        dec._tf_resized(run, torch.zeros(130, 6, 512), torch.zeros(130, 2, 512), torch.zeros(130, 80, 3), torch.full((130,), 5),
                        torch.full((130,), 2), False)
