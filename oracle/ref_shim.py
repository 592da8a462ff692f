"""Import the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``tacotron2_subword_b200``; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline leg may use it.

The reference cannot be imported as-is in this image:
  * ``layers.py:2`` / ``utils.py:4-7`` / ``stft.py`` import ``librosa`` and
    ``matplotlib`` which are not installed -> empty ``sys.modules`` stubs (only
    the STFT / plotting code uses them, never the decoder);
  * ``utils.py:12`` hard-codes ``torch.cuda.LongTensor`` -> replaced with the
    same expression on ``lengths.device``;
  * every attention except SMA never constructs ``attention_layer_bert``
    (``model.py:158-191``) although ``decode`` uses it (``model.py:356``) ->
    ``build_reference_decoder`` attaches an unmodified reference LSA module.

``/root/reference`` does not exist on the GPU box: callers must go through
``reference_available()``.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TACO2_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model.py"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


_cached = None


def import_reference():
    """Returns (model, attention, hparams) reference modules."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _stub("librosa", filters=_stub("librosa.filters", mel=None),
          util=_stub("librosa.util", normalize=None, pad_center=None, tiny=None))
    _stub("matplotlib", use=lambda *a, **k: None, pylab=_stub("matplotlib.pylab"))
    import importlib

    # the reference is a flat script directory; it also has its own ``utils``,
    # ``layers``, ``hparams`` ... module names, so import under a scoped path.
    saved = {k: sys.modules.pop(k) for k in
             ("model", "attention", "layers", "utils", "hparams", "stft", "audio_processing")
             if k in sys.modules}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        ref_utils = importlib.import_module("utils")
        ref_model = importlib.import_module("model")
        ref_attention = importlib.import_module("attention")
        ref_hparams = importlib.import_module("hparams")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    import torch

    def get_mask_from_lengths(lengths):  # utils.py:10-14 on lengths.device
        max_len = torch.max(lengths).item()
        ids = torch.arange(0, max_len, device=lengths.device, dtype=torch.long)
        return (ids < lengths.unsqueeze(1)).bool()

    ref_utils.get_mask_from_lengths = get_mask_from_lengths
    ref_model.get_mask_from_lengths = get_mask_from_lengths
    # keep the reference modules reachable under private names, restore ours
    for k in ("model", "attention", "layers", "utils", "hparams", "stft", "audio_processing"):
        if k in sys.modules:
            sys.modules["_taco2ref_" + k] = sys.modules[k]
            # leave them registered too: reference modules import each other lazily
    for k, v in saved.items():
        sys.modules[k] = v
    _cached = (ref_model, ref_attention, ref_hparams)
    return _cached


def build_reference_decoder(attention: str = "StepwiseMonotonicAttention", **hp_overrides):
    """``model.Decoder(hparams)`` from the reference, plus the LSA shim."""
    ref_model, ref_attention, ref_hparams = import_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        hp = ref_hparams.create_hparams()
        hp.attention = attention
        for k, v in hp_overrides.items():
            setattr(hp, k, v)
        dec = ref_model.Decoder(hp)
        if attention != "StepwiseMonotonicAttention":
            dec.attention_layer_bert = ref_attention.LocationSensitiveAttention(
                hp.attention_rnn_dim, hp.encoder_embedding_dim, hp.attention_dim,
                hp.attention_location_n_filters, hp.attention_location_kernel_size)
    return dec, hp
