"""Helpers for the -m gpu parity tests: build the product Decoder from oracle weights."""
from __future__ import annotations

import torch

from oracle.synth import LSA, SMA, DecoderDims
from tacotron2_subword_b200 import Decoder, DropoutReplay, create_hparams


def make_decoder(weights, attention=SMA, n_streams=2, device="cuda", exact=True, **hp_over):
    """``exact=True``: batched calls on the default ``decoder_path="auto"`` stay on the fp32-exact generic kernel (the
    product default sends 2 <= B <= 128 to the fp16-operand tensor path; tests of that path select it explicitly)."""
    hp = create_hparams()
    hp.attention = attention
    for k, v in hp_over.items():
        hp[k] = v
    dec = Decoder(hp, n_streams=n_streams)
    missing = dec.load_state_dict(weights, strict=True)
    dec.batched_precision = "fp32" if exact else "fp16"
    return dec.to(device)


def replay_of(plan):
    return DropoutReplay(prenet_keep=plan.prenet_keep, lstm_keep=plan.lstm_keep, sma_noise=plan.sma_noise)
