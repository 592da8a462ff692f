"""Shared test plumbing: rebuild weights / inputs / dropout plans from a golden recipe."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from oracle.synth import (LSA, SMA, DecoderDims, make_decoder_weights, make_dropout_plan,
                          make_inputs, weights_checksum)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR)
                  if f.endswith(".npz") and not f.startswith(("grad_", "postnet_", "memprep", "loss_")))


def grad_golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith("grad_"))


def load_grad_golden(name: str):
    """(recipe, {tensor name: digest dict or None for parameters the reference leaves without gradient})."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    recipe = json.loads(str(z["recipe"]))
    digests = {}
    for k in z.files:
        if "/" not in k:
            continue
        n, field = k.rsplit("/", 1)
        if field == "none":
            digests[n] = None
        else:
            digests.setdefault(n, {})[field] = z[k]
    return recipe, digests, str(z["weights_checksum"])


def check_grad_digest(name: str, g: torch.Tensor, want: dict, rtol: float) -> float:
    """Compare a gradient with a reference digest; every field is held to rtol * max|g_ref|.  Returns the worst ratio."""
    from oracle.synth import grad_digest
    got = grad_digest(name, g)
    scale = float(want["max"])
    n = g.numel()
    errs = [abs(float(got[name + "/max"]) - scale) / scale,
            abs(float(got[name + "/proj"]) - float(want["proj"])) / scale,            # projection is normalised by sqrt(n)
            abs(float(got[name + "/sum"]) - float(want["sum"])) / (scale * max(1.0, n ** 0.5)),
            float(np.abs(got[name + "/head"] - want["head"]).max()) / scale]
    worst = max(errs)
    assert worst <= rtol, f"{name}: gradient digest off by {worst:.2e} of max|g| (fields max/proj/sum/head: {errs})"
    return worst


def load_golden(name: str):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    recipe = json.loads(str(z["recipe"]))
    out = {k: torch.from_numpy(z[k]) for k in z.files if k not in ("recipe", "weights_checksum")}
    return recipe, out, str(z["weights_checksum"])


def materialise(recipe: dict):
    """(weights, inputs, plan) exactly as oracle/make_golden.py built them."""
    seed = recipe["seed"]
    w = make_decoder_weights(recipe["attention"], seed=seed, gate_bias=recipe.get("gate_bias"))
    B, T_in, T_sub = recipe["B"], recipe["T_in"], recipe["T_sub"]
    if recipe["mode"] == "tf":
        T = recipe["T"]
        inp = make_inputs(B, T_in, T_sub, T, seed=seed, ragged=recipe["ragged"])
        plan = make_dropout_plan(B, T + 1, T, T_in, T_sub, recipe["training"], seed=seed + 1)
    else:
        ms = recipe["max_steps"]
        inp = make_inputs(B, T_in, T_sub, 1, seed=seed)
        plan = make_dropout_plan(B, ms, ms, T_in, T_sub, False, seed=seed + 1)
    return w, inp, plan


def maxabs(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).abs().max()) if a.numel() else 0.0
