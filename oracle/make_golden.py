"""Generate golden vectors from the UNMODIFIED reference (build container only).

    python -m oracle.make_golden            # rewrites tests/golden/*.npz

The reference ships no golden vectors for the decoder, so the
pins are outputs of the reference's own ``model.Decoder`` (/root/reference/model.py:128-492)
run here on CPU fp32 with

  * weights from ``oracle.synth.make_decoder_weights`` loaded via ``load_state_dict(strict=True)``,
  * inputs from ``oracle.synth.make_inputs``,
  * dropout masks / SMA noise from ``oracle.synth.make_dropout_plan`` replayed by patching
    ``torch.nn.functional.dropout`` and ``Tensor.normal_`` for the duration of the call
    (call order documented in SURVEY.md 8c and ``DropoutPlan``).

Only the *outputs* (+ the recipe to regenerate weights/inputs/masks and a weights
checksum) are stored; ``/root/reference`` does not travel to the GPU box.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.synth import (LSA, SMA, DecoderDims, DropoutPlan, make_decoder_weights,  # noqa: E402
                          make_dropout_plan, make_inputs, weights_checksum, seeded_loss, grad_digest)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> recipe.  Kept small so the whole CPU suite stays fast; "cfg1" is
# BASELINE.json configs[0] (B=1, 80 phones, 400 frames, teacher-forced, CPU).
CASES = {
    "tf_sma_eval_ragged": dict(mode="tf", attention=SMA, B=3, T_in=23, T_sub=9, T=12, ragged=True, training=False),
    "tf_sma_train_ragged": dict(mode="tf", attention=SMA, B=2, T_in=17, T_sub=6, T=10, ragged=True, training=True),
    "tf_lsa_eval_ragged": dict(mode="tf", attention=LSA, B=3, T_in=40, T_sub=13, T=9, ragged=True, training=False),
    "tf_lsa_train": dict(mode="tf", attention=LSA, B=2, T_in=33, T_sub=11, T=6, ragged=True, training=True),
    "tf_sma_cfg1": dict(mode="tf", attention=SMA, B=1, T_in=80, T_sub=26, T=400, ragged=False, training=False),
    "fr_sma_stop": dict(mode="fr", attention=SMA, B=1, T_in=30, T_sub=10, max_steps=60, gate_bias="auto", seed=7),
    "fr_sma_maxsteps": dict(mode="fr", attention=SMA, B=1, T_in=21, T_sub=7, max_steps=25, gate_bias=-20.0),
    "fr_sma_first_frame": dict(mode="fr", attention=SMA, B=1, T_in=12, T_sub=4, max_steps=10, gate_bias=None),
    "fr_lsa_stop": dict(mode="fr", attention=LSA, B=1, T_in=30, T_sub=10, max_steps=60, gate_bias="auto", seed=6),
}


@contextlib.contextmanager
def replay(plan: DropoutPlan, mode: str, training: bool, T: int, p_att=0.1, p_dec=0.1):
    """Patch F.dropout / Tensor.normal_ so the reference consumes ``plan`` in its own call order."""
    import torch.nn.functional as F

    queue = []   # (keep uint8 tensor, p)
    noise_q = []
    if mode == "tf":
        for s in range(len(plan.prenet_keep)):           # model.py:412 then :413
            queue.append(plan.prenet_keep[s][0])
            queue.append(plan.prenet_keep[s][1])
        if training:
            for t in range(T):                           # model.py:341-346, 372-373
                queue.extend(plan.lstm_keep[t, i] for i in range(6))
                noise_q.extend(n[t] for n in plan.sma_noise)   # attention.py:346-348 via model.py:355-356
    else:
        for t in range(T):                               # model.py:449-450 / 470-471
            for s in range(len(plan.prenet_keep)):
                queue.append(plan.prenet_keep[s][0][t])
                queue.append(plan.prenet_keep[s][1][t])
    queue.reverse()
    noise_q.reverse()
    orig_dropout, orig_normal = F.dropout, torch.Tensor.normal_

    def dropout(x, p=0.5, training=True, inplace=False):
        if not training:
            return x
        keep = queue.pop()
        assert keep.shape == x.shape, (keep.shape, x.shape)
        return x * (keep.to(x.dtype) * torch.tensor(1.0 / (1.0 - p), dtype=x.dtype))

    def normal_(self, *a, **k):
        n = noise_q.pop()
        assert n.shape == self.shape
        return self.copy_(n)

    F.dropout = dropout
    torch.Tensor.normal_ = normal_
    try:
        yield queue
    finally:
        F.dropout = orig_dropout
        torch.Tensor.normal_ = orig_normal


def _auto_gate_bias(case: dict, seed: int) -> float:
    """Random-init gate logits are nearly flat, so a fixed bias stops at frame 0 or never.
    Run once with bias -20 (never stops), then place logit(gate_threshold=0.001) halfway
    between the running maximum and the previous maximum so the stop lands mid-utterance."""
    probe = run_reference(dict(case, gate_bias=-20.0), seed)
    g = probe["gate"].reshape(-1).astype(np.float64) + 20.0
    k = int(np.argmax(g))
    prev = float(g[:k].max()) if k > 0 else float(g[k]) - 1.0
    cut = 0.5 * (float(g[k]) + prev)
    return float(np.float32(np.log(0.001 / 0.999) - cut))


def run_reference(case: dict, seed: int = None):
    """Run the reference decoder for one recipe; returns dict of numpy outputs."""
    seed = case.get("seed", 1234) if seed is None else seed
    if case.get("gate_bias") == "auto":
        case = dict(case, gate_bias=_auto_gate_bias(case, seed))
    attention = case["attention"]
    dec, hp = ref_shim.build_reference_decoder(attention)
    w = make_decoder_weights(attention, seed=seed, gate_bias=case.get("gate_bias"))
    if attention == LSA:
        # the LSA branch builds no decoder-side bert attention; ref_shim attached one.
        pass
    dec.load_state_dict(w, strict=True)
    B, T_in, T_sub = case["B"], case["T_in"], case["T_sub"]
    out = {}
    with torch.no_grad():
        if case["mode"] == "tf":
            T = case["T"]
            inp = make_inputs(B, T_in, T_sub, T, seed=seed, ragged=case["ragged"])
            plan = make_dropout_plan(B, T + 1, T, T_in, T_sub, case["training"], seed=seed + 1)
            dec.train(case["training"])
            with replay(plan, "tf", case["training"], T) as q:
                mel, gate, al, alb = dec(inp["memory"], inp["embeddings"], inp["mels"],
                                         inp["memory_lengths"], inp["bert_lengths"])
                assert not q, "dropout replay queue not drained"
            out.update(mel=mel, gate=gate, align=al, align_bert=alb)
        else:
            max_steps = case["max_steps"]
            inp = make_inputs(B, T_in, T_sub, 1, seed=seed)
            plan = make_dropout_plan(B, max_steps, max_steps, T_in, T_sub, False, seed=seed + 1)
            dec.eval()
            dec.max_decoder_steps = max_steps
            with replay(plan, "fr", False, max_steps), contextlib.redirect_stdout(io.StringIO()):
                mel, gate, al, alb, flag = dec.inference(inp["memory"], inp["embeddings"])
            out.update(mel=mel, gate=gate, align=al, align_bert=alb,
                       flag=torch.tensor(int(flag)), n_frames=torch.tensor(mel.shape[2]))
    res = {k: v.detach().cpu().numpy() for k, v in out.items()}
    res["weights_checksum"] = np.array(weights_checksum(w))
    res["recipe"] = np.array(json.dumps(dict(case, seed=seed)))
    return res


# ---------------------------------------------------------------------------------------------------
# Gradient fixtures: the reference's own autograd (loss.backward() over model.py:392-428) on a seeded loss.
# Full gradients are ~200 MB, so a fixture keeps per-tensor digests: max|g|, sum(g), a seeded random projection
# <g, r> and the first 32 elements.  tests/ compare the oracle's autograd (and, on the GPU, the CUDA backward)
# against the same digests.
# ---------------------------------------------------------------------------------------------------
GRAD_CASES = {
    "grad_sma_train_B16": dict(mode="tf", attention=SMA, B=16, T_in=24, T_sub=8, T=5, ragged=True, training=True, seed=516),
    "grad_sma_eval_B3": dict(mode="tf", attention=SMA, B=3, T_in=11, T_sub=4, T=4, ragged=True, training=False, seed=33),
    "grad_lsa_train_B4": dict(mode="tf", attention=LSA, B=4, T_in=19, T_sub=7, T=5, ragged=True, training=True, seed=44),
}


def run_reference_grads(case: dict, loss_seed: int = 5):
    """Reference Decoder.forward + autograd for one teacher-forced recipe -> digests of every gradient."""
    seed = case["seed"]
    attention = case["attention"]
    dec, hp = ref_shim.build_reference_decoder(attention)
    w = make_decoder_weights(attention, seed=seed)
    dec.load_state_dict(w, strict=True)
    B, T_in, T_sub, T = case["B"], case["T_in"], case["T_sub"], case["T"]
    inp = make_inputs(B, T_in, T_sub, T, seed=seed, ragged=case["ragged"])
    plan = make_dropout_plan(B, T + 1, T, T_in, T_sub, case["training"], seed=seed + 1)
    dec.train(case["training"])
    mem = inp["memory"].clone().requires_grad_(True)
    emb = inp["embeddings"].clone().requires_grad_(True)
    with replay(plan, "tf", case["training"], T) as q:
        outs = dec(mem, emb, inp["mels"], inp["memory_lengths"], inp["bert_lengths"])
        assert not q, "dropout replay queue not drained"
    seeded_loss(outs, loss_seed).backward()
    res = {}
    for n, p_ in dec.named_parameters():
        if p_.grad is None:
            res[n + "/none"] = np.array(1)
        else:
            res.update(grad_digest(n, p_.grad))
    res.update(grad_digest("memory", mem.grad))
    res.update(grad_digest("embeddings", emb.grad))
    res["recipe"] = np.array(json.dumps(dict(case, loss_seed=loss_seed)))
    res["weights_checksum"] = np.array(weights_checksum(w))
    return res


def run_reference_postnet(seed: int = 21, B: int = 3, T: int = 37):
    """Unmodified reference Postnet (model.py:27-70) in eval mode on seeded weights / BatchNorm statistics."""
    from oracle.postnet_oracle import make_postnet_weights
    ref_model, _, ref_hparams = ref_shim.import_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        hp = ref_hparams.create_hparams()
    net = ref_model.Postnet(hp).eval()
    w = make_postnet_weights(seed)
    sd = net.state_dict()
    for k in sd:
        if k in w:
            sd[k] = w[k]
    net.load_state_dict(sd, strict=True)
    x = torch.randn(B, hp.n_mel_channels, T, generator=torch.Generator().manual_seed(seed + 1))
    with torch.no_grad():
        y = net(x)
    return {"x": x.numpy(), "y": y.numpy(), "seed": np.array(seed)}


def run_reference_memprep(seed: int = 31, B: int = 2, T: int = 9):
    """The reference's own modules for the decoder-input step: layers.LinearNorm wired as model.py:548-549 (linear_converter
    over cat(encoder_outputs, cls)) and as model.py:258-261 (the attention layer's memory_layer), on seeded weights."""
    from oracle.memprep_oracle import make_memprep_inputs, make_memprep_weights
    ref_shim.import_reference()
    import layers as ref_layers
    w = make_memprep_weights(seed)
    conv = ref_layers.LinearNorm(512 + 768, 512)
    meml = ref_layers.LinearNorm(512, 128, bias=False, w_init_gain="tanh")
    conv.load_state_dict({"linear_layer.weight": w["linear_converter.linear_layer.weight"],
                          "linear_layer.bias": w["linear_converter.linear_layer.bias"]}, strict=True)
    meml.load_state_dict({"linear_layer.weight": w["memory_layer.linear_layer.weight"]}, strict=True)
    enc, cls = make_memprep_inputs(B, T, seed + 1)
    with torch.no_grad():
        memory = conv(torch.cat([enc, cls], 2))
        pm = meml(memory)
    return {"memory": memory.numpy(), "processed_memory": pm.numpy(), "seed": np.array(seed), "B": np.array(B), "T": np.array(T)}


def run_reference_loss(alignloss: str, seed: int = 61, B: int = 3, T: int = 21, T_in: int = 13):
    """Unmodified reference Tacotron2Loss (loss_function.py:7-66) + torch.autograd on a seeded case: loss terms and, per
    output, max|g|, sum(g) and the first 16 elements of d total / d output."""
    from oracle.loss_oracle import make_loss_case
    ref_shim.import_reference()
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    try:
        import loss_function as ref_loss
    finally:
        sys.path.remove(ref_shim.REFERENCE_ROOT)
    c = make_loss_case(B, T, T_in, seed)
    outs = [c[k].clone().requires_grad_(True) for k in ("mel", "mel_postnet", "gate", "align", "align_bert")]
    targets = (c["mel_target"].clone(), c["gate_target"].clone(), c["align_target"].clone())
    total, mel_loss, gate_loss, al, alb = ref_loss.Tacotron2Loss(alignloss)(outs, targets, None, 0)
    total.backward()
    res = {"total": total.detach().numpy(), "mel_loss": mel_loss.detach().numpy(), "gate_loss": gate_loss.detach().numpy(),
           "align_loss": np.array(float("nan") if al is None else float(al)), "align_bert_loss": np.array(float("nan") if alb is None else float(alb)),
           "alignloss": np.array(alignloss), "seed": np.array(seed), "B": np.array(B), "T": np.array(T), "T_in": np.array(T_in)}
    for k, o in zip(("mel", "mel_postnet", "gate", "align", "align_bert"), outs):
        if o.grad is None:
            continue
        g = o.grad
        res[f"g_{k}/max"] = g.abs().max().numpy()
        res[f"g_{k}/sum"] = g.double().sum().numpy()
        res[f"g_{k}/head"] = g.reshape(-1)[:16].numpy()
    return res


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for al in ("", "L2"):
        res = run_reference_loss(al)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"loss_{al or 'default'}.npz"), **res)
        print(f"loss_{al or 'default'}: total {float(res['total']):.6f}")
    res = run_reference_memprep()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "memprep.npz"), **res)
    print(f"memprep: memory{res['memory'].shape}")
    res = run_reference_postnet()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "postnet_eval.npz"), **res)
    print(f"postnet_eval: y{res['y'].shape}")
    for name, case in GRAD_CASES.items():
        res = run_reference_grads(case)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **res)
        print(f"{name}: {sum(1 for k in res if k.endswith('/max'))} gradient digests -> {os.path.getsize(path)//1024} KiB")
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name, case in CASES.items():
        res = run_reference(case)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **res)
        extra = f" n_frames={int(res['n_frames'])} flag={int(res['flag'])}" if "flag" in res else ""
        print(f"{name}: mel{res['mel'].shape} wsum={res['weights_checksum']}{extra} -> {os.path.getsize(path)//1024} KiB")


if __name__ == "__main__":
    main()
