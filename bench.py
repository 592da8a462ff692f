#!/usr/bin/env python
"""bench.py -- mel frames/s of the Tacotron2 dual-stream decoder hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Headline workload (BASELINE.json configs[1], the configuration the metric is quoted on): free-running
``Decoder.inference``, batch 1, 150-phone / 50-sub-word synthetic memory, max_decoder_steps=1000, default hparams
(dual-stream, StepwiseMonotonicAttention), gate bias -20 so every utterance runs exactly 1000 frames.  One "step" = one
utterance.  N GPUs = N independent utterances, one per rank (utterance sharding, no data-path collective; weak scaling).

The JSON line (rank 0):
  value         frames/s, inputs resident in HBM (device-timed with CUDA events, max over ranks)
  e2e           the same metric through ``Decoder.inference`` with pinned HOST inputs and ALL five outputs copied back
  roofline      the binding roof of the latency kernel.  Its weights live in shared memory, tensor memory and L2 (DRAM pipe
                ~3 % busy), so the HBM-algorithmic figure (kept as ``hbm_algorithmic``) is not a bound: the frame time is
                a chain of cross-CTA exchanges through L2 plus one L2 pass over the non-resident weights.  Both are
                measured on this device in this run (``taco2dec_measure_machine``): floor = hops x hop latency + L2 bytes /
                L2 read rate; frac = floor / achieved.  ``traffic`` comes from the committed ncu capture and is only
                reported when that capture was taken with the library that is loaded now (sha256 stamp).
  cpu_baseline  the reference arm (below) run as a subprocess on the same host before any GPU work: same function, same
                sample, same process conditions as ``--impl reference``
  sub_records   the BASELINE configs that have a bottleneck when sharded: cfg 3 (64 utterances in total, strong scaling,
                persistent tcgen05 kernel), cfg 4 (128 utterances in total, teacher-forced GTA shape), cfg 5 (decoder
                forward + backward + NCCL gradient all-reduce, 64 utterances per GPU, weak scaling)

``--impl reference``: the CPU oracle port of the reference decoder (the reference is Python and does not travel to the
GPU box) on all host threads, full 1000-frame utterances.
"""
from __future__ import annotations

import argparse
import contextlib
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(B=1, T_in=150, T_sub=50, max_steps=1000, attention="StepwiseMonotonicAttention", gate_bias=-20.0, seed=1234)
W_ACT = 32_082_257        # weights touched per frame, dual-stream SMA
LAT_HOPS_PER_FRAME = 7    # dependent cross-CTA exchanges of the latency kernel per free-running frame:
                          # prenet -> h1 -> q partials -> context -> h2 -> projection/stop -> prenet L0 partials -> prenet
LIB = os.path.join(ROOT, "tacotron2_subword_b200", "csrc", "libtaco2dec.so")

_REAL_STDOUT = None


def _claim_stdout() -> None:
    """The contract is ONE JSON line on stdout; libraries print there too, so stdout is pointed at stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)
        sys.stdout = sys.stderr


def emit(line: dict) -> None:
    payload = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(payload.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, payload)


def algorithmic_bytes_per_frame(B, T_in, T_sub, bytes_per_weight=4):
    """W_act*s_w + B*[(T_in+T_sub)*(512+128)*4 + (T_in+T_sub)*12 + 57,988]."""
    tt = T_in + T_sub
    return W_ACT * bytes_per_weight + B * (tt * (512 + 128) * 4 + tt * 12 + 57_988)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def lib_stamp() -> str:
    """Stamp that ties an ncu capture to the code it profiled: hash of the kernel sources (nvcc output is not
    byte-reproducible, so a hash of the .so would go stale on every rebuild; `__graft_entry__.build()` rebuilds the library
    whenever a source is newer than it)."""
    try:
        from tacotron2_subword_b200 import build as _b
        if not os.path.isfile(LIB):
            return "missing"
        return _b.source_stamp()
    except Exception:
        return "missing"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_problem():
    from oracle.synth import make_decoder_weights, make_inputs
    w = make_decoder_weights(CFG["attention"], seed=CFG["seed"], gate_bias=CFG["gate_bias"])
    inp = make_inputs(CFG["B"], CFG["T_in"], CFG["T_sub"], 1, seed=CFG["seed"])
    return w, inp


def time_cpu_port(w, inp, frames, repeats, threads):
    """The oracle port of the reference decoder (torch CPU fp32, all host threads)."""
    import torch
    from oracle.decoder_oracle import DecoderOracle
    from oracle.synth import make_dropout_plan
    torch.set_num_threads(threads)
    orc = DecoderOracle(w, CFG["attention"])
    plan = make_dropout_plan(1, frames, frames, CFG["T_in"], CFG["T_sub"], False, seed=CFG["seed"] + 1)
    times = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            mel, *_ = orc.inference(inp["memory"], inp["embeddings"], plan, max_decoder_steps=frames)
            times.append(time.perf_counter() - t0)
            assert mel.shape[2] == frames
    return times


def workload_config():
    return {"workload": "cfg2: free-running Decoder.inference, B=1/GPU, 150 phones + 50 sub-words, "
                        "max_decoder_steps=1000 (gate bias -20 => exactly 1000 frames), dual-stream SMA, default hparams",
            "frames_per_step": CFG["max_steps"], "utterances_per_gpu": 1,
            "l2": "L2 flushed (256 MiB write) between timed steps; per-frame working set 128.9 MB > 126 MB L2",
            "sharding": "one utterance per rank, no collective on the data path"}


def run_reference_arm(args, rank):
    """--impl reference: the reference's own CPU algorithm (oracle port) on all host threads, full 1000-frame utterances,
    same config / metric / unit as the b200 arm."""
    if rank != 0:
        return
    w, inp = make_problem()
    threads = os.cpu_count() or 1
    frames = CFG["max_steps"]
    time_cpu_port(w, inp, 50, max(1, min(args.warmup, 2)), threads)   # warm-up on a short utterance (thread pool, page-in)
    times = time_cpu_port(w, inp, frames, args.steps, threads)
    total = sum(times)
    value = frames * args.steps / total
    line = {
        "impl": "reference", "metric": "mel_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} full {frames}-frame utterances of the workload "
                                   f"(oracle/decoder_oracle.py = CPU restatement of model.Decoder.inference, torch CPU fp32, "
                                   f"{threads} threads, own process)"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def cpu_baseline_subprocess(steps=3):
    """The reference arm in its own process BEFORE this process touches the GPU: identical code and conditions, so
    cpu_baseline and the driver's --impl reference run cannot disagree by construction."""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(steps), "--warmup", "1"],
                             capture_output=True, text=True, timeout=600)
        for ln in out.stdout.splitlines()[::-1]:
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
    except Exception as e:      # noqa: BLE001
        return {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    return {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port", "sample": "no output"}


# ------------------------------------------------------------------------------------------------------------------
# sub-records: the configurations whose scaling has a bottleneck
# ------------------------------------------------------------------------------------------------------------------
def sub_records(torch, dist, dev, rank, world, w_never_stop):
    from oracle.synth import SMA, make_decoder_weights, make_inputs
    from tacotron2_subword_b200 import Decoder, create_hparams
    recs = []

    def timed(fn, reps=2):
        fn()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([min(ts)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dec = Decoder(create_hparams()); dec.load_state_dict(w_never_stop, strict=True); dec = dec.to(dev).eval()
    dec.rng_seed = 7 + rank
    eng = dec._engine(dev)
    # cfg 3: 64 utterances in total, free-running 1000 frames, sharded over the ranks (strong scaling)
    total, T = 64, 1000
    Bp = total // world
    inp = make_inputs(total, 120, 40, 1, seed=3, ragged=True)
    sl = slice(rank * Bp, (rank + 1) * Bp)
    mem, emb = inp["memory"][sl].to(dev), inp["embeddings"][sl].to(dev)
    ml, bl = inp["memory_lengths"][sl].to(dev), inp["bert_lengths"][sl].to(dev)
    T_in, T_sub = int(ml.max()), int(bl.max())
    mem, emb = mem[:, :T_in].contiguous(), emb[:, :T_sub].contiguous()
    with torch.no_grad():
        ms = timed(lambda: dec.inference_batched(mem, emb, ml, bl, max_decoder_steps=T))
    recs.append({"config": "cfg3: batched free-running inference, 64 utterances in total, 120 phones / 40 sub-words, 1000 frames",
                 "scaling": "strong", "utterances_per_gpu": Bp, "path": eng.last_path(), "ms": ms, "us_per_frame_step": 1e3 * ms / T,
                 "value": total * T / (ms * 1e-3), "unit": "frames/s", "collective": "none (utterance sharding)"})
    # cfg 4: GTA shape, 128 utterances in total, teacher-forced 800 frames
    total, T = 128, 800
    Bp = total // world
    inp = make_inputs(total, 160, 53, T, seed=4, ragged=True)
    sl = slice(rank * Bp, (rank + 1) * Bp)
    ml, bl = inp["memory_lengths"][sl].to(dev), inp["bert_lengths"][sl].to(dev)
    T_in, T_sub = int(ml.max()), int(bl.max())
    mem, emb = inp["memory"][sl, :T_in].contiguous().to(dev), inp["embeddings"][sl, :T_sub].contiguous().to(dev)
    mels = inp["mels"][sl].to(dev)
    with torch.no_grad():
        ms = timed(lambda: dec(mem, emb, mels, ml, bl, independent=True))
    recs.append({"config": "cfg4: GTA extraction shape, teacher-forced, 128 utterances in total, 160 / 53, 800 frames",
                 "scaling": "strong", "utterances_per_gpu": Bp, "path": eng.last_path(), "ms": ms, "us_per_frame_step": 1e3 * ms / T,
                 "value": total * T / (ms * 1e-3), "unit": "frames/s", "collective": "none (utterance sharding)"})
    del dec
    # cfg 5: decoder training step, 64 utterances per GPU (weak), forward + backward + gradient all-reduce over NCCL
    B, T = 64, 800
    dect = Decoder(create_hparams()); dect.load_state_dict(make_decoder_weights(SMA, seed=1234), strict=True)
    dect = dect.to(dev).train()
    dect.rng_seed = 11
    n_coll = 0
    if world > 1:
        from tacotron2_subword_b200.distributed import apply_gradient_allreduce
        apply_gradient_allreduce(dect)
    inp = make_inputs(B, 160, 53, T, seed=5 + rank, ragged=True)
    mem, emb = inp["memory"].to(dev).requires_grad_(True), inp["embeddings"].to(dev).requires_grad_(True)
    mels, ml, bl = inp["mels"].to(dev), inp["memory_lengths"].to(dev), inp["bert_lengths"].to(dev)
    target = torch.randn(B, 80, T, device=dev)

    def step():
        dect.zero_grad(set_to_none=True)
        mel, gate, al, alb = dect(mem, emb, mels, ml, bl)
        loss = torch.nn.functional.mse_loss(mel, target) + \
            torch.nn.functional.binary_cross_entropy_with_logits(gate, torch.zeros_like(gate))
        loss.backward()

    ms = timed(step)
    if world > 1:
        n_coll = dect._grad_bucketer.n_collectives
        gsum = torch.stack([p.grad.double().abs().sum() for p in dect.parameters() if p.grad is not None]).sum()
        glist = [torch.zeros_like(gsum) for _ in range(world)]
        dist.all_gather(glist, gsum)
        assert all(torch.equal(g_, glist[0]) for g_ in glist), "gradients differ across ranks after the all-reduce"
    grad_bytes = sum(p.grad.numel() * 4 for p in dect.parameters() if p.grad is not None)
    recs.append({"config": "cfg5: decoder training step (teacher-forced forward + BPTT backward + NCCL gradient all-reduce), "
                           "64 utterances per GPU, 160 / 53, 800 frames, train mode",
                 "scaling": "weak", "utterances_per_gpu": B, "path": dect._engine(dev).last_path(), "ms": ms,
                 "us_per_frame_step": 1e3 * ms / T, "value": world * B * T / (ms * 1e-3), "unit": "frames/s",
                 "collective": f"all-reduce of {grad_bytes / 1e6:.0f} MB of decoder gradients per step over NCCL" if world > 1 else "none (1 GPU)",
                 "nccl_calls_total": n_coll})
    return recs


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-records", action="store_true")
    ap.add_argument("--weights", default="fp32", choices=["fp32", "fp16"],
                    help="storage of the packed LSTM matrices on the latency path (headline = fp32, the reference's dtype)")
    ap.add_argument("--path", default="auto", choices=["auto", "generic", "latency"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    # CPU baseline first, in its own process, before this process creates a CUDA context
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline_subprocess()

    import torch
    import torch.distributed as dist
    from tacotron2_subword_b200 import Decoder, create_hparams

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    w, inp = make_problem()
    hp = create_hparams()
    hp.max_decoder_steps = CFG["max_steps"]
    dec = Decoder(hp)
    dec.load_state_dict(w, strict=True)
    dec = dec.to(dev).eval()
    dec.rng_seed = 2024 + rank               # production mode: Philox prenet dropout in-kernel
    dec.decoder_path, dec.weight_dtype = args.path, args.weights
    frames = CFG["max_steps"]

    mem_host = inp["memory"].pin_memory()
    emb_host = inp["embeddings"].pin_memory()
    mem_dev, emb_dev = mem_host.to(dev), emb_host.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    eng = dec._engine(dev)
    eng.set_profiling(True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def resident_step():
        with torch.no_grad():
            return dec.inference_batched(mem_dev, emb_dev)

    host_out = {}

    def to_host(name, t):
        b = host_out.get(name)
        if b is None or b.shape != t.shape:
            b = host_out[name] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        b.copy_(t, non_blocking=True)
        return b

    def e2e_step():
        with torch.no_grad():
            m = mem_host.to(dev, non_blocking=True)
            e = emb_host.to(dev, non_blocking=True)
            with contextlib.redirect_stdout(sys.stderr):      # the API prints "Warning! Reached max decoder steps"
                mel, gate, al, alb, flag = dec.inference(m, e)
            # everything Decoder.inference returns, into pinned host buffers, one stream synchronisation for the four copies
            outs = tuple(to_host(n, t) for n, t in (("mel", mel), ("gate", gate), ("align", al), ("align_bert", alb)))
            torch.cuda.current_stream(dev).synchronize()
            return outs + (flag,)

    for _ in range(args.warmup):
        out = resident_step()
        assert int(out[4][0]) == frames and int(out[5][0]) == 1
    e2e_step()
    path_taken = eng.last_path()
    mel_ref = out[0].clone()

    # ---------------- value: inputs resident in HBM, device-timed ----------------
    sampler = ClockSampler(local_rank)
    launches0 = eng.launch_count()
    barrier()
    sampler.start()
    step_ms, kern_ms = [], []
    for _ in range(args.steps):
        flush.fill_(1)                       # evict L2 between timed iterations (untimed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        resident_step()
        e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        kern_ms.append(eng.last_kernel_ms())
    barrier()
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) / 1e3
    value = n_gpus * frames * args.steps / total_s

    # ---------------- e2e: pinned host inputs -> API call -> host outputs ----------------
    barrier()
    e2e_ms = []
    for _ in range(args.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        e0.record()
        outs_h = e2e_step()
        e1.record()
        e1.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        e2e_ms.append(max(e0.elapsed_time(e1), wall))   # host-side work is part of the end-to-end call
    barrier()
    e2e_total = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_total, op=dist.ReduceOp.MAX)
    e2e_value = n_gpus * frames * args.steps / (float(e2e_total.item()) / 1e3)
    h2d = mem_host.numel() * 4 + emb_host.numel() * 4
    d2h = sum(t.numel() * 4 for t in outs_h[:4])

    # ---------------- p50 latency without the L2 flush (a server decoding back-to-back utterances) ----------------
    warm_ms = []
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        resident_step()
        e1.record()
        e1.synchronize()
        warm_ms.append(e0.elapsed_time(e1))

    # ---------------- secondary variant: fp16 storage of the three LSTM matrices (same kernel family) ------
    variant = None
    if args.weights == "fp32" and path_taken == "latency":
        dec.weight_dtype = "fp16"
        for _ in range(3):
            out16 = resident_step()
        v_ms = []
        for _ in range(max(3, args.steps // 2)):
            flush.fill_(1)
            resident_step()
            v_ms.append(eng.last_kernel_ms())
        torch.cuda.synchronize(dev)
        err = float((out16[0] - mel_ref).abs().max())
        dec.weight_dtype = "fp32"
        variant = {"weights": "fp16 LSTM matrices (fp32 accumulate, everything else fp32)", "kernel_ms": statistics.mean(v_ms),
                   "us_per_frame": 1e3 * statistics.mean(v_ms) / frames, "frames_per_sec_1gpu": frames / (statistics.mean(v_ms) * 1e-3),
                   "mel_max_abs_vs_fp32_over_1000_free_running_frames": err}

    l2_gbs, hop_ns = eng.measure_machine()
    subs = None
    if not args.no_sub_records:
        subs = sub_records(torch, dist, dev, rank, world, w)

    if rank == 0:
        peak, peak_src = measured_peaks()
        wb = 4 if args.weights == "fp32" else 2
        bpf = (algorithmic_bytes_per_frame(1, CFG["T_in"], CFG["T_sub"], 4) if wb == 4
               else 31_457_280 * 2 + 624_977 * 4 + (algorithmic_bytes_per_frame(1, CFG["T_in"], CFG["T_sub"], 4) - W_ACT * 4))
        k_ms = statistics.mean(kern_ms)
        us_frame = 1e3 * k_ms / frames
        hbm_alg = bpf * frames / (k_ms * 1e-3) / 1e9
        # committed ncu capture of this kernel: only valid for the library it was taken with
        traffic = l2_traffic = None
        traffic_note = "no capture"
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                tj = json.load(f)
            if tj.get("lib_sha256_16") == lib_stamp():
                traffic, l2_traffic = tj.get("dram_bytes_per_launch"), tj.get("l2_bytes_per_launch")
                traffic_note = f"ncu capture {tj.get('source')} (library {tj.get('lib_sha256_16')})"
            else:
                traffic_note = f"stale: capture was taken with library {tj.get('lib_sha256_16')}, loaded library is {lib_stamp()}"
        except Exception:
            pass
        # bytes the latency kernel pulls through L2 per frame: everything that is not resident in shared / tensor memory
        l2_bpf = (l2_traffic / frames) if l2_traffic else (bpf - (24.5e6 + 33.5e6 if wb == 4 else 52.0e6))
        floor_us = LAT_HOPS_PER_FRAME * hop_ns * 1e-3 + l2_bpf / (l2_gbs * 1e9) * 1e6
        is_lat = path_taken == "latency"
        line = {
            "metric": "mel_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": statistics.mean(step_ms), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.weights == "fp32" else "f16-weights/f32-accumulate",
            "data": "synthetic", "config": dict(workload_config(), kernel_path=path_taken, weights=args.weights),
            "latency_ms_p50": statistics.median(step_ms), "latency_ms_p50_no_l2_flush": statistics.median(warm_ms),
            "us_per_frame": us_frame,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "latency_ms_p50": statistics.median(e2e_ms),
                    "outputs_copied_back": "mel, gate, alignments, alignments_bert (everything Decoder.inference returns)"},
            "gpu_launches": launches,
            "roofline": {
                "bound": "l2+latency" if is_lat else "hbm",
                "achieved": (l2_bpf * frames / (k_ms * 1e-3) / 1e9) if is_lat else hbm_alg,
                "peak": l2_gbs if is_lat else peak, "unit": "GB/s",
                "frac": (floor_us / us_frame) if is_lat else hbm_alg / peak,
                "traffic": traffic, "traffic_source": traffic_note,
                "kernel": ("lat::decoder_latency<%d, %s>" % (wb, "true" if wb == 4 else "false")) if is_lat else "decoder_persistent<1>", "kernel_ms": k_ms,
                "model": {"what": "per-frame floor = hops x measured cross-CTA exchange latency + L2 bytes / measured L2 read rate; "
                                  "frac = floor / achieved frame time (achieved / peak are the L2 byte rate and the measured L2 read rate)",
                          "hops_per_frame": LAT_HOPS_PER_FRAME, "hop_ns_measured": hop_ns, "l2_read_gbs_measured": l2_gbs,
                          "l2_bytes_per_frame": l2_bpf, "floor_us_per_frame": floor_us, "achieved_us_per_frame": us_frame},
                "hbm_algorithmic": {"achieved_gbs": hbm_alg, "peak_gbs": peak, "frac": hbm_alg / peak,
                                    "algorithmic_bytes_per_launch": bpf * frames,
                                    "peak_source": f"{peak_src} (MEASURED_PEAKS.json hbm_gbs)",
                                    "note": "not a bound: > 90 % of these bytes are served from shared memory, tensor memory and L2"},
            },
        }
        if variant is not None:
            line["variants"] = [variant]
        if subs is not None:
            line["sub_records"] = subs
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
