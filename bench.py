#!/usr/bin/env python
"""bench.py -- mel frames/s of the Tacotron2 dual-stream decoder hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): free-running
``Decoder.inference``, batch 1, 150-phone / 50-sub-word synthetic memory, max_decoder_steps=1000,
default hparams (dual-stream, StepwiseMonotonicAttention), gate bias -20 so every utterance runs
exactly 1000 frames.  One "step" = one utterance.  N GPUs = N independent utterances, one per rank
(utterance sharding, no data-path collective; weak scaling).

Printed JSON line (rank 0): value = frames/s with inputs resident in HBM (device-timed, CUDA
events, max over ranks); e2e = the same metric through the Python drop-in API with pinned HOST
inputs (H2D + D2H inside the timed region); roofline = algorithmic bytes of the persistent kernel /
its event-timed duration vs the measured HBM peak; cpu_baseline = the CPU oracle port of the
reference decoder timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import contextlib
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
REF_SAMPLE_FRAMES = 250   # frames per step for the CPU reference arm (bounded sample of the 1000-frame workload)
W_ACT = 32_082_257        # weights touched per frame, dual-stream SMA


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout when
# NCCL_DEBUG is set on the box, the reference-style decoder prints a warning at max_decoder_steps), so the process's
# stdout is pointed at stderr for the whole run and the result line goes to the saved descriptor.
_REAL_STDOUT = None


def _claim_stdout() -> None:
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
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


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


def run_reference_arm(args, rank):
    """--impl reference: the reference's own CPU algorithm (oracle port; the reference is Python
    and does not travel to the GPU box) on all host threads, same config/metric/unit."""
    if rank != 0:
        return
    w, inp = make_problem()
    threads = os.cpu_count() or 1
    time_cpu_port(w, inp, 20, max(1, args.warmup), threads)  # warm-up on a short utterance
    times = time_cpu_port(w, inp, REF_SAMPLE_FRAMES, args.steps, threads)
    total = sum(times)
    value = REF_SAMPLE_FRAMES * args.steps / total
    line = {
        "impl": "reference", "metric": "mel_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} x {REF_SAMPLE_FRAMES}-frame prefix of the 1000-frame utterance "
                                   f"(oracle/decoder_oracle.py, torch CPU fp32, {threads} threads)"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config():
    return {"workload": "cfg2: free-running Decoder.inference, B=1/GPU, 150 phones + 50 sub-words, "
                        "max_decoder_steps=1000 (gate bias -20 => exactly 1000 frames), dual-stream SMA, default hparams",
            "frames_per_step": CFG["max_steps"], "utterances_per_gpu": 1, "weights": "fp32 (reference layouts)",
            "l2": "L2 flushed (256 MiB write) between timed steps; per-frame working set 128.9 MB > 126 MB L2",
            "sharding": "one utterance per rank, no collective on the data path"}


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    def e2e_step():
        with torch.no_grad():
            m = mem_host.to(dev, non_blocking=True)
            e = emb_host.to(dev, non_blocking=True)
            with contextlib.redirect_stdout(sys.stderr):      # the API prints "Warning! Reached max decoder steps"
                mel, gate, al, alb, flag = dec.inference(m, e)
            return mel.cpu(), gate.cpu(), flag

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
        mel_h, gate_h, flag = e2e_step()
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
    d2h = mel_h.numel() * 4 + gate_h.numel() * 4

    # ---------------- p50 latency without the L2 flush (a server decoding back-to-back utterances; SURVEY.md 8d asks for both) ----
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
    if rank == 0:
        peak, peak_src = measured_peaks()
        wb = 4 if args.weights == "fp32" else 2
        bpf = (algorithmic_bytes_per_frame(1, CFG["T_in"], CFG["T_sub"], 4) if wb == 4
               else 31_457_280 * 2 + 624_977 * 4 + (algorithmic_bytes_per_frame(1, CFG["T_in"], CFG["T_sub"], 4) - W_ACT * 4))
        k_ms = statistics.mean(kern_ms)
        achieved = bpf * frames / (k_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": "mel_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": statistics.mean(step_ms), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.weights == "fp32" else "f16-weights/f32-accumulate",
            "data": "synthetic", "config": dict(workload_config(), kernel_path=path_taken, weights=args.weights),
            "latency_ms_p50": statistics.median(step_ms), "latency_ms_p50_no_l2_flush": statistics.median(warm_ms),
            "us_per_frame": 1e3 * k_ms / frames,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "latency_ms_p50": statistics.median(e2e_ms)},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic if (args.weights == "fp32" and path_taken == "latency") else None,
                         "kernel": ("lat::decoder_latency<%d>" % wb) if path_taken == "latency" else "decoder_persistent<1>",
                         "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": bpf * frames, "peak_source": f"{peak_src} (MEASURED_PEAKS.json hbm_gbs)"},
        }
        if variant is not None:
            line["variants"] = [variant]
        if n_gpus == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            time_cpu_port(w, inp, 20, 1, threads)
            t = time_cpu_port(w, inp, frames, 2, threads)
            line["cpu_baseline"] = {"value": frames * len(t) / sum(t), "unit": "frames/s", "cores": threads, "kind": "port",
                                    "sample": f"{len(t)} full {frames}-frame utterances of the same workload "
                                              f"(oracle/decoder_oracle.py = CPU restatement of model.Decoder.inference, "
                                              f"torch CPU fp32, {threads} threads)"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
