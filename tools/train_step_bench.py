"""BASELINE cfg 5: decoder training step (teacher-forced forward + backward [+ gradient all-reduce]) with synthetic
upstream gradients.  B is per GPU (weak scaling).
usage: python tools/train_step_bench.py [B] [T] [--cpu]   (defaults 64 800; --cpu also times the CPU oracle on 16 frames)
       python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/train_step_bench.py"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.synth import SMA, make_decoder_weights, make_inputs, make_dropout_plan
from tacotron2_subword_b200 import Decoder, create_hparams

args = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(args[0]) if len(args) > 0 else 64
T = int(args[1]) if len(args) > 1 else 800
T_in, T_sub = 160, 53
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
w = make_decoder_weights(SMA, seed=1234)
dec = Decoder(create_hparams()); dec.load_state_dict(w); dec = dec.cuda().train()
dec.decoder_path, dec.weight_dtype = "tensor", "fp16"
if world > 1:
    import torch.distributed as dist
    from tacotron2_subword_b200.distributed import apply_gradient_allreduce
    dist.init_process_group("nccl", init_method="env://")
    apply_gradient_allreduce(dec)          # broadcast + bucketed all-reduce hooks (distributed.py:132-179)
inp = make_inputs(B, T_in, T_sub, T, seed=3 + rank, ragged=True)
mem, emb, mels = inp["memory"].cuda().requires_grad_(True), inp["embeddings"].cuda().requires_grad_(True), inp["mels"].cuda()
ml, bl = inp["memory_lengths"].cuda(), inp["bert_lengths"].cuda()
target = torch.randn(B, 80, T, device="cuda")


def step():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    dec.zero_grad(set_to_none=True)
    ev[0].record()
    mel, gate, al, alb = dec(mem, emb, mels, ml, bl)
    loss = torch.nn.functional.mse_loss(mel, target) + torch.nn.functional.binary_cross_entropy_with_logits(gate, torch.zeros_like(gate))
    ev[1].record()
    loss.backward()       # with world > 1 the bucketed all-reduce runs inside / right after this
    ev[2].record()
    ev[2].synchronize()
    return ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), float(loss)


step()
res = [step() for _ in range(3)]
fw = min(r[0] for r in res); bw = min(r[1] for r in res)
if world > 1:
    tmax = torch.tensor([fw, bw], device="cuda")
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)       # a step is as slow as its slowest rank
    fw, bw = float(tmax[0]), float(tmax[1])
    gsum = torch.stack([p.grad.abs().sum() for p in dec.parameters() if p.grad is not None]).sum()
    glist = [torch.zeros_like(gsum) for _ in range(world)]
    dist.all_gather(glist, gsum)
    assert all(torch.equal(g_, glist[0]) for g_ in glist), "gradients differ across ranks after the all-reduce"
if rank != 0:
    dist.destroy_process_group()
    sys.exit(0)
out = dict(config=f"cfg5 train step B={B}/GPU T={T} {T_in}/{T_sub} SMA train-mode", n_gpus=world, forward_ms=round(fw, 2),
           backward_ms=round(bw, 2), step_ms=round(fw + bw, 2), frames_per_s=round(world * B * T / ((fw + bw) * 1e-3)),
           us_per_frame_fwd=round(1e3 * fw / T, 1),
           us_per_frame_bwd=round(1e3 * bw / T, 1), loss=res[-1][2], peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2**30, 2))
print(json.dumps(out), flush=True)

if "--profile" in sys.argv:          # kernel-time table of one step (CUPTI through torch.profiler; not a bench number)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:28]
    tot = sum(e.device_time_total for e in prof.key_averages())
    print(f"kernel time of one step: {tot / 1e3:.2f} ms")
    for e in rows:
        print(f"{e.device_time_total / 1e3:9.3f} ms {e.count:6d} x {e.device_time_total / max(1, e.count):9.1f} us  {e.key[:90]}")

if "--cpu" in sys.argv:
    from oracle.decoder_oracle import DecoderOracle
    Tc = 16
    torch.set_num_threads(os.cpu_count())
    orc = DecoderOracle(w, SMA)
    orc.w = {k: v.clone().requires_grad_(True) for k, v in orc.w.items()}
    ci = make_inputs(B, T_in, T_sub, Tc, seed=3, ragged=True)
    plan = make_dropout_plan(B, Tc + 1, Tc, T_in, T_sub, True, seed=4)
    t0 = time.time()
    o = orc.forward(ci["memory"], ci["embeddings"], ci["mels"], ci["memory_lengths"], ci["bert_lengths"], plan, training=True)
    (o[0].pow(2).mean() + o[1].mean()).backward()
    dt = time.time() - t0
    cpu = dict(cpu_oracle_frames=Tc, cpu_s=round(dt, 2), cpu_frames_per_s=round(B * Tc / dt), cores=os.cpu_count())
    print(json.dumps(cpu), flush=True)
    out.update(cpu)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/train_step_{world}gpu.json", "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
