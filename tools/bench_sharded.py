"""BASELINE cfg 3 / cfg 4 across the GPUs of one box: utterances are sharded by rank, no collective on the data path
(SURVEY.md 8e); the job time is the slowest rank's device time.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_sharded.py"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from oracle.synth import SMA, make_decoder_weights, make_inputs
from tacotron2_subword_b200 import Decoder, create_hparams
from tacotron2_subword_b200.distributed import shard_range

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", init_method="env://")
w = make_decoder_weights(SMA, seed=1234, gate_bias=-20.0)
dec = Decoder(create_hparams()); dec.load_state_dict(w); dec = dec.cuda().eval(); dec.rng_seed = 7
eng = dec._engine(torch.device("cuda", torch.cuda.current_device()))
rows = []


def run(name, mode, B_total, T_in, T_sub, T):
    """B_total utterances for the whole job; this rank decodes shard_range(B_total, rank, world)."""
    idx = shard_range(B_total, rank, world)
    B = len(idx)
    inp = make_inputs(B, T_in, T_sub, T if mode == "tf" else 1, seed=3 + rank, ragged=True)   # this rank's shard only
    mem, emb = inp["memory"].cuda(), inp["embeddings"].cuda()
    ml, bl = inp["memory_lengths"].cuda(), inp["bert_lengths"].cuda()
    dec.decoder_path, dec.weight_dtype = "auto", "fp16"      # 2 <= B <= 128 per GPU -> tensor path
    if mode == "tf":
        mels = inp["mels"].cuda()
        fn = lambda: dec(mem, emb, mels, ml, bl)
    else:
        fn = lambda: dec.inference_batched(mem, emb, ml, bl, max_decoder_steps=T)
    ts = []
    with torch.no_grad():
        fn(); torch.cuda.synchronize()
        for _ in range(3):
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
    ms = torch.tensor([min(ts)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    if rank == 0:
        r = dict(config=name, n_gpus=world, utterances_total=B_total, utterances_per_gpu=B, frames=T, path=eng.last_path(),
                 ms_slowest_rank=round(ms, 3), frames_per_s=round(B_total * T / (ms * 1e-3)))
        rows.append(r); print(json.dumps(r), flush=True)


run("cfg3 strong: free-running, 64 utterances in total, 120/40, 1000 steps", "fr", 64, 120, 40, 1000)
run("cfg3 weak: free-running, 64 utterances per GPU", "fr", 64 * world, 120, 40, 1000)
run("cfg4: GTA teacher-forced, 128 utterances in total, 160/53, 800 frames", "tf", 128, 160, 53, 800)
run("cfg4 weak: GTA teacher-forced, 128 utterances per GPU", "tf", 128 * world, 160, 53, 800)
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open(f"gpurun_out/bench_sharded_{world}gpu.json", "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
