"""End-to-end throughput of the batched GTA driver (tacotron2_subword_b200/gta.py) on synthetic decoder-level items,
next to the reference's procedure (GTA.py:35-61: one utterance at a time, batch 1, a host sync + np.save per utterance)
run through the same CUDA decoder.  usage: python tools/gta_bench.py [n_utterances] [max_batch]"""
import sys, os, json, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle.synth import SMA, make_decoder_weights
from tacotron2_subword_b200 import Decoder, create_hparams
from tacotron2_subword_b200.gta import GtaItem, gta_extract

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
max_batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
w = make_decoder_weights(SMA, seed=1234)
dec = Decoder(create_hparams()); dec.load_state_dict(w); dec = dec.cuda().eval()
dec.weight_dtype, dec.rng_seed = "fp16", 5
g = torch.Generator().manual_seed(0)
items = []
for i in range(n):
    T = int(torch.randint(300, 801, (1,), generator=g))
    T_in = max(20, T // 5)
    items.append(GtaItem(f"utt_{i:05d}", (0.5 * torch.randn(T_in, 512, generator=g)).pin_memory(),
                         torch.randn(80, T, generator=g).pin_memory(), (0.5 * torch.randn(max(7, T_in // 3), 512, generator=g)).pin_memory()))
frames = sum(int(it.mel.shape[1]) for it in items)
out = tempfile.mkdtemp(prefix="gta_")
try:
    gta_extract(dec, items[:max_batch], out, max_batch=max_batch)          # warm-up (weight repack, graph, allocator)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    files = gta_extract(dec, items, out, max_batch=max_batch)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    assert len(files) == n
    # the reference's loop shape: batch 1, sync + save per utterance (on a subset, scaled)
    sub = items[: min(n, 48)]
    sub_frames = sum(int(it.mel.shape[1]) for it in sub)
    with torch.no_grad():
        it = sub[0]
        dec(it.memory[None].cuda(), it.embeddings[None].cuda(), it.mel[None].cuda(), torch.tensor([it.memory.shape[0]]).cuda(),
            torch.tensor([it.embeddings.shape[0]]).cuda())
        torch.cuda.synchronize(); t1 = time.perf_counter()
        for it in sub:
            mel = dec(it.memory[None].cuda(), it.embeddings[None].cuda(), it.mel[None].cuda(),
                      torch.tensor([it.memory.shape[0]]).cuda(), torch.tensor([it.embeddings.shape[0]]).cuda())[0]
            np.save(os.path.join(out, "b1_" + it.name), mel.cpu().numpy())
        dt1 = time.perf_counter() - t1
    res = dict(utterances=n, frames=frames, max_batch=max_batch, batched_s=round(dt, 3), batched_frames_per_s=round(frames / dt),
               batched_utterances_per_s=round(n / dt, 1), batch1_loop_frames_per_s=round(sub_frames / dt1),
               batch1_loop_sample=len(sub), speedup=round((frames / dt) / (sub_frames / dt1), 1))
    print(json.dumps(res), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/gta_bench.json", "w"), indent=1)
finally:
    shutil.rmtree(out, ignore_errors=True)
