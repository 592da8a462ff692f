"""Batched synthesis (BERT_Tacotron2.inference_batch) vs the reference's loop shape (one inference() per utterance) through
the same model: 64 utterances, 120 phones + 40 sub-words, gate bias -20 (1000 frames each).  usage: python tools/synth_bench.py [n]"""
import sys, os, json, time, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tacotron2_subword_b200 import BERT_Tacotron2, create_hparams

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(1234)
hp = create_hparams()
model = BERT_Tacotron2(hp).cuda().eval()
with torch.no_grad():
    model.decoder.gate_layer.linear_layer.bias.fill_(-20.0)
model.decoder.max_decoder_steps = 1000
model.decoder.rng_seed = 1
g = torch.Generator().manual_seed(0)
T_ins = torch.randint(60, 121, (n,), generator=g).tolist(); T_ins[0] = 120
seqs = [torch.randint(0, hp.n_symbols, (1, t), generator=g).cuda() for t in T_ins]
subs = [torch.randint(0, hp.sub_n_symbols, (1, max(2, t // 3)), generator=g).cuda() for t in T_ins]
pcls = [torch.randn(1, t, hp.BERT_embedding_dim, generator=g).cuda() for t in T_ins]
bcls = [torch.randn(1, s.shape[1], hp.BERT_embedding_dim, generator=g).cuda() for s in subs]


def wall(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts)


with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
    model.decoder.weight_dtype = "fp16"
    t_batch = wall(lambda: model.inference_batch(seqs, subs, pcls, bcls))
    t_enc = wall(lambda: [model._memories(seqs[i], subs[i], pcls[i], bcls[i]) for i in range(n)])
    model.decoder.weight_dtype = "fp32"
    k = min(n, 8)
    t_loop = wall(lambda: [model.inference(seqs[i], subs[i], pcls[i], bcls[i]) for i in range(k)], reps=2) * n / k
frames = n * 1000
res = dict(utterances=n, frames=frames, batched_s=round(t_batch, 4), batched_frames_per_s=round(frames / t_batch),
           per_utterance_encoding_loop_s=round(t_enc, 4), per_utterance_loop_s=round(t_loop, 4),
           per_utterance_loop_frames_per_s=round(frames / t_loop), speedup=round(t_loop / t_batch, 1))
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/synth_bench.json", "w"), indent=1)
