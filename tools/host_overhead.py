"""Where the time of one Decoder.inference_batched call goes outside the persistent kernel (cfg 2, B = 1):
device-timed call vs kernel, wall clock, and a cProfile of the host side.
usage: python tools/host_overhead.py [calls]"""
import sys, os, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_problem, CFG
from tacotron2_subword_b200 import Decoder, create_hparams

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
w, inp = make_problem()
hp = create_hparams(); hp.max_decoder_steps = CFG["max_steps"]
dec = Decoder(hp); dec.load_state_dict(w); dec = dec.cuda().eval(); dec.rng_seed = 1
eng = dec._engine(torch.device("cuda", 0)); eng.set_profiling(True)
mem, emb = inp["memory"].cuda(), inp["embeddings"].cuda()
with torch.no_grad():
    for _ in range(3):
        dec.inference_batched(mem, emb)
    torch.cuda.synchronize()
    dev_ms, wall_ms, k_ms = [], [], []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        dec.inference_batched(mem, emb)
        e1.record(); e1.synchronize()
        wall_ms.append((time.perf_counter() - t0) * 1e3); dev_ms.append(e0.elapsed_time(e1)); k_ms.append(eng.last_kernel_ms())
    med = lambda v: sorted(v)[len(v) // 2]
    print(f"call (device events) {med(dev_ms):.3f} ms, wall {med(wall_ms):.3f} ms, persistent kernel {med(k_ms):.3f} ms, "
          f"outside the kernel {med(dev_ms) - med(k_ms):.3f} ms")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        dec.inference_batched(mem, emb)
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
    print(s.getvalue()[:6000])
