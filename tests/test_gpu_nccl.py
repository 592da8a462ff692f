"""-m gpu, 2 ranks over NCCL (skipped on a box with one GPU): the decoder's hand-written backward publishes its gradients to
the all-reduce while the remaining contractions are still running (distributed.GradientBucketer.early_reduce); the result
must be the mean of the two ranks' LOCAL gradients and identical on both ranks (/root/reference/distributed.py:132-179)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle.synth import SMA, make_decoder_weights, make_inputs
    from tacotron2_subword_b200 import Decoder, create_hparams
    from tacotron2_subword_b200.distributed import apply_gradient_allreduce

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        B, T, T_in, T_sub = 8, 6, 30, 10
        w = make_decoder_weights(SMA, seed=77)

        def build():
            d = Decoder(create_hparams())
            d.load_state_dict(w)
            d = d.cuda().train()
            d.rng_seed = 4321 + rank          # same dropout draws in the local and in the reduced pass
            return d

        inp = make_inputs(B, T_in, T_sub, T, seed=900 + rank, ragged=True)     # every rank has its own utterances
        args = (inp["memory"].cuda(), inp["embeddings"].cuda(), inp["mels"].cuda(), inp["memory_lengths"].cuda(),
                inp["bert_lengths"].cuda())

        def grads_of(d):
            d.zero_grad(set_to_none=True)
            outs = d(*args)
            (outs[0].square().mean() + outs[1].square().mean() + outs[2].sum() * 1e-3).backward()
            torch.cuda.synchronize()
            return {n: p.grad.detach().clone() for n, p in d.named_parameters() if p.grad is not None}

        local = grads_of(build())
        dec = build()
        apply_gradient_allreduce(dec, bucket_mb=8.0)
        n0 = dec._grad_bucketer.n_collectives
        got = grads_of(dec)
        n_coll = dec._grad_bucketer.n_collectives - n0
        worst, n_early = 0.0, 0
        for n, g in local.items():
            parts = [torch.empty_like(g) for _ in range(world)]
            dist.all_gather(parts, g)
            want = sum(parts) / world
            scale = float(want.abs().max()) or 1.0
            worst = max(worst, float((got[n] - want).abs().max()) / scale)
            same = [torch.empty_like(g) for _ in range(world)]
            dist.all_gather(same, got[n])
            assert torch.equal(same[0], same[1]), f"{n}: ranks disagree after the all-reduce"
            n_early += g.numel() * 4 >= dec._grad_bucketer.early_min_bytes
        q.put((rank, worst, n_coll, n_early, len(local)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_decoder_gradients_allreduced_over_nccl_match_mean_of_local_gradients():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, worst, n_coll, n_early, n_grads in out:
        assert n_grads == 26                                   # every live decoder parameter (decoder_rnn_bert is dead)
        # early (per-tensor, in place) reductions for the large tensors + at least one bucket for the small ones
        assert n_early >= 6 and n_coll >= n_early + 1, (n_coll, n_early)
        # same Philox seed, same kernels: the only difference is the order of the fp32 atomics inside the contractions
        assert worst < 5e-3, f"rank {rank}: all-reduced gradient off by {worst:.2e} of its scale"
