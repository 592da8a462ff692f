"""CPU, gloo, world_size 2: utterance sharding and the bucketed gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tacotron2_subword_b200.distributed import (apply_gradient_allreduce, reduce_tensor, shard_by_length,
                                                shard_range)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(7, 5)
        self.b = torch.nn.Linear(5, 3)
        self.dead = torch.nn.Linear(4, 4)      # never used: no gradient, like decoder.decoder_rnn_bert
        self.c = torch.nn.Linear(3, 2)

    def forward(self, x):
        return self.c(torch.relu(self.b(torch.relu(self.a(x)))))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)              # ranks start DIFFERENT: broadcast must fix that
        net = _Net()
        apply_gradient_allreduce(net, bucket_mb=1e-4)   # tiny buckets -> several collectives
        w0 = net.a.weight.detach().clone()
        results = []
        for step in range(2):
            torch.manual_seed(1000 + 10 * step + rank)
            x = torch.randn(6, 7)
            net.zero_grad(set_to_none=True)
            net(x).pow(2).sum().backward()
            results.append({n: (None if p.grad is None else p.grad.numpy().copy()) for n, p in net.named_parameters()})
        loss_mean = reduce_tensor(torch.tensor(float(rank + 1)), world)
        # numpy arrays are pickled by value (torch tensors would travel as shared-memory handles)
        q.put((rank, w0.numpy().copy(), results, float(loss_mean), net._grad_bucketer.n_collectives,
               len(net._grad_bucketer.buckets)))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_matches_mean_of_local_grads():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, w0a, res_a, lm_a, ncoll, nb), (_, w0b, res_b, lm_b, _, _) = out
    assert (w0a == w0b).all(), "rank-0 parameters must be broadcast at wrap time"
    assert lm_a == lm_b == 1.5
    assert nb >= 3 and ncoll >= 2 * 3, "several buckets, each reduced every step"
    # reference result: average of the two ranks' LOCAL gradients, recomputed serially
    torch.manual_seed(100)
    ref = _Net()
    for step in range(2):
        local = []
        for rank in range(world):
            torch.manual_seed(1000 + 10 * step + rank)
            x = torch.randn(6, 7)
            ref.zero_grad(set_to_none=True)
            ref(x).pow(2).sum().backward()
            local.append({n: (None if p.grad is None else p.grad.clone()) for n, p in ref.named_parameters()})
        for n in local[0]:
            if local[0][n] is None:
                assert res_a[step][n] is None and res_b[step][n] is None     # dead params stay grad-less
                continue
            want = (local[0][n] + local[1][n]) / 2
            assert torch.allclose(torch.from_numpy(res_a[step][n]), want, atol=1e-6), n
            assert (res_a[step][n] == res_b[step][n]).all(), n


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 64, 129):
        for w in (1, 2, 4, 8):
            got = [i for r in range(w) for i in shard_range(n, r, w)]
            assert got == list(range(n))
            sizes = [len(shard_range(n, r, w)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


def test_shard_by_length_is_balanced_and_complete():
    g = torch.Generator().manual_seed(0)
    lengths = torch.randint(60, 160, (64,), generator=g).tolist()
    for w in (2, 4, 8):
        shards = shard_by_length(lengths, w)
        assert sorted(i for s in shards for i in s) == list(range(64))
        assert {len(s) for s in shards} == {64 // w}
        loads = [sum(lengths[i] for i in s) for s in shards]
        assert (max(loads) - min(loads)) / max(loads) < 0.03


# ---------------------------------------------------------------------------------------------------------------
# early path: a module that produces several gradients inside ONE autograd node (the decoder's hand-written BPTT)
# publishes them one by one; large ones are averaged in place while the node is still computing the others
# ---------------------------------------------------------------------------------------------------------------
class _TwoGrads(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, big, small):
        ctx.mod = mod
        ctx.save_for_backward(x)
        return x @ big + small.sum()

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        mod = ctx.mod
        d_big = x.t() @ g
        if getattr(mod, "_grad_ready", None) is not None:
            mod._grad_ready(mod.big, d_big)           # published before the node returns (once per node: sub-batches publish twice)
        d_small = g.sum() * torch.ones_like(mod.small)
        if getattr(mod, "_grad_ready", None) is not None:
            mod._grad_ready(mod.small, d_small)       # below the size threshold: stays on the bucket path
        if getattr(mod, "_grad_ready_finish", None) is not None:
            mod._grad_ready_finish()
        return None, None, d_big, d_small


class _FakeDecoder(torch.nn.Module):
    """Looks like the decoder to GradientBucketer (it keys on _weight_tensors / decoder_rnn)."""

    def __init__(self):
        super().__init__()
        self.big = torch.nn.Parameter(torch.randn(16, 8))
        self.small = torch.nn.Parameter(torch.randn(3))
        self.decoder_rnn = torch.nn.Identity()

    def _weight_tensors(self):
        return [self.big, self.small]

    def forward(self, x):
        if getattr(self, "sub_batches", False):       # two nodes per pass, as Decoder._tf_resized produces for B > 128
            h = x.shape[0] // 2
            return torch.cat([_TwoGrads.apply(self, x[:h], self.big, self.small), _TwoGrads.apply(self, x[h:], self.big, self.small)])
        return _TwoGrads.apply(self, x, self.big, self.small)


def _early_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(5)
        net = _FakeDecoder()
        apply_gradient_allreduce(net)
        net._grad_bucketer.early_min_bytes = 256      # `big` (512 B) takes the early path, `small` (12 B) does not
        out = []
        for step in range(3):
            net.sub_batches = step == 2               # third step: the same parameter is published by two nodes of one pass
            torch.manual_seed(50 + step * 10 + rank)
            x = torch.randn(4, 16)
            net.zero_grad(set_to_none=True)
            net(x).pow(2).sum().backward()
            out.append((net.big.grad.numpy().copy(), net.small.grad.numpy().copy()))
        q.put((rank, out, net._grad_bucketer.n_collectives))
    finally:
        dist.destroy_process_group()


def test_early_published_gradients_are_averaged_once():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_early_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(5)
    ref = _FakeDecoder()
    for step in range(3):
        local = []
        for rank in range(world):
            torch.manual_seed(50 + step * 10 + rank)
            x = torch.randn(4, 16)
            ref.zero_grad(set_to_none=True)
            ref(x).pow(2).sum().backward()
            local.append((ref.big.grad.clone(), ref.small.grad.clone()))
        for k in range(2):
            want = (local[0][k] + local[1][k]) / 2
            for rank in range(world):
                assert torch.allclose(torch.from_numpy(out[rank][1][step][k]), want, atol=1e-5), (step, k, rank)
    assert out[0][2] == 7       # per step: one early all-reduce per publishing node (big) + one bucket (small); two nodes in step 3
