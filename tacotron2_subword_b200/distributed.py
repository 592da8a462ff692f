"""Data-parallel plumbing for the decoder path: utterance sharding (inference / GTA: no collective)
and a bucketed, overlapped gradient all-reduce for training (the only collective on this path).

Drop-in for /root/reference/distributed.py:132-179 (``apply_gradient_allreduce``) and
train.py:23-42 (``reduce_tensor``, ``init_distributed``).  Differences, all deliberate:

  * the reference flattens ALL gradients into one 207.5 MB buffer after the whole backward has
    finished (one ``torch.cat`` + one all-reduce + per-tensor copies, zero overlap).  Here parameters
    are grouped into buckets in reverse registration order (~ the order autograd finishes them);
    a bucket's all-reduce is launched asynchronously the moment its last gradient is accumulated,
    so NCCL (NVLink 5 / NVSwitch, in-switch reduction when available) overlaps the rest of backward;
  * the 1/world_size scale is folded into the bucket before the reduce;
  * the initial parameter broadcast is one call per bucket instead of one per tensor (~140 calls);
  * parameters that never receive a gradient (``decoder.decoder_rnn_bert.*`` is dead in decode(),
    model.py:375-378) are skipped exactly like the reference's ``param.grad is not None`` test.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch.autograd import Variable


def init_distributed(hparams, n_gpus: int, rank: int, group_name: Optional[str] = None) -> None:
    """train.py:30-42: one process per GPU, NCCL, TCP rendezvous from hparams.dist_url."""
    assert torch.cuda.is_available(), "Distributed mode requires CUDA."
    torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group(backend=hparams.dist_backend, init_method=hparams.dist_url,
                            world_size=n_gpus, rank=rank, group_name=group_name or "")


def reduce_tensor(tensor: torch.Tensor, n_gpus: int) -> torch.Tensor:
    """train.py:23-27: mean over ranks of a (scalar) tensor."""
    rt = tensor.clone()
    dist.all_reduce(rt, op=dist.ReduceOp.SUM)
    rt /= n_gpus
    return rt


def shard_range(n_items: int, rank: int, world_size: int) -> range:
    """Contiguous utterance shard of rank (sizes differ by at most one)."""
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def shard_by_length(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Length-balanced sharding for batched inference / GTA: sort by length (the reference's
    collate_fn already sorts, data_utils.py:146-160) and deal round-robin in a serpentine so every
    rank gets the same count (+-1) and nearly the same number of frames; utterances are independent,
    so no collective is needed afterwards."""
    order = sorted(range(len(lengths)), key=lambda i: -int(lengths[i]))
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for k, idx in enumerate(order):
        lap, pos = divmod(k, world_size)
        shards[pos if lap % 2 == 0 else world_size - 1 - pos].append(idx)
    return shards


class _Bucket:
    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params
        self.pending = 0
        self.ready: List[torch.nn.Parameter] = []
        self.flat: Optional[torch.Tensor] = None
        self.work = None
        self.launched = False

    def reset(self):
        self.pending = len(self.params)
        self.ready = []
        self.flat, self.work, self.launched = None, None, False


class GradientBucketer:
    """Bucketed asynchronous mean-all-reduce of ``module``'s gradients."""

    def __init__(self, module: torch.nn.Module, bucket_bytes: int = 32 << 20, group=None):
        self.module = module
        self.group = group
        self.world = dist.get_world_size(group)
        params = [p for p in module.parameters() if p.requires_grad]
        self.buckets: List[_Bucket] = []
        cur: List[torch.nn.Parameter] = []
        size = 0
        for p in reversed(params):                       # ~ the order gradients become final
            if cur and (size + p.numel() * p.element_size() > bucket_bytes or p.dtype != cur[0].dtype):
                self.buckets.append(_Bucket(cur))
                cur, size = [], 0
            cur.append(p)
            size += p.numel() * p.element_size()
        if cur:
            self.buckets.append(_Bucket(cur))
        self._bucket_of: Dict[int, _Bucket] = {id(p): b for b in self.buckets for p in b.params}
        self._callback_queued = False
        self.n_collectives = 0
        self._early: List = []                 # (work handle) of gradients reduced before autograd saw them
        self._prereduced = set()               # ids of parameters whose gradient of this pass is already the mean
        self.early_min_bytes = 1 << 20
        for b in self.buckets:
            b.reset()
        for p in params:
            p.register_post_accumulate_grad_hook(self._on_grad)
        # modules that produce several gradients in one autograd node (the decoder's hand-written BPTT) publish them one by
        # one through these two attributes; see model._DecoderTF.backward
        for m in module.modules():
            if hasattr(m, "_weight_tensors") and hasattr(m, "decoder_rnn"):
                m._grad_ready = self.early_reduce
                m._grad_ready_finish = self.early_finish

    # -- early path: reduce a finished gradient in place while its producer keeps computing the others ------------------
    def early_reduce(self, p: torch.nn.Parameter, g: torch.Tensor) -> None:
        if id(p) not in self._bucket_of or g.numel() * g.element_size() < self.early_min_bytes or p.grad is not None:
            return                              # small tensors ride in their bucket; accumulating passes take the normal path
        g.mul_(1.0 / self.world)
        self._early.append(dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self._prereduced.add(id(p))
        self.n_collectives += 1

    def early_finish(self) -> None:
        for wk in self._early:
            wk.wait()
        self._early = []

    # -- hooks ------------------------------------------------------------------------------
    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if not self._callback_queued:
            self._callback_queued = True
            Variable._execution_engine.queue_callback(self._finish)   # end of this backward pass
        b = self._bucket_of[id(p)]
        if id(p) in self._prereduced:           # already averaged in place by early_reduce
            self._prereduced.discard(id(p))
        else:
            b.ready.append(p)
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def _launch(self, b: _Bucket) -> None:
        grads = [p.grad for p in b.ready if p.grad is not None]
        if not grads:
            b.launched = True
            return
        b.flat = torch.cat([g.reshape(-1) for g in grads]) if len(grads) > 1 else grads[0].reshape(-1).clone()
        b.flat.mul_(1.0 / self.world)                                  # mean folded into the pre-scale
        b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        b.launched = True
        self.n_collectives += 1

    def _finish(self) -> None:
        # buckets holding parameters that got no gradient this step never filled up: flush them now
        for b in self.buckets:
            if not b.launched:
                self._launch(b)
        for b in self.buckets:
            if b.work is not None:
                b.work.wait()
                off = 0
                for p in b.ready:
                    if p.grad is None:
                        continue
                    n = p.grad.numel()
                    p.grad.copy_(b.flat[off:off + n].view_as(p.grad))
                    off += n
            b.reset()
        self._callback_queued = False


def broadcast_parameters(module: torch.nn.Module, src: int = 0, bucket_bytes: int = 64 << 20, group=None) -> int:
    """Rank ``src``'s parameters and buffers to everyone, coalesced (distributed.py:138-141 does one
    broadcast per tensor).  Returns the number of collectives issued."""
    tensors = [t for t in module.state_dict().values() if torch.is_tensor(t)]
    by_dtype: Dict[torch.dtype, List[torch.Tensor]] = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    calls = 0
    for dt, ts in by_dtype.items():
        chunk: List[torch.Tensor] = []
        size = 0
        def flush():
            nonlocal chunk, size, calls
            if not chunk:
                return
            flat = torch.cat([t.reshape(-1) for t in chunk])
            dist.broadcast(flat, src, group=group)
            off = 0
            for t in chunk:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()
            calls += 1
            chunk, size = [], 0
        for t in ts:
            if size + t.numel() * t.element_size() > bucket_bytes:
                flush()
            chunk.append(t)
            size += t.numel() * t.element_size()
        flush()
    return calls


def apply_gradient_allreduce(module: torch.nn.Module, bucket_mb: float = 32.0, group=None) -> torch.nn.Module:
    """Same name / contract as the reference (distributed.py:132-179): broadcast rank 0's state, then
    average gradients across ranks after every backward.  The module's class is unchanged."""
    if not dist.is_initialized():
        raise RuntimeError("apply_gradient_allreduce needs an initialised process group")
    with torch.no_grad():
        broadcast_parameters(module, 0, group=group)
    if not hasattr(module, "_grad_bucketer"):
        module._grad_bucketer = GradientBucketer(module, int(bucket_mb * (1 << 20)), group=group)
    return module
