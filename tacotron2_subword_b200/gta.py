"""Batched ground-truth-aligned (GTA) mel extraction for vocoder fine-tuning.

The reference does this one utterance at a time (/root/reference/GTA.py:35-61): eval-mode model, teacher-forced
``Decoder.forward`` at batch 1, ``np.save(<name>, mel_outputs)`` with ``mel_outputs`` of shape [1, n_mel, T] (the
decoder output *before* the postnet), one host synchronisation per utterance.  Here utterances are bucketed by length,
decoded together with ``Decoder.forward(..., independent=True)`` -- every row is exactly the batch-1 result on its own
un-padded memory, so batching does not change what is written -- truncated to their own length and handed to a
background writer, so the GPU never waits for the disk.  Across GPUs the utterance list is sharded by rank
(``distributed.shard_by_length``); there is no collective.

Only the decoder is batched: the encoder of the reference model is not batch-invariant (its convolutions see the
padding symbols), so callers encode each utterance on its own (``encode_tacotron2``) and pass decoder-level items.
"""
from __future__ import annotations

import os
import queue
import threading
from dataclasses import dataclass
from typing import Callable, Iterable, List, Optional, Sequence

import numpy as np
import torch

from .distributed import shard_by_length


@dataclass
class GtaItem:
    """One utterance at the decoder boundary."""
    name: str                                  # output file stem (GTA.py:39: basename of the text file)
    memory: torch.Tensor                       # [T_in, enc]   encoder output (char / phoneme stream)
    mel: torch.Tensor                          # [n_mel, T]    ground-truth mel (teacher-forcing input)
    embeddings: Optional[torch.Tensor] = None  # [T_sub, enc]  sub-word stream (BERT_Tacotron2 only)


def plan_batches(n_frames: Sequence[int], max_batch: int, max_frames_per_batch: Optional[int] = None) -> List[List[int]]:
    """Group utterance indices into batches of similar length (longest first, like the reference's collate,
    data_utils.py:146-160).  A batch costs max(T) frame-steps for all its rows, so mixing lengths wastes work;
    ``max_frames_per_batch`` bounds rows x max(T) (activation memory)."""
    if max_batch < 1:
        raise ValueError("max_batch must be >= 1")
    order = sorted(range(len(n_frames)), key=lambda i: (-int(n_frames[i]), i))
    batches: List[List[int]] = []
    cur: List[int] = []
    for i in order:
        t_max = int(n_frames[cur[0]]) if cur else int(n_frames[i])
        if cur and (len(cur) >= max_batch or (max_frames_per_batch and (len(cur) + 1) * t_max > max_frames_per_batch)):
            batches.append(cur)
            cur = []
        cur.append(i)
    if cur:
        batches.append(cur)
    return batches


class AsyncNpyWriter:
    """``np.save`` on a background thread.  ``put`` never blocks on the disk unless ``max_pending`` arrays are waiting;
    ``close`` joins the thread and re-raises the first I/O error."""

    def __init__(self, out_dir: str, max_pending: int = 256):
        os.makedirs(out_dir, exist_ok=True)
        self.out_dir = out_dir
        self._q: "queue.Queue" = queue.Queue(maxsize=max_pending)
        self._err: Optional[BaseException] = None
        self.written: List[str] = []
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            name, arr = item
            try:
                if self._err is None:
                    path = os.path.join(self.out_dir, name)
                    np.save(path, arr)                   # np.save appends ".npy" when missing, as in GTA.py:61
                    self.written.append(path if path.endswith(".npy") else path + ".npy")
            except BaseException as e:                   # keep draining so producers never dead-lock
                self._err = e

    def put(self, name: str, arr: np.ndarray) -> None:
        if self._err is not None:
            raise self._err
        self._q.put((name, arr))

    def close(self) -> List[str]:
        self._q.put(None)
        self._thread.join()
        if self._err is not None:
            raise self._err
        return self.written


def encode_tacotron2(model, text: torch.Tensor) -> torch.Tensor:
    """Encoder output of ONE utterance for the single-stream compat model (GTA.py:51-59 + model.py Tacotron2.forward):
    text LongTensor [T_in] -> memory [T_in, enc].  Batch 1 on purpose (see module docstring)."""
    dev = next(model.parameters()).device
    text = text.to(dev).long().unsqueeze(0)
    lengths = torch.tensor([text.shape[1]], device=dev)
    emb = model.embedding(text).transpose(1, 2)
    return model.encoder(emb, lengths)[0]


@torch.no_grad()
def gta_extract(decoder, items: Sequence[GtaItem], out_dir: str, max_batch: int = 64, rank: int = 0, world_size: int = 1,
                max_frames_per_batch: Optional[int] = None, on_batch: Optional[Callable[[int, int], None]] = None) -> List[str]:
    """Write ``<out_dir>/<name>.npy`` = float32 [1, n_mel, T] for this rank's share of ``items``.

    decoder: ``tacotron2_subword_b200.Decoder`` in eval mode on a CUDA device (prenet dropout stays on, model.py:23).
    Returns the list of files written by this rank."""
    if decoder.training:
        raise ValueError("GTA extraction runs the model in eval mode (GTA.py:25-28); call .eval() first")
    dev = decoder.gate_layer.linear_layer.weight.device
    two = decoder.n_streams == 2
    mine = shard_by_length([int(it.mel.shape[1]) for it in items], world_size)[rank]
    batches = plan_batches([int(items[i].mel.shape[1]) for i in mine], max_batch, max_frames_per_batch)
    writer = AsyncNpyWriter(out_dir)
    copy_stream = torch.cuda.Stream(device=dev)
    pending = []                                         # (event, host tensor, names, lengths) of batches in flight
    try:
        for bi, batch in enumerate(batches):
            idx = [mine[j] for j in batch]
            B = len(idx)
            T = max(int(items[i].mel.shape[1]) for i in idx)
            T_in = max(int(items[i].memory.shape[0]) for i in idx)
            n_mel, enc = items[idx[0]].mel.shape[0], items[idx[0]].memory.shape[1]
            memory = torch.zeros(B, T_in, enc, device=dev)
            mels = torch.zeros(B, n_mel, T, device=dev)
            mlen = torch.empty(B, dtype=torch.int64)
            tlen = []
            emb = blen = None
            if two:
                T_sub = max(int(items[i].embeddings.shape[0]) for i in idx)
                emb = torch.zeros(B, T_sub, enc, device=dev)
                blen = torch.empty(B, dtype=torch.int64)
            for r, i in enumerate(idx):
                it = items[i]
                memory[r, : it.memory.shape[0]] = it.memory.to(dev, non_blocking=True)
                mels[r, :, : it.mel.shape[1]] = it.mel.to(dev, non_blocking=True)
                mlen[r] = it.memory.shape[0]
                tlen.append(int(it.mel.shape[1]))
                if two:
                    emb[r, : it.embeddings.shape[0]] = it.embeddings.to(dev, non_blocking=True)
                    blen[r] = it.embeddings.shape[0]
            mel_out = decoder(memory, emb, mels, mlen.to(dev), blen.to(dev) if two else None, independent=True)[0]
            # device -> pinned host on a side stream; the next batch's kernels are enqueued meanwhile
            host = torch.empty(mel_out.shape, dtype=torch.float32, pin_memory=True)
            done = torch.cuda.Event()
            copy_stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(copy_stream):
                host.copy_(mel_out, non_blocking=True)
                done.record(copy_stream)
            mel_out.record_stream(copy_stream)
            pending.append((done, host, [items[i].name for i in idx], tlen))
            while len(pending) > 1:                      # hand the previous batch to the writer
                _flush(pending.pop(0), writer, decoder)
            if on_batch is not None:
                on_batch(bi, len(batches))
        while pending:
            _flush(pending.pop(0), writer, decoder)
    finally:
        files = writer.close()
    return files


def _flush(entry, writer: AsyncNpyWriter, decoder=None) -> None:
    done, host, names, tlen = entry
    done.synchronize()
    if decoder is not None:
        # the batch has finished: if an in-kernel watchdog aborted it, its outputs are garbage -- raise instead of writing
        # them (reads the sticky abort word on a private stream, so the next batch keeps running meanwhile)
        decoder.check(sync=False)
    arr = host.numpy()
    for r, (name, t) in enumerate(zip(names, tlen)):
        writer.put(name, np.ascontiguousarray(arr[r: r + 1, :, :t]))     # [1, n_mel, T], GTA.py:61
