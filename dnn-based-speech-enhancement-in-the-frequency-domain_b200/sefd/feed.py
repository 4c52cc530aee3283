"""Rank-sharded on-device data feed: the step either side of the train step.

Replaces the reference's `create_dataloader('train')` (dataloader.py:11-31: whole `.npy` of shape [N, 2, L] in RAM,
DataLoader(batch_size, shuffle=True, drop_last=True, pin_memory=True)) with a feeder that keeps the next batch in
flight: indices -> pinned staging buffer -> asynchronous H2D copy on a side stream, double-buffered, so the copy of
batch i+1 overlaps the train step of batch i.  Utterances are sharded over ranks exactly like
torch.utils.data.DistributedSampler (same permutation for a given seed / epoch, wrap-around padding, rank-strided).
Yields (inputs [B, L], targets [B, L]) CONTIGUOUS float32 CUDA tensors (separate device buffers: the kernels index rows
with stride L), like the reference loop expects after `.to(DEVICE)`; they are views of a ring of three slots, valid until
the iteration after the next one starts (consume them in the loop).  Three slots, so that the host gather of batch i+1
never has to wait for the consumer of the slot it overwrites (batch i-2) while step i is still to be enqueued.
"""
import math

import numpy as np
import torch


def shard_indices(n, epoch=0, seed=0, shuffle=True, rank=0, world=1):
    """Indices of this rank for one epoch: DistributedSampler(dataset, world, rank, shuffle, seed, drop_last=False)."""
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        idx = torch.randperm(n, generator=g).tolist()
    else:
        idx = list(range(n))
    total = math.ceil(n / world) * world
    pad = total - len(idx)
    if pad > 0:
        idx += (idx * math.ceil(pad / len(idx)))[:pad]
    return idx[rank:total:world]


class WaveFeeder:
    SLOTS = 3

    def __init__(self, data, batch, device="cuda", shuffle=True, drop_last=True, seed=0, rank=None, world=None):
        if isinstance(data, str):
            data = np.load(data)                                    # dataloader.py:42
        if data.ndim != 3 or data.shape[1] != 2:
            raise ValueError(f"expected an array [N, 2, L] of (noisy, clean) pairs, got {data.shape}")
        if rank is None or world is None:
            import torch.distributed as dist
            on = dist.is_available() and dist.is_initialized()
            rank, world = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
        self.data, self.batch, self.device = data, int(batch), torch.device(device)
        self.shuffle, self.drop_last, self.seed, self.rank, self.world = shuffle, drop_last, seed, rank, world
        self.epoch = 0
        if self.device.type != "cuda":
            raise RuntimeError("WaveFeeder stages batches for a CUDA device (no CPU path)")
        L = data.shape[2]
        self._pinned = [torch.empty(self.batch, 2, L, dtype=torch.float32).pin_memory() for _ in range(self.SLOTS)]
        self._dev = [(torch.empty(self.batch, L, dtype=torch.float32, device=self.device),
                      torch.empty(self.batch, L, dtype=torch.float32, device=self.device)) for _ in range(self.SLOTS)]
        self._stream = torch.cuda.Stream(device=self.device)
        self._ready = [torch.cuda.Event() for _ in range(self.SLOTS)]      # copy into slot finished
        self._free = [torch.cuda.Event() for _ in range(self.SLOTS)]       # consumer finished with slot

    def set_epoch(self, epoch):
        self.epoch = int(epoch)

    def _batches(self):
        idx = shard_indices(len(self.data), self.epoch, self.seed, self.shuffle, self.rank, self.world)
        nb = len(idx) // self.batch if self.drop_last else math.ceil(len(idx) / self.batch)
        return [idx[i * self.batch:(i + 1) * self.batch] for i in range(nb)]

    def __len__(self):
        return len(self._batches())

    def _stage(self, slot, ids):
        n = len(ids)
        self._free[slot].synchronize()                             # the consumer no longer reads this slot
        buf = self._pinned[slot]
        np.take(self.data, ids, axis=0, out=buf.numpy()[:n])       # gather straight into pinned memory
        with torch.cuda.stream(self._stream):
            self._dev[slot][0][:n].copy_(buf[:n, 0], non_blocking=True)
            self._dev[slot][1][:n].copy_(buf[:n, 1], non_blocking=True)
            self._ready[slot].record(self._stream)
        return n

    def __iter__(self):
        batches = self._batches()
        if not batches:
            return
        cur = torch.cuda.current_stream(self.device)
        for e in self._free:
            e.record(cur)
        n_next = self._stage(0, batches[0])
        for i in range(len(batches)):
            slot, n = i % self.SLOTS, n_next
            if i + 1 < len(batches):
                n_next = self._stage((i + 1) % self.SLOTS, batches[i + 1])     # in flight while batch i is consumed
            cur.wait_event(self._ready[slot])
            noisy, clean = self._dev[slot]
            yield noisy[:n], clean[:n]
            self._free[slot].record(cur)
