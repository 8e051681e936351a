"""Multi-GPU execution: one process per GPU, the cutout's TIME axis sharded.

Every (time step, cell) is independent and the reduction runs over space
only, so rank r owns the contiguous steps ``shard_bounds(nt, world, r)`` of
every field, the CSR plan and the small tables are replicated, and the only
communication is re-assembling the ``(time, bus)`` result (one all-gather of
a few MB over NCCL/NVLink) or, for time-aggregated per-cell outputs, one
all-reduce of a ``(y, x)`` plane.  Heat demand shards must be cut on day
boundaries (``align=24`` steps for hourly data).

Works with any initialised ``torch.distributed`` backend (``nccl`` on GPUs,
``gloo`` in the CPU tests).
"""

from __future__ import annotations

import numpy as np
import pandas as pd


def shard_bounds(nt, world, rank, align=1):
    """Contiguous [begin, end) of `rank`: blocks of `align` steps are dealt as
    evenly as possible; earlier ranks take the remainder."""
    nblk = -(-nt // align)
    base, rem = divmod(nblk, world)
    b0 = rank * base + min(rank, rem)
    b1 = b0 + base + (1 if rank < rem else 0)
    return min(b0 * align, nt), min(b1 * align, nt)


class TimeShard:
    """Attached to a ``Cutout`` holding one rank's time shard."""

    def __init__(self, group=None):
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def _device_for_comm(self, t):
        import torch
        import torch.distributed as dist

        if dist.get_backend(self.group) == "nccl":
            return t if t.is_cuda else t.cuda()
        return t.cpu()

    def gather_time(self, local, labels=None, counts=None, async_op=False):
        """Concatenate per-rank (nt_r, ...) results along time, in rank order.
        ``labels=None`` skips the (host-side, pickled) gather of the time labels;
        ``counts`` (steps per rank, when the caller knows the shard layout) skips
        the size exchange and its host synchronisation.  ``async_op`` (equal shards,
        no labels) returns ``(out, work)`` without making the current stream wait, so
        the next conversion overlaps the NVLink transfer; ``work.wait()`` before ``out``
        is read."""
        import torch
        import torch.distributed as dist

        t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
        t = self._device_for_comm(t.contiguous())
        if counts is None:
            n_local = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
            cl = [torch.zeros_like(n_local) for _ in range(self.world)]
            dist.all_gather(cl, n_local, group=self.group)
            counts = [int(c.item()) for c in cl]
        assert len(counts) == self.world and counts[self.rank] == t.shape[0]
        if len(set(counts)) == 1:
            out = torch.empty((self.world * counts[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            if async_op and labels is None:
                return out, dist.all_gather_into_tensor(out, t, group=self.group, async_op=True)
            dist.all_gather_into_tensor(out, t, group=self.group)
        else:  # ragged shards: pad to the longest
            m = max(counts)
            pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            pad[: t.shape[0]] = t
            parts = [torch.empty_like(pad) for _ in range(self.world)]
            dist.all_gather(parts, pad, group=self.group)
            out = torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
        if labels is None:
            return out, None
        all_labels = [None] * self.world
        dist.all_gather_object(all_labels, np.asarray(pd.Index(labels).values), group=self.group)
        lab = np.concatenate(all_labels)
        if np.issubdtype(lab.dtype, np.datetime64):
            lab = pd.DatetimeIndex(lab)
        return out, lab

    def sum_planes(self, *planes):
        """All-reduce (sum) per-cell planes, e.g. the NaN-skipping time sum and the number
        of valid steps per cell of a time-aggregated per-cell output (one collective)."""
        import torch
        import torch.distributed as dist

        ts = [p if isinstance(p, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(p)) for p in planes]
        t = self._device_for_comm(torch.stack([x.to(torch.float32) for x in ts]).contiguous())
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return tuple(t[i] for i in range(len(ts)))

    def sum_over_ranks(self, local, n_t):
        """All-reduce a per-cell time sum and the number of steps it covers."""
        import torch
        import torch.distributed as dist

        t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
        t = self._device_for_comm(t.contiguous().clone())
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        n = torch.tensor([n_t], dtype=torch.int64, device=t.device)
        dist.all_reduce(n, op=dist.ReduceOp.SUM, group=self.group)
        return t, int(n.item())
