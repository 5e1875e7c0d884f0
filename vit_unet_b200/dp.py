"""Batch-sharded data parallelism for the ViT-UNet training step (one process per GPU, NCCL over NVLink).

The reference has no multi-device torch path ("NCCL/MPI/Gloo call sites: none"); the unit
of sharding is the image: every rank holds a full parameter replica, runs the
forward/backward kernels on its slice of the batch, and the only exchange step is ONE averaged all-reduce of the
gradients per step.  BatchNorm statistics of the Re-Attention maps stay per-rank (non-synchronised BN, the DDP
default).

Overlap: the engine writes all gradients into one flat fp32 buffer laid out in forward-execution order, so
backward completes it as a growing SUFFIX.  Each time a block's gradients are final the bucketer is told; once
at least ``bucket_numel`` elements are ready, that contiguous slice is all-reduced asynchronously (the NCCL
kernel runs on the process group's stream while the next block's backward kernels run on the compute stream).
The first bucket is the reconstruction conv + the last skip connection (a 3072^2 projection for Base).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist


class GradBucketer:
    """Host-side bucketing logic; device-agnostic so it is testable with gloo on CPU."""

    def __init__(self, group_starts: Dict[str, int], flat_numel: int, bucket_numel: int = 8 << 20,
                 process_group=None):
        self.group_starts = dict(group_starts)     # prefix -> offset of the group's first element in the flat buffer
        self.flat_numel = flat_numel
        self.bucket_numel = bucket_numel
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._flat: Optional[torch.Tensor] = None
        self._works: List = []
        self._sent_from = flat_numel      # everything in [_sent_from, flat_numel) has been handed to all_reduce
        self.launched: List[tuple] = []   # (start, end) of every bucket, for tests / introspection
        backend = dist.get_backend(process_group) if dist.is_initialized() else ""
        self._avg_native = backend == "nccl"

    def begin(self, flat: torch.Tensor) -> None:
        assert flat.numel() == self.flat_numel
        self._flat, self._works, self._sent_from, self.launched = flat, [], self.flat_numel, []

    def on_ready(self, prefix: str) -> None:
        """All gradients of parameter group `prefix` and of every later group are final."""
        if self.world == 1 or self._flat is None:
            return
        start = self.group_starts.get(prefix)
        if start is None:             # the model has no parameters under this prefix (e.g. no output conv)
            return
        if start < self._sent_from and self._sent_from - start >= self.bucket_numel:
            self._launch(start, self._sent_from)

    def _launch(self, a: int, b: int) -> None:
        chunk = self._flat[a:b]
        if self._avg_native:
            w = dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.pg, async_op=True)
        else:
            w = dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        self._works.append((w, a, b))
        self.launched.append((a, b))
        self._sent_from = a

    def finish(self) -> None:
        """Flush the remaining prefix of the buffer and make the compute stream wait for every bucket."""
        if self.world == 1 or self._flat is None:
            return
        if self._sent_from > 0:
            self._launch(0, self._sent_from)
        for w, a, b in self._works:
            w.wait()
            if not self._avg_native:
                self._flat[a:b].div_(self.world)
        self._works = []


class DataParallel(torch.nn.Module):
    """Wraps a vit_unet_b200 model: same call surface, gradients averaged across ranks during backward."""

    def __init__(self, module, process_group=None, bucket_mb: float = 32.0, broadcast_from: Optional[int] = 0,
                 high_priority: Optional[bool] = None):
        super().__init__()
        self.module = module
        # The all-reduce kernels run next to persistent compute kernels that occupy every SM; on a high-priority stream
        # their CTAs are scheduled as soon as a slot frees instead of queueing behind the compute grid (timeline:
        # tools/dp_timeline.py).  Default: a dedicated NCCL group with high-priority streams when none is given
        # (VU_DP_HIGH_PRIORITY=0 keeps the default group).
        if high_priority is None:
            high_priority = os.environ.get("VU_DP_HIGH_PRIORITY", "1") == "1"
        if (process_group is None and high_priority and dist.is_initialized() and dist.get_backend() == "nccl"
                and dist.get_world_size() > 1):
            try:
                opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
                process_group = dist.new_group(backend="nccl", pg_options=opts)
            except (AttributeError, TypeError, RuntimeError):     # older / newer torch spellings: fall back to the default group
                process_group = None
        pd = dict(module.named_parameters())
        starts: Dict[str, int] = {}
        for name, off in zip(module._param_names, module._flat_offsets):
            prefix = _group_of(name)
            starts.setdefault(prefix, off)
            # finer notification inside a block: the projection weight (D x D: 37.7 MB at Base level 0, the largest tensor of
            # the model) is final right after its wgrad, long before the block's attention-map backward ends, and everything
            # behind it in the flat order (LN, FeedForward of the same block and all later groups) is already final
            if name.endswith("proj.weight"):
                starts.setdefault(name[:-len("weight")], off)
        self.bucketer = GradBucketer(starts, module._flat_numel, int(bucket_mb * (1 << 20) / 4), process_group)
        module._dp = self.bucketer
        if broadcast_from is not None and dist.is_initialized() and dist.get_world_size(process_group) > 1:
            gsrc = dist.get_global_rank(process_group, broadcast_from) if process_group is not None else broadcast_from
            for t in list(pd.values()) + [b for _, b in module.named_buffers()]:
                dist.broadcast(t.data, src=gsrc, group=process_group)

    def forward(self, *a, **k):
        return self.module(*a, **k)

    def sync_buffers(self, src: int = 0) -> None:
        """COLLECTIVE -- every rank of the process group must call it.  Broadcast the buffers (Re-Attention BatchNorm
        running_mean / running_var / num_batches_tracked) of group-rank `src` to every rank.  BatchNorm statistics
        are per-rank during training (standard non-sync-BN data parallelism); call this on all ranks
        before evaluating or checkpointing so they agree on one set of running stats.  `src` is a rank INSIDE the
        process group (0 = its first member), so sub-groups that do not contain global rank 0 work."""
        pg = self.bucketer.pg
        if not (dist.is_initialized() and dist.get_world_size(pg) > 1):
            return
        gsrc = dist.get_global_rank(pg, src) if pg is not None else src
        for _, b in self.module.named_buffers():
            dist.broadcast(b.data, src=gsrc, group=pg)

    def state_dict(self, *a, sync_from: Optional[int] = None, **k):
        """The wrapped model's state_dict (the reference's key layout, no 'module.' prefix).  No collective runs by
        default, so the usual ``if rank == 0: torch.save(model.state_dict())`` cannot hang; call ``sync_buffers()``
        on EVERY rank first (or pass ``sync_from=r`` on every rank) to checkpoint one agreed set of BN statistics."""
        if sync_from is not None:
            self.sync_buffers(sync_from)
        return self.module.state_dict(*a, **k)

    def load_state_dict(self, sd, *a, **k):
        return self.module.load_state_dict(sd, *a, **k)


def _group_of(name: str) -> str:
    """'Encoders.3.ReAttn.proj.weight' -> 'Encoders.3.' ; 'conv2d.weight' -> 'conv2d.' ; 'PE.x' -> 'PE.'"""
    parts = name.split(".")
    if parts[0] in ("Encoders", "BottleNeck", "Decoders", "SkipConnections"):
        return f"{parts[0]}.{parts[1]}."
    return parts[0] + "."
