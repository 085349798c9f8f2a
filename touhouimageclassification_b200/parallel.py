"""Data-parallel training across the GPUs of one NVSwitch box: one process per GPU, replicated parameters,
the global batch split into contiguous per-rank shards, ONE exchange step -- a sum all-reduce of the flat
gradient arena, issued bucket by bucket on a side stream while the backward of earlier layers is still running
(SURVEY.md section 8e). The reference is single-GPU only (``L.Trainer(devices=1)``, ntrain.py:240); its
correctness contract is "N-GPU step on global batch G == 1-GPU step on G".

The loss gradient is scaled by 1 / (per-rank batch * world size) before backward, so the summed gradients are
the global-batch mean and every rank applies the identical AdamW update (parameters stay bit-identical).
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .finetune import fused_train_step
from .model import ViTForImageClassification
from .optim import FusedAdamW


def stage_grad_ranges(model: ViTForImageClassification) -> List[Tuple[int, int]]:
    """Element range of the gradient arena completed by each backward stage (0 = head, 1..L = layers, L+1 = embed)."""
    lib = _lib.load()
    c = model.config.to_c()
    out = []
    for stage in range(model.config.num_hidden_layers + 2):
        b, e = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(lib.tic_vit_stage_grad_range(ctypes.byref(c), ctypes.c_int(stage), ctypes.byref(b), ctypes.byref(e)))
        out.append((b.value, e.value))
    return out


def make_buckets(ranges: List[Tuple[int, int]], target_elems: int) -> List[Tuple[int, int, int, int]]:
    """Group consecutive backward stages into buckets of about ``target_elems`` gradient elements.

    Returns ``(stage_begin, stage_end, elem_begin, elem_end)``; stages 1..L walk the arena downwards, so a run of
    consecutive layer stages is one contiguous slice. Stage 0 (classifier, top of the arena) and the last stage
    (embeddings, bottom of the arena) are merged with their neighbours only when contiguous."""
    buckets = []
    cur = None
    for s, (b, e) in enumerate(ranges):
        if cur is not None and (cur[2] == e or cur[3] == b) and (cur[3] - cur[2]) < target_elems:
            cur = (cur[0], s + 1, min(cur[2], b), max(cur[3], e))
        else:
            if cur is not None:
                buckets.append(cur)
            cur = (s, s + 1, b, e)
    buckets.append(cur)
    return buckets


class GradBucketer:
    """Sum-all-reduces slices of a flat gradient tensor; device-agnostic (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, process_group=None):
        self.group = process_group
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1

    def all_reduce(self, flat: torch.Tensor, begin: int, end: int):
        if self.world_size > 1 and end > begin:
            dist.all_reduce(flat[begin:end], op=dist.ReduceOp.SUM, group=self.group)


class DataParallelTrainer:
    """Fused train step + bucketed gradient all-reduce overlapped with backward."""

    def __init__(self, model: ViTForImageClassification, optimizer: FusedAdamW, process_group=None,
                 bucket_mb: float = 96.0, local: bool = False):
        self.model, self.optimizer = model, optimizer
        self.bucketer = GradBucketer(process_group) if not local else None
        self.world_size = self.bucketer.world_size if self.bucketer is not None else 1
        self.buckets = make_buckets(stage_grad_ranges(model), int(bucket_mb * (1 << 20) / 4))
        # side stream: the all-reduce of a bucket (N > 1) and then AdamW over that bucket run here while the main stream
        # continues with the backward of the earlier layers
        self.comm_stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        self.update_at_end = os.environ.get("TIC_ADAMW_AT_END") is not None  # development A/B: one AdamW launch after backward

    def broadcast_parameters(self, src: int = 0):
        """Make every replica start from rank ``src``'s weights."""
        if self.world_size > 1:
            if not self.model._arena_ok():
                self.model._repack()
            dist.broadcast(self.model._arena, src=src, group=self.bucketer.group)
            self.model.refresh_shadow(force=True)

    def _grad_sync(self, model, dlogits, batch, head_only):
        g = model.grad_arena()
        opt = self.optimizer
        opt.begin_step()
        main = torch.cuda.current_stream() if self.comm_stream is not None else None
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(main)  # the arena was zeroed (and the forward ran) on the main stream
        for (s0, s1, b, e) in self.buckets:
            model.engine_backward(dlogits, batch, head_only=head_only, stage_begin=s0, stage_end=s1)
            if self.comm_stream is None:  # CPU tests (gloo): in order
                if self.bucketer is not None:
                    self.bucketer.all_reduce(g, b, e)
                continue
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                if self.bucketer is not None:
                    self.bucketer.all_reduce(g, b, e)
                # the bucket's gradients are final and its layers' backward is over: nobody reads these weights again
                # in this step, so the update (fp32 parameters, moments, bf16 shadow) can run under the rest of backward
                if not self.update_at_end:
                    opt.step_range(b, e)
        if self.comm_stream is not None:
            main.wait_stream(self.comm_stream)
        if self.comm_stream is None or self.update_at_end:
            opt.step_range(0, g.numel())
        opt.end_step()

    def step(self, inputs=None, target=None, patches=None) -> torch.Tensor:
        return fused_train_step(self.model, self.optimizer, inputs, target, patches=patches,
                                grad_sync=self._grad_sync, world_size=self.world_size)
