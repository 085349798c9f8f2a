"""Fused AdamW over the model's flat parameter arena (one kernel launch per contiguous trainable range).

Stands where ``torch.optim.AdamW(self.parameters(), lr, weight_decay)`` stands in the reference
(``ntrain.py:39-41`` [a16], ``finetune.py:314``): one parameter group, decoupled weight decay applied to
every tensor (biases and LayerNorm included), betas (0.9, 0.999), eps 1e-8. The same launch refreshes the
bf16 shadow weights the tensor-core GEMMs read, so no separate cast pass is needed after the step.

It subclasses ``torch.optim.Optimizer`` so LR schedulers (``get_linear_schedule_with_warmup``,
``finetune.py:324``) and ``state_dict()`` / ``load_state_dict()`` keep working; the per-parameter state
tensors (``exp_avg``, ``exp_avg_sq``) are views into two flat arenas, in ``named_parameters()`` order --
the index order torch's own AdamW state_dict uses, so tuple checkpoints written by ``finetune.py:249-258``
can be resumed.
"""
from __future__ import annotations

import torch

from . import ops
from .model import ViTForImageClassification


def find_engine_model(obj) -> ViTForImageClassification:
    if isinstance(obj, ViTForImageClassification):
        return obj
    if isinstance(obj, torch.nn.Module):
        for m in obj.modules():
            if isinstance(m, ViTForImageClassification):
                return m
    raise TypeError("FusedAdamW needs the B200 ViTForImageClassification (or a module containing it)")


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.model = find_engine_model(model)
        if not self.model._arena_ok():
            self.model._repack()
        params = [p for p in self.model._params_in_order() if p.requires_grad]
        if not params:
            raise ValueError("FusedAdamW: the model has no trainable parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._step = 0
        self._m = None
        self._v = None
        self._ranges = None
        self._arena_ptr = None
        self.grads_in_arena = False  # set by the fused train step: gradients already live in the grad arena
        self.arena_clean = False     # True right after zero_grad(): the fused step need not zero the arena again

    # ---- arenas -----------------------------------------------------------------------------------
    def _ensure_state(self):
        model = self.model
        if not model._arena_ok():
            model._repack()
        arena = model._arena
        if self._m is not None and self._arena_ptr == arena.data_ptr():
            return
        old_m, old_v = self._m, self._v
        self._m = torch.zeros_like(arena)
        self._v = torch.zeros_like(arena)
        if old_m is not None and old_m.numel() == arena.numel():  # model moved devices: carry the moments over
            self._m.copy_(old_m)
            self._v.copy_(old_v)
        self._arena_ptr = arena.data_ptr()
        trainable = set(id(p) for g in self.param_groups for p in g["params"])
        spans = []
        for p, o, n in zip(model._params_in_order(), model._offsets, model._numels):
            if id(p) not in trainable:
                continue
            spans.append((o, o + (n + 63) // 64 * 64))
            st = self.state[p]
            st["step"] = torch.tensor(float(self._step))
            st["exp_avg"] = self._m[o:o + n].view(p.shape)
            st["exp_avg_sq"] = self._v[o:o + n].view(p.shape)
        ranges = []
        for a, b in sorted(spans):  # arena order differs from named_parameters() order (q/k/v are interleaved)
            if ranges and ranges[-1][1] == a:
                ranges[-1][1] = b
            else:
                ranges.append([a, b])
        self._ranges = [(a, min(b, arena.numel())) for a, b in ranges]

    # ---- torch.optim surface ------------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = True):
        for g in self.param_groups:
            for p in g["params"]:
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()
        self.model.grad_arena().zero_()
        self.grads_in_arena = False
        self.arena_clean = True

    def _gather_grads(self):
        """Generic path (``loss.backward()``): make sure every .grad is inside the gradient arena."""
        views = self.model.grad_views()
        index = {id(p): i for i, p in enumerate(self.model._params_in_order())}
        for g in self.param_groups:
            for p in g["params"]:
                v = views[index[id(p)]]
                if p.grad is None:
                    v.zero_()
                elif p.grad.data_ptr() != v.data_ptr():
                    v.copy_(p.grad)

    # The update can be applied slice by slice (the arena is flat and AdamW is elementwise): the train step runs it for
    # each gradient bucket on a side stream as soon as that bucket's gradients are final, in the shadow of the backward
    # of the earlier layers. begin_step / step_range / end_step are the three parts of step().
    @torch.no_grad()
    def begin_step(self):
        self._ensure_state()
        self._step += 1
        if self.model._shadow is None:
            self.model.refresh_shadow(force=True)

    @torch.no_grad()
    def step_range(self, begin: int, end: int, grad_scale: float = 1.0):
        """AdamW over the trainable part of arena elements [begin, end) on the current stream."""
        model = self.model
        group = self.param_groups[0]
        g = model.grad_arena()
        b1, b2 = group["betas"]
        for ra, rb in self._ranges:
            a, b = max(begin, ra), min(end, rb)
            if a < b:
                ops.adamw_step(model._arena[a:b], g[a:b], self._m[a:b], self._v[a:b], model._shadow[a:b],
                               float(group["lr"]), b1, b2, group["eps"], group["weight_decay"], self._step, grad_scale)

    @torch.no_grad()
    def end_step(self):
        for st in self.state.values():
            if "step" in st:
                st["step"].fill_(float(self._step))
        self.model.mark_shadow_fresh()
        self.grads_in_arena = False
        self.arena_clean = False

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._ensure_state()
        if not self.grads_in_arena:
            self._gather_grads()
        self.begin_step()
        self.step_range(0, self.model._arena.numel(), grad_scale)
        self.end_step()
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        loaded = {id(p): dict(self.state[p]) for g in self.param_groups for p in g["params"] if p in self.state}
        self._m = None
        self._ensure_state()
        steps = []
        for g in self.param_groups:
            for p in g["params"]:
                st = loaded.get(id(p))
                if st and "exp_avg" in st:
                    self.state[p]["exp_avg"].copy_(st["exp_avg"])
                    self.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
                    steps.append(int(float(st.get("step", 0))))
        self._step = max(steps) if steps else 0
        for st in self.state.values():
            st["step"] = torch.tensor(float(self._step))
