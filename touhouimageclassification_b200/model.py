"""Drop-in ``ViTForImageClassification`` for the TouhouIC hot path, backed by the sm_100a engine.

Boundary: the object ``TIC/ViT/model.py:8-47`` returns -- an ``nn.Module`` called as
``m(x).logits`` (``finetune.py:59-60``, ``ntrain.py:47``, ``serve.py:101-106``, ``web/runtime.py:116-117``)
whose ``state_dict()`` has exactly the HuggingFace key layout (200 keys for ViT-B/16,
392 for ViT-L/16, fp32, no buffers), so ``nViT_epoch*.pth`` checkpoints load with ``strict=True``.

The parameters are ordinary fp32 ``nn.Parameter`` objects, but they are *views into one flat arena* whose
layout the native engine dictates (q/k/v adjacent so the QKV projection is a single GEMM). The engine also
keeps a bf16 shadow of the arena for the tensor-core GEMMs. There is no CPU path: calling the module with
CPU tensors raises.
"""
from __future__ import annotations

import ctypes
import threading
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import c_i64, c_int, c_void_p


# ----------------------------------------------------------------------------------------------------
# configuration
# ----------------------------------------------------------------------------------------------------
class TicVitConfigC(ctypes.Structure):
    """Mirror of ``tic_vit_config`` in include/tic_b200.h."""
    _fields_ = [("image_size", ctypes.c_int32), ("patch_size", ctypes.c_int32), ("hidden", ctypes.c_int32),
                ("layers", ctypes.c_int32), ("heads", ctypes.c_int32), ("mlp", ctypes.c_int32),
                ("num_labels", ctypes.c_int32), ("ln_eps", ctypes.c_float)]


@dataclass
class ViTConfig:
    """Subset of ``transformers.ViTConfig`` the hot path depends on (configuration_vit.py:50-65)."""
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    image_size: int = 224
    patch_size: int = 16
    num_channels: int = 3
    num_labels: int = 120
    layer_norm_eps: float = 1e-12
    initializer_range: float = 0.02
    hidden_act: str = "gelu"
    qkv_bias: bool = True

    def to_c(self) -> TicVitConfigC:
        return TicVitConfigC(self.image_size, self.patch_size, self.hidden_size, self.num_hidden_layers,
                             self.num_attention_heads, self.intermediate_size, self.num_labels, self.layer_norm_eps)


_PRESETS = {
    "google/vit-base-patch16-224": dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                                        intermediate_size=3072, image_size=224),
    "google/vit-large-patch16-224": dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                                         intermediate_size=4096, image_size=224),
    "google/vit-large-patch16-384": dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                                         intermediate_size=4096, image_size=384),
}
for _k in list(_PRESETS):
    _PRESETS[_k + "-in21k"] = _PRESETS[_k]


def config_from_name(model_name: str, num_labels: int) -> ViTConfig:
    key = model_name
    if key not in _PRESETS:
        raise ValueError(f"unknown ViT preset {model_name!r}; known: {sorted(_PRESETS)}")
    return ViTConfig(num_labels=num_labels, **_PRESETS[key])


class ImageClassifierOutput:
    """What callers read: ``.logits`` (and ``.loss`` when labels were passed), like the HF output class."""

    def __init__(self, logits, loss=None):
        self.logits = logits
        self.loss = loss
        self.hidden_states = None
        self.attentions = None

    def __getitem__(self, i):
        items = ([self.loss] if self.loss is not None else []) + [self.logits]
        return items[i]

    def __iter__(self):
        return iter(([self.loss] if self.loss is not None else []) + [self.logits])


# ----------------------------------------------------------------------------------------------------
# module skeleton with the HuggingFace parameter names
# ----------------------------------------------------------------------------------------------------
class _Bag(nn.Module):
    """A parameter container (the arithmetic lives in the engine, not in ``forward``)."""

    def forward(self, *a, **k):  # pragma: no cover - never called
        raise RuntimeError("submodules of the B200 ViT are parameter containers; call the top-level model")


class _Linear(_Bag):
    def __init__(self, w, b):
        super().__init__()
        self.weight, self.bias = w, b
        self.out_features, self.in_features = w.shape[0], w.shape[1]


class _LayerNorm(_Bag):
    def __init__(self, w, b, eps):
        super().__init__()
        self.weight, self.bias, self.eps = w, b, eps


class _Conv(_Bag):
    def __init__(self, w, b):
        super().__init__()
        self.weight, self.bias = w, b


def _shapes(cfg: ViTConfig):
    """(name, shape) in HF ``named_parameters()`` order == tic_vit_param_layout order."""
    D, F, C = cfg.hidden_size, cfg.intermediate_size, cfg.num_labels
    N = (cfg.image_size // cfg.patch_size) ** 2 + 1
    out = [("vit.embeddings.cls_token", (1, 1, D)), ("vit.embeddings.position_embeddings", (1, N, D)),
           ("vit.embeddings.patch_embeddings.projection.weight", (D, 3, 16, 16)),
           ("vit.embeddings.patch_embeddings.projection.bias", (D,))]
    for i in range(cfg.num_hidden_layers):
        p = f"vit.encoder.layer.{i}."
        out += [(p + "attention.attention.query.weight", (D, D)), (p + "attention.attention.query.bias", (D,)),
                (p + "attention.attention.key.weight", (D, D)), (p + "attention.attention.key.bias", (D,)),
                (p + "attention.attention.value.weight", (D, D)), (p + "attention.attention.value.bias", (D,)),
                (p + "attention.output.dense.weight", (D, D)), (p + "attention.output.dense.bias", (D,)),
                (p + "intermediate.dense.weight", (F, D)), (p + "intermediate.dense.bias", (F,)),
                (p + "output.dense.weight", (D, F)), (p + "output.dense.bias", (D,)),
                (p + "layernorm_before.weight", (D,)), (p + "layernorm_before.bias", (D,)),
                (p + "layernorm_after.weight", (D,)), (p + "layernorm_after.bias", (D,))]
    out += [("vit.layernorm.weight", (D,)), ("vit.layernorm.bias", (D,)),
            ("classifier.weight", (C, D)), ("classifier.bias", (C,))]
    return out


def param_layout(cfg: ViTConfig):
    """Arena element offsets / sizes from the native engine (single source of truth)."""
    lib = _lib.load()
    lib.tic_vit_param_arena_elems.restype = ctypes.c_int64
    lib.tic_vit_head_offset.restype = ctypes.c_int64
    c = cfg.to_c()
    total = lib.tic_vit_param_arena_elems(ctypes.byref(c))
    if total < 0:
        _lib.check(1)
    count = 4 + 16 * cfg.num_hidden_layers + 4
    offs = (ctypes.c_int64 * count)()
    nums = (ctypes.c_int64 * count)()
    n = lib.tic_vit_param_layout(ctypes.byref(c), offs, nums, c_int(count))
    if n != count:
        _lib.check(1)
    head = lib.tic_vit_head_offset(ctypes.byref(c))
    return int(total), list(offs), list(nums), int(head)


def _on_model_device(fn):
    """Run an engine call with the model's device current. Kernel launches, tensor-map encodes and the stream handed to
    the C-ABI belong to the CURRENT device; a replica on ``cuda:1``, or a worker thread (whose current device defaults
    to 0), would otherwise launch on the wrong GPU with this model's pointers."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        dev = self._arena.device
        if dev.type != "cuda" or dev.index == torch.cuda.current_device():
            return fn(self, *args, **kwargs)
        with torch.cuda.device(dev):
            return fn(self, *args, **kwargs)
    return wrapper


class ViTForImageClassification(nn.Module):
    """B200-native stand-in for ``transformers.ViTForImageClassification`` (modeling_vit.py:605-653)."""

    base_model_prefix = "vit"

    def __init__(self, config: ViTConfig):
        super().__init__()
        if config.num_channels != 3 or config.patch_size != 16:
            raise ValueError("the B200 ViT engine supports 3-channel, 16x16-patch models only")
        self.config = config
        self.num_labels = config.num_labels
        self._total, self._offsets, self._numels, self._head_offset = param_layout(config)
        self._names = [n for n, _ in _shapes(config)]
        self._shapes = [s for _, s in _shapes(config)]
        for (n, s), ne in zip(_shapes(config), self._numels):
            assert int(torch.Size(s).numel()) == ne, (n, s, ne)
        arena = torch.zeros(self._total, dtype=torch.float32)
        params = [nn.Parameter(arena[o:o + n].view(s)) for o, n, s in zip(self._offsets, self._numels, self._shapes)]
        self._arena = arena
        self._build_tree(params)
        self._init_weights()
        # engine state (created lazily on the parameters' device)
        self._shadow = None           # bf16 copy of the arena
        self._shadow_key = None       # version stamp the shadow corresponds to
        self._grad_arena = None
        self._workspaces = {}
        self._ws_generation = 0
        self._lock = threading.RLock()
        # "bf16": tensor-core engine with autocast rounding points (training, ntrain.py:241 bf16-mixed).
        # "fp32": split-bf16 GEMMs, fp32 everything else -- the reference's no-autocast inference (serve.py:99-101);
        #         forward only, selected with set_precision("fp32").
        self.precision = "bf16"
        self._w6 = None
        self._w6_key = None
        # CUDA graphs of the inference forward for small batches (the web / serve path is launch-bound at batch 1-64).
        # Batches are padded up to the next bucket, so at most len(graph_buckets) graphs + workspaces are ever retained
        # (a dynamic batcher produces every size from 1 to 64).
        self.graph_max_batch = 64
        self.graph_buckets = (1, 2, 4, 8, 16, 32, 64)
        self._graphs = {}
        self._capture_stream = None

    # ---- structure ------------------------------------------------------------------------------
    def _build_tree(self, params):
        cfg = self.config
        it = iter(params)
        vit = _Bag()
        emb = _Bag()
        emb.cls_token = next(it)
        emb.position_embeddings = next(it)
        pe = _Bag()
        pe.projection = _Conv(next(it), next(it))
        emb.patch_embeddings = pe
        vit.embeddings = emb
        layers = []
        for _ in range(cfg.num_hidden_layers):
            layer = _Bag()
            att = _Bag()
            inner = _Bag()
            inner.query = _Linear(next(it), next(it))
            inner.key = _Linear(next(it), next(it))
            inner.value = _Linear(next(it), next(it))
            att.attention = inner
            ao = _Bag()
            ao.dense = _Linear(next(it), next(it))
            att.output = ao
            layer.attention = att
            inter = _Bag()
            inter.dense = _Linear(next(it), next(it))
            layer.intermediate = inter
            outp = _Bag()
            outp.dense = _Linear(next(it), next(it))
            layer.output = outp
            layer.layernorm_before = _LayerNorm(next(it), next(it), cfg.layer_norm_eps)
            layer.layernorm_after = _LayerNorm(next(it), next(it), cfg.layer_norm_eps)
            layers.append(layer)
        enc = _Bag()
        enc.layer = nn.ModuleList(layers)
        vit.encoder = enc
        vit.layernorm = _LayerNorm(next(it), next(it), cfg.layer_norm_eps)
        self.vit = vit
        self.classifier = _Linear(next(it), next(it))

    @property
    def base_model(self):
        """Everything except the classifier (ntrain.py:35-37 freezes ``self.vit.base_model.parameters()``)."""
        return self.vit

    def _init_weights(self):
        """Same distributions as modeling_vit.py:384-398 (trunc-normal 0.02 / zeros / ones)."""
        std = self.config.initializer_range
        with torch.no_grad():
            for name, p in self.named_parameters():
                if name.endswith("layernorm.weight") or "layernorm_before.weight" in name or "layernorm_after.weight" in name:
                    p.fill_(1.0)
                elif name.endswith(".bias"):
                    p.zero_()
                else:
                    nn.init.trunc_normal_(p, mean=0.0, std=std)

    # ---- arena maintenance ------------------------------------------------------------------------
    def _params_in_order(self):
        return [p for _, p in self.named_parameters()]

    def _repack(self, device=None):
        """Re-point every parameter at a view of one flat arena (after .to()/.cuda() moved them apart)."""
        params = self._params_in_order()
        device = params[0].device if device is None else device
        arena = torch.zeros(self._total, dtype=torch.float32, device=device)
        with torch.no_grad():
            for p, o, n, s in zip(params, self._offsets, self._numels, self._shapes):
                view = arena[o:o + n].view(s)
                view.copy_(p.detach().to(device=device, dtype=torch.float32))
                p.data = view
                p.grad = None
        self._arena = arena
        self._shadow = None
        self._shadow_key = None
        self._w6 = None
        self._w6_key = None
        self._grad_arena = None
        self._workspaces = {}
        self._graphs = {}

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        p0 = next(self.parameters())
        if p0.dtype != torch.float32:
            raise TypeError("parameters of the B200 ViT stay fp32 (the engine keeps its own bf16 shadow)")
        self._repack(p0.device)
        return self

    def _arena_ok(self) -> bool:
        a = self._arena
        base = a.data_ptr()
        for p, o in zip(self._params_in_order(), self._offsets):
            if p.device != a.device or p.data_ptr() != base + 4 * o:
                return False
        return True

    def _version_key(self):
        return sum(p._version for p in self._params_in_order())

    @_on_model_device
    def refresh_shadow(self, force: bool = False):
        """Bring the bf16 shadow up to date with the fp32 parameters (one cast kernel over the arena)."""
        if not self._arena_ok():
            self._repack()
        key = self._version_key()
        if self._shadow is None:
            self._shadow = torch.empty(self._total, dtype=torch.bfloat16, device=self._arena.device)
            force = True
        if force or key != self._shadow_key:
            _lib.check(_lib.load().tic_cast_f32_to_bf16(c_void_p(self._arena.data_ptr()), c_void_p(self._shadow.data_ptr()),
                                                        c_i64(self._total), _stream(self._arena.device)))
            self._shadow_key = key

    def set_precision(self, precision: str):
        """``"bf16"`` (default) or ``"fp32"`` (inference only; logits within 1e-4 relative of the fp32 reference)."""
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        self.precision = precision
        if precision == "bf16":
            self._w6 = None
            self._w6_key = None
        return self

    @_on_model_device
    def _refresh_w6(self):
        """Split weights for the fp32 mode: every GEMM weight as three bf16 terms laid out [N, 6K]."""
        if not self._arena_ok():
            self._repack()
        lib = _lib.load()
        key = self._version_key()
        c = self.config.to_c()
        if self._w6 is None or self._w6.device != self._arena.device:
            lib.tic_vit_w6_elems.restype = ctypes.c_int64
            n = lib.tic_vit_w6_elems(ctypes.byref(c))
            if n < 0:
                _lib.check(1)
            self._w6 = torch.empty(n, dtype=torch.bfloat16, device=self._arena.device)
            self._w6_key = None
        if key != self._w6_key:
            _lib.check(lib.tic_vit_prepare_w6(ctypes.byref(c), c_void_p(self._arena.data_ptr()),
                                              c_void_p(self._w6.data_ptr()), _stream(self._arena.device)))
            self._w6_key = key

    @_on_model_device
    def engine_forward_f32(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """fp32-mode forward (tic_vit_forward_f32): fp32 NCHW pixels -> fp32 logits."""
        with self._lock:
            self._check_input(pixel_values)
            self._refresh_w6()
            x = pixel_values.detach()
            if x.dtype != torch.float32 or not x.is_contiguous():
                x = x.float().contiguous()
            batch = x.shape[0]
            lib = _lib.load()
            c = self.config.to_c()
            key = (batch, "f32")
            ws = self._workspaces.get(key)
            if ws is None or ws.device != self._arena.device:
                lib.tic_vit_workspace_bytes_f32.restype = ctypes.c_int64
                nbytes = lib.tic_vit_workspace_bytes_f32(ctypes.byref(c), c_int(batch))
                if nbytes < 0:
                    _lib.check(1)
                for k in [k for k in self._workspaces if k[1] == "f32"]:
                    del self._workspaces[k]
                ws = torch.empty(nbytes, dtype=torch.uint8, device=self._arena.device)
                self._workspaces[key] = ws
            logits = torch.empty((batch, self.config.num_labels), dtype=torch.float32, device=self._arena.device)
            _lib.check(lib.tic_vit_forward_f32(
                ctypes.byref(c), c_void_p(self._arena.data_ptr()), c_void_p(self._w6.data_ptr()), c_void_p(x.data_ptr()),
                c_int(batch), c_void_p(ws.data_ptr()), c_i64(ws.numel()), c_void_p(logits.data_ptr()), _stream(self._arena.device)))
            return logits

    def mark_shadow_fresh(self):
        """Called by the fused optimizer, which writes the shadow itself."""
        self._shadow_key = self._version_key()

    def grad_arena(self) -> torch.Tensor:
        if self._grad_arena is None or self._grad_arena.device != self._arena.device:
            self._grad_arena = torch.zeros(self._total, dtype=torch.float32, device=self._arena.device)
        return self._grad_arena

    def grad_views(self):
        g = self.grad_arena()
        return [g[o:o + n].view(s) for o, n, s in zip(self._offsets, self._numels, self._shapes)]

    def _workspace(self, batch: int, training: bool) -> torch.Tensor:
        key = (batch, bool(training))
        ws = self._workspaces.get(key)
        if ws is None or ws.device != self._arena.device:
            lib = _lib.load()
            lib.tic_vit_workspace_bytes.restype = ctypes.c_int64
            c = self.config.to_c()
            nbytes = lib.tic_vit_workspace_bytes(ctypes.byref(c), c_int(batch), c_int(int(training)))
            if nbytes < 0:
                _lib.check(1)
            if training:  # keep a single training workspace alive (tens of GB at batch 256)
                for k in [k for k in self._workspaces if k[1] is True]:
                    del self._workspaces[k]
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self._arena.device)
            self._workspaces[key] = ws
        return ws

    # ---- engine calls -----------------------------------------------------------------------------
    def _check_input(self, pixel_values: torch.Tensor):
        if not isinstance(pixel_values, torch.Tensor) or pixel_values.dim() != 4:
            raise ValueError("pixel_values must be a [batch, channels, height, width] tensor")
        b, ch, h, w = pixel_values.shape
        if ch != self.config.num_channels:
            raise ValueError("Make sure that the channel dimension of the pixel values match with the one set in the "
                             f"configuration. Expected {self.config.num_channels} but got {ch}.")
        s = self.config.image_size
        if h != s or w != s:
            raise ValueError(f"Input image size ({h}*{w}) doesn't match model ({s}*{s}).")
        if not pixel_values.is_cuda or not self._arena.is_cuda:
            raise RuntimeError("the B200 ViT runs on CUDA only: move the model and pixel_values to the GPU "
                               "(there is no CPU fallback)")

    @_on_model_device
    def engine_forward(self, pixel_values=None, patches=None, training=False) -> torch.Tensor:
        """Raw engine forward: fp32 NCHW pixels or bf16 patch rows -> fp32 logits [B, num_labels]."""
        with self._lock:
            self.refresh_shadow()
            if pixel_values is not None:
                self._check_input(pixel_values)
                x = pixel_values.detach()
                if x.dtype != torch.float32 or not x.is_contiguous():
                    x = x.float().contiguous()
                batch = x.shape[0]
            else:
                P = (self.config.image_size // 16) ** 2
                assert patches.dtype == torch.bfloat16 and patches.is_cuda and patches.is_contiguous()
                batch = patches.shape[0] // P
                x = None
            if not training and batch <= self.graph_max_batch and not torch.cuda.is_current_stream_capturing():
                return self._graph_forward(x if x is not None else patches, batch, from_patches=x is None)
            ws = self._workspace(batch, training)
            if training:
                self._ws_generation += 1
            logits = torch.empty((batch, self.config.num_labels), dtype=torch.float32, device=self._arena.device)
            c = self.config.to_c()
            _lib.check(_lib.load().tic_vit_forward(
                ctypes.byref(c), c_void_p(self._arena.data_ptr()), c_void_p(self._shadow.data_ptr()),
                c_void_p(0 if x is None else x.data_ptr()), c_void_p(0 if patches is None else patches.data_ptr()),
                c_int(batch), c_void_p(ws.data_ptr()), c_i64(ws.numel()), c_int(int(training)),
                c_void_p(logits.data_ptr()), _stream(self._arena.device)))
            return logits

    def _graph_forward(self, x: torch.Tensor, batch: int, from_patches: bool = False) -> torch.Tensor:
        """Inference forward of a small batch as ONE CUDA-graph launch: the ~220 kernel launches (and their tensor-map
        encodes) of a ViT-L forward cost more host time than device time below batch 64. The graph is captured once per
        batch size over static input / logits / workspace buffers; weights are read through the (stable) shadow arena,
        which refresh_shadow() keeps current outside the graph."""
        bucket = next((b for b in self.graph_buckets if b >= batch), batch)
        key = (bucket, from_patches)
        entry = self._graphs.get(key)
        if entry is None or entry["shadow_ptr"] != self._shadow.data_ptr() or entry["arena_ptr"] != self._arena.data_ptr():
            dev = self._arena.device
            ws = self._workspace(bucket, False)
            rows = x.shape[0] // batch if from_patches else 1  # patch rows per image (bf16 [B * P, 768] input)
            xin = torch.zeros((bucket * rows,) + tuple(x.shape[1:]), dtype=x.dtype, device=dev)
            out = torch.empty((bucket, self.config.num_labels), dtype=torch.float32, device=dev)
            c = self.config.to_c()

            def launch():
                _lib.check(_lib.load().tic_vit_forward(
                    ctypes.byref(c), c_void_p(self._arena.data_ptr()), c_void_p(self._shadow.data_ptr()),
                    c_void_p(0 if from_patches else xin.data_ptr()), c_void_p(xin.data_ptr() if from_patches else 0),
                    c_int(bucket), c_void_p(ws.data_ptr()), c_i64(ws.numel()),
                    c_int(0), c_void_p(out.data_ptr()), _stream(dev)))

            # an explicit capture stream ON THIS DEVICE: torch.cuda.graph's default one is a class-level singleton created
            # on whichever device captured first, so a replica on another GPU would capture an empty graph
            if self._capture_stream is None or self._capture_stream.device != dev:
                self._capture_stream = torch.cuda.Stream(device=dev)
            # eager warm-up ON THE CAPTURE STREAM: one-time function attributes, driver entry points and anything else the
            # launchers set up per device / stream on first use happen outside the capture
            cur = torch.cuda.current_stream(dev)
            self._capture_stream.wait_stream(cur)
            with torch.cuda.stream(self._capture_stream):
                launch()
            self._capture_stream.synchronize()
            graph = torch.cuda.CUDAGraph()
            # thread_local: another thread of the process (a DataLoader pin-memory thread, a second replica) may call
            # cudaHostAlloc / cudaMalloc while this one captures; the default 'global' mode would fail the capture
            with torch.cuda.graph(graph, stream=self._capture_stream, capture_error_mode="thread_local"):
                launch()
            entry = dict(graph=graph, x=xin, out=out, ws=ws, shadow_ptr=self._shadow.data_ptr(),
                         arena_ptr=self._arena.data_ptr())
            self._graphs[key] = entry
        entry["x"][:x.shape[0]].copy_(x)
        entry["graph"].replay()
        return entry["out"][:batch].clone()

    @_on_model_device
    def engine_backward(self, dlogits: torch.Tensor, batch: int, head_only: bool = False, stage_begin: int = 0,
                        stage_end: Optional[int] = None):
        """Accumulate parameter gradients of the last training forward into ``grad_arena()``."""
        with self._lock:
            ws = self._workspaces.get((batch, True))
            if ws is None:
                raise RuntimeError("engine_backward called without a preceding training forward of the same batch size")
            g = self.grad_arena()
            if stage_end is None:
                stage_end = self.config.num_hidden_layers + 2
            c = self.config.to_c()
            dl = dlogits.detach()
            if dl.dtype != torch.float32 or not dl.is_contiguous():
                dl = dl.float().contiguous()
            _lib.check(_lib.load().tic_vit_backward(
                ctypes.byref(c), c_void_p(self._arena.data_ptr()), c_void_p(self._shadow.data_ptr()), c_int(batch),
                c_void_p(ws.data_ptr()), c_i64(ws.numel()), c_void_p(dl.data_ptr()), c_void_p(g.data_ptr()),
                c_int(stage_begin), c_int(stage_end), c_int(int(head_only)), _stream(self._arena.device)))

    # ---- nn.Module surface ------------------------------------------------------------------------
    def forward(self, pixel_values=None, labels=None, interpolate_pos_encoding=None, **kwargs):
        if pixel_values is None:
            raise ValueError("You have to specify pixel_values")
        if interpolate_pos_encoding:
            raise NotImplementedError("interpolate_pos_encoding is not used by the reference (SURVEY section 5) and "
                                      "is not implemented; build the model with the target image_size instead")
        self._check_input(pixel_values)
        params = self._params_in_order()
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if self.precision == "fp32":
            if needs_grad:
                raise RuntimeError("precision='fp32' is the inference mode (serve.py:99-101); the reference never "
                                   "trains in fp32 -- use torch.no_grad() / requires_grad_(False), or set_precision('bf16')")
            logits = self.engine_forward_f32(pixel_values)
        elif needs_grad:
            logits = _ViTFunction.apply(self, pixel_values, *params)
        else:
            logits = self.engine_forward(pixel_values, training=False)
        if torch.is_autocast_enabled():
            logits = logits.to(torch.get_autocast_dtype("cuda"))
        loss = None
        if labels is not None:
            from . import ops
            loss = ops.cross_entropy(logits.view(-1, self.num_labels), labels.view(-1))
        return ImageClassifierOutput(logits=logits, loss=loss)


def _stream(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _ViTFunction(torch.autograd.Function):
    """Generic autograd bridge: lets ``loss.backward()`` + any torch optimizer drive the engine."""

    @staticmethod
    def forward(ctx, model, pixel_values, *params):
        logits = model.engine_forward(pixel_values, training=True)
        ctx.model = model
        ctx.batch = pixel_values.shape[0]
        ctx.generation = model._ws_generation
        ctx.needs = [p.requires_grad for p in params]
        ctx.params = params
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model = ctx.model
        if ctx.generation != model._ws_generation:
            raise RuntimeError("the B200 ViT keeps the activations of ONE training forward: call backward() before "
                               "the next forward in training mode")
        views = model.grad_views()
        g = model.grad_arena()
        trainable = [i for i, n in enumerate(ctx.needs) if n]
        aliased = [i for i in trainable if ctx.params[i].grad is not None
                   and ctx.params[i].grad.data_ptr() == views[i].data_ptr()]
        accumulate_in_place = len(aliased) == len(trainable) and len(trainable) > 0
        if not accumulate_in_place:
            for i in aliased:  # rare mixed state: detach the old gradient from the arena before it is reused
                ctx.params[i].grad = ctx.params[i].grad.clone()
            g.zero_()
        head_only = not any(ctx.needs[i] for i in range(len(ctx.needs) - 2))
        model.engine_backward(dlogits, ctx.batch, head_only=head_only)
        if accumulate_in_place:
            grads = [None] * len(ctx.needs)  # p.grad already aliases the arena the engine accumulated into
        else:
            grads = [views[i] if ctx.needs[i] else None for i in range(len(ctx.needs))]
        return (None, None, *grads)


# ----------------------------------------------------------------------------------------------------
# factory mirroring TIC/ViT/model.py:8-47
# ----------------------------------------------------------------------------------------------------
def ViT(num_classes: int, pretrained: bool = True, model_name: str = None, wrap_model_name=True,
        image_size: Optional[int] = None) -> ViTForImageClassification:
    """Same signature and defaults as the reference factory (default name ``google/vit-large-patch16-224-in21k``).

    ``pretrained=True`` needs the HuggingFace checkpoint on local disk (the reference downloads it,
    ``TIC/utils/ensure.py:11-15``); it is loaded through ``transformers`` and copied into the arena.
    """
    if model_name is None:
        model_name = "google/vit-large-patch16-224-in21k"
    cfg = config_from_name(model_name, num_classes)
    if image_size is not None:
        cfg.image_size = image_size
    model = ViTForImageClassification(cfg)
    if pretrained:
        try:
            from transformers import ViTForImageClassification as HFViT
            hf = HFViT.from_pretrained(model_name, num_labels=num_classes, ignore_mismatched_sizes=True)
        except Exception as e:  # no network / no local snapshot
            raise OSError(f"pretrained=True needs a local snapshot of {model_name!r} ({e})") from e
        if hf.config.image_size != 224:
            raise ValueError(f"Pretrained model's image size {hf.config.image_size} does not match "
                             f"the specified image size 224.")
        model.load_state_dict(hf.state_dict(), strict=True)
    return model
