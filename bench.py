#!/usr/bin/env python
"""Benchmark of the hot path: ViT-L/16 224x224 bf16 fine-tune step (BASELINE.json metric), per-GPU batch 256.

  python bench.py --gpus N --steps K --warmup W            # this repo (engine on sm_100a)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU PyTorch path on the host cores

For N > 1 launch with torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.
"value" = device-timed images/s with inputs resident in HBM; "e2e" = the same step through the public API
(finetune.train_step) with pinned host inputs, H2D copies and the loss read-back inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                image_size=224, num_labels=120)
PER_GPU_BATCH = 256
METRIC = "vit_l16_224_train_images_per_sec"
WORKLOAD_NAME = "ViT-L/16 224x224"
# The headline (BASELINE.json metric) is the default; the other BASELINE configs can be measured with --workload.
WORKLOADS = {
    "vitl224": (dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096, image_size=224,
                     num_labels=120), 256, "vit_l16_224_train_images_per_sec", "ViT-L/16 224x224"),
    "vitb224": (dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, image_size=224,
                     num_labels=120), 256, "vit_b16_224_train_images_per_sec", "ViT-B/16 224x224"),
    "vitl384": (dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096, image_size=384,
                     num_labels=120), 128, "vit_l16_384_train_images_per_sec", "ViT-L/16 384x384"),
}


def flops_per_image_forward(c):
    N = (c["image_size"] // 16) ** 2 + 1
    D, L = c["hidden_size"], c["num_hidden_layers"]
    return L * (24 * N * D * D + 4 * N * N * D) + 2 * (N - 1) * 768 * D + 2 * D * c["num_labels"]


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(source="measured", hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"],
                    bf16_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]))
    return dict(source="fallback", hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0)


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons))
        return dict(sm_mhz=statistics.median(self.samples), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons))


# ------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own PyTorch path on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_model():
    """The class TIC/ViT/model.py:45 instantiates, built from an explicit config (the factory needs the HF hub)."""
    import torch
    try:
        from transformers import ViTConfig, ViTForImageClassification
        m = ViTForImageClassification(ViTConfig(**WORKLOAD))
        return m, "reference", "transformers.ViTForImageClassification"
    except Exception:  # transformers missing: fall back to the oracle port of the same arithmetic
        from oracle import vit_oracle as O
        sd = O.deterministic_state_dict(WORKLOAD, 0.02)

        class Port(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.p = torch.nn.ParameterDict({k.replace(".", "/"): torch.nn.Parameter(v) for k, v in sd.items()})

            def forward(self, x):
                out = O.vit_forward({k.replace("/", "."): v for k, v in self.p.items()}, x, WORKLOAD["num_attention_heads"])
                return type("Out", (), {"logits": out})()
        return Port(), "port", "oracle.vit_oracle.vit_forward"


REFERENCE_ROOT = "/root/reference"


def reference_train_step():
    """The reference's own ``train_step`` (TIC/ViT/finetune.py:54-67), imported from the reference tree when that tree is
    on this machine (the build container); the GPU boxes do not have it, so there the same statement sequence restated
    below runs. The only change to the environment the reference code runs in is the device string: it moves every batch
    ``.to("cuda")``, mapped to "cpu" here (the shim tests/golden/make_golden_ref.py uses); ``autocast('cuda')`` does not
    touch CPU tensors and the GradScaler is disabled, so the step is the plain fp32 one."""
    import torch
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "TIC", "ViT")):
        try:
            sys.path.insert(0, REFERENCE_ROOT)
            import warnings
            warnings.filterwarnings("ignore")
            import TIC.ViT.finetune as RF
            real_to = torch.Tensor.to

            def to(self, *a, **k):
                a = tuple("cpu" if (isinstance(v, str) and v.startswith("cuda")) else v for v in a)
                if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
                    k["device"] = "cpu"
                return real_to(self, *a, **k)

            def step(model, data, opt, crit, scaler):
                torch.Tensor.to = to
                try:
                    return RF.train_step(model, data, opt, crit, scaler)
                finally:
                    torch.Tensor.to = real_to
            return step, "imported " + REFERENCE_ROOT + "/TIC/ViT/finetune.py:54-67 train_step (device string 'cuda' -> 'cpu')"
        except Exception as e:  # a reference dependency missing here: say so and use the restatement
            why = f"import of the reference's finetune.py failed ({type(e).__name__}); "
    else:
        why = "reference tree absent on this box; "

    def step(model, data, opt, crit, scaler):   # finetune.py:54-67, statement for statement, without .to("cuda")
        model.train()
        opt.zero_grad()
        inputs, labels = data
        loss = crit(model(inputs).logits, labels)
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        return loss.item()
    return step, why + "train_step restated from TIC/ViT/finetune.py:54-67"


def cpu_train_steps(steps, warmup, time_budget_s, batch=8):
    """``finetune.train_step`` on the CPU: zero_grad -> forward -> CrossEntropyLoss -> backward -> AdamW(lr 1e-5, wd 0.01)
    -> loss.item(); fp32, all host threads, on a bounded sample (``batch`` images per step) of the per-GPU batch.
    Returns img/s, ms per step, kind, cores, description, the batch actually run."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    model, kind, what = cpu_reference_model()
    step_fn, step_src = reference_train_step()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5, weight_decay=0.01)
    crit = torch.nn.CrossEntropyLoss()
    scaler = torch.amp.GradScaler(enabled=False)

    def one(b):
        x = torch.randn(b, 3, WORKLOAD["image_size"], WORKLOAD["image_size"])
        y = torch.randint(0, WORKLOAD["num_labels"], (b,))
        t0 = time.perf_counter()
        step_fn(model, (x, y), opt, crit, scaler)
        return time.perf_counter() - t0

    t_first = one(batch)  # also warms the allocator / thread pool
    # keep the whole run inside the time budget by shrinking the per-step sample
    while batch > 1 and (steps + warmup) * t_first * 0.8 > time_budget_s:
        batch //= 2
        t_first = one(batch)
    for _ in range(max(0, warmup - 1)):
        one(batch)
    times = [one(batch) for _ in range(steps)]
    ms = statistics.median(times) * 1e3
    sample = (f"{what} fp32, {step_src}, AdamW lr 1e-5 wd 0.01, on {WORKLOAD_NAME}: {batch} images per step (a bounded "
              f"sample of the per-GPU batch {PER_GPU_BATCH}), {steps} timed steps after {warmup} warm-up, median; "
              f"{cores} host threads")
    return batch / (ms / 1e3), ms, kind, cores, sample, batch


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, ms, kind, cores, sample, batch = cpu_train_steps(args.steps, max(1, args.warmup), time_budget_s=200.0)
    line = base_line(args, n_gpus=args.gpus)
    # the config is this arm's: the CPU steps a bounded sample of the batch, and the line says which
    line["config"].update(per_gpu_batch=batch, global_batch=batch, parallelism="cpu", l2_policy=None,
                          sample_of=f"per-GPU batch {PER_GPU_BATCH} of the GPU arm")
    line.update(impl="reference", value=value, ms_per_step=ms, dtype="f32", vs_baseline=None,
                cpu_baseline=dict(value=value, unit="img/s", cores=cores, kind=kind, sample=sample),
                e2e=dict(value=value, unit="img/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0, roofline=None, clocks=None)
    emit(line)


def torch_gpu_diagnostic(dev):
    """Stock PyTorch on the same GPU: transformers.ViTForImageClassification (the class TIC/ViT/model.py:45 builds), bf16
    autocast as ntrain.py:241 asks Lightning for, SDPA attention, torch.optim.AdamW(fused=True). Library kernels only --
    reported as context beside the product's number, never as part of it."""
    import torch
    try:
        import warnings
        warnings.filterwarnings("ignore")
        from transformers import ViTConfig, ViTForImageClassification
        cfg = ViTConfig(**WORKLOAD)
        try:
            ref = ViTForImageClassification._from_config(cfg, attn_implementation="sdpa")
        except Exception:
            ref = ViTForImageClassification(cfg)
        ref = ref.to(dev).train()
        opt = torch.optim.AdamW(ref.parameters(), lr=1e-5, weight_decay=0.01, fused=True)
        S = WORKLOAD["image_size"]
        for batch in (PER_GPU_BATCH, PER_GPU_BATCH // 2, PER_GPU_BATCH // 4):
            try:
                x = torch.randn(batch, 3, S, S, device=dev)
                y = torch.randint(0, WORKLOAD["num_labels"], (batch,), device=dev)

                def step():
                    opt.zero_grad(set_to_none=True)
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        loss = torch.nn.functional.cross_entropy(ref(x).logits.float(), y)
                    loss.backward()
                    opt.step()
                    return loss
                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                return dict(value=round(batch / (ms / 1e3), 1), unit="img/s", ms_per_step=round(ms, 3), batch=batch,
                            what="transformers.ViTForImageClassification, torch.autocast(bf16), attn_implementation=" +
                                 str(getattr(ref.config, "_attn_implementation", "?")) + ", torch.optim.AdamW(fused=True), "
                                 "inputs resident, 5 timed steps after 3 warm-up (cuBLASLt / SDPA / ATen kernels)",
                            role="context only: stock PyTorch on the same GPU")
            except torch.OutOfMemoryError:
                opt.zero_grad(set_to_none=True)
                torch.cuda.empty_cache()
        return dict(unavailable="out of memory at every batch tried")
    except Exception as e:
        return dict(unavailable=f"{type(e).__name__}: {e}"[:200])
    finally:
        torch.cuda.empty_cache()


def committed_gemm_traffic(workload):
    import glob
    from scripts.ncu_launch_summary import gemm_sources_sha
    sha = gemm_sources_sha()
    stale = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_launches_{workload}_*_summary.json")), reverse=True):
        try:
            with open(path) as f:
                d = json.load(f)
            name = os.path.relpath(path, ROOT)
            if d.get("gemm_sources_sha") == sha:
                return float(d["gemm"]["dram_bytes_per_launch"]), (f"{name} (ncu dram__bytes_read+write, mean over the GEMM "
                                                                     f"launches of the capture, GEMM sources {sha})")
            stale = stale or f"null: {name} was captured from other GEMM sources ({d.get('gemm_sources_sha')} != {sha})"
        except Exception:
            continue
    return None, stale


def base_line(args, n_gpus):
    return dict(metric=METRIC, value=None, unit="img/s", n_gpus=n_gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=None, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="bf16", data="synthetic",
                config=dict(workload=WORKLOAD_NAME + " bf16 fine-tune step (forward, softmax-CE, backward, AdamW), "
                                     "random-init weights, synthetic N(0,1) images, int labels",
                            per_gpu_batch=PER_GPU_BATCH, global_batch=PER_GPU_BATCH * n_gpus,
                            tokens=(WORKLOAD["image_size"] // 16) ** 2 + 1,
                            parallelism=f"dp{n_gpus}", l2_policy="inputs_exceed_l2 (tens of GB of activations per step)"))


# ------------------------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from touhouimageclassification_b200 import _lib
    from touhouimageclassification_b200.finetune import train_step
    from touhouimageclassification_b200.model import ViTConfig, ViTForImageClassification
    from touhouimageclassification_b200.optim import FusedAdamW
    from touhouimageclassification_b200.parallel import DataParallelTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("for --gpus N > 1 launch with: python -m torch.distributed.run --nproc-per-node N bench.py ...")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    lib.tic_launch_count.restype = __import__("ctypes").c_int64
    lib.tic_prof_collect.restype = __import__("ctypes").c_int64

    torch.manual_seed(1234)
    model = ViTForImageClassification(ViTConfig(**WORKLOAD)).to(dev).train()
    opt = FusedAdamW(model, lr=1e-5, weight_decay=0.01)
    trainer = DataParallelTrainer(model, opt, bucket_mb=args.bucket_mb)
    trainer.broadcast_parameters(0)
    if args.diag_no_allreduce:
        trainer.bucketer.all_reduce = lambda flat, begin, end: None
    B, S = PER_GPU_BATCH, WORKLOAD["image_size"]
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x_dev = torch.randn(B, 3, S, S, device=dev, generator=g)
    y_dev = torch.randint(0, WORKLOAD["num_labels"], (B,), device=dev, generator=g)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- device-resident throughput
    for _ in range(max(3, args.warmup)):
        loss = trainer.step(x_dev, y_dev)
    barrier()
    launches0 = lib.tic_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            loss = trainer.step(x_dev, y_dev)
        e1.record()
        barrier()
    launches = lib.tic_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = PER_GPU_BATCH * world / (ms_step / 1e3)
    loss_val = float(loss.item())
    # data-parallel invariant: after every step the parameter arenas are bit-identical on all ranks. Checked on
    # an exact checksum (sum of the arena's bit patterns as int64) whose max and min over the ranks must coincide.
    ranks_identical = None
    if world > 1:
        chk = model._arena.view(torch.int32).to(torch.int64).sum().reshape(1)
        hi, lo = chk.clone(), chk.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        ranks_identical = bool((hi == lo).item())
        if not ranks_identical:
            raise SystemExit(f"rank {rank}: parameter arenas diverged across ranks after {args.steps} steps")

    # ---- one extra profiled step: per-kernel CUDA-event durations on the launching stream
    lib.tic_prof_enable(1)
    trainer.step(x_dev, y_dev)
    torch.cuda.synchronize()
    import ctypes
    buf = ctypes.create_string_buffer(1 << 16)
    n = lib.tic_prof_collect(buf, ctypes.c_int64(len(buf)))
    lib.tic_prof_enable(0)
    kernels = {}
    for ln in buf.raw[:n].decode().splitlines():
        name, cnt, ms, fl, by = ln.split("\t")
        kernels[name] = dict(launches=int(cnt), ms=float(ms), flops=float(fl), bytes=float(by))
    gemm = [v for k, v in kernels.items() if k.startswith("gemm_")]
    gemm_ms = sum(v["ms"] for v in gemm)
    gemm_flops = sum(v["flops"] for v in gemm)
    gemm_launches = sum(v["launches"] for v in gemm)
    prof_ms = sum(v["ms"] for v in kernels.values())
    peaks = measured_peaks()
    peak = peaks["bf16_tflops_sustained"]  # kernel timed inside a long step -> sustained figure
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
    step_flops = 3 * flops_per_image_forward(WORKLOAD) * PER_GPU_BATCH
    # DRAM traffic per GEMM launch: from the newest committed ncu launch list of this command (dram__bytes_read + _write
    # summed over every GEMM launch / launches) -- quoted only when that capture was made from the GEMM sources of this
    # tree (hash stamped by scripts/ncu_launch_summary.py); a capture of older kernels yields null, not a stale number
    traffic, traffic_src = committed_gemm_traffic(args.workload)
    gemm_bytes = sum(v["bytes"] for v in gemm)
    roofline = dict(bound="tensor", kernel="gemm_bf16_tcgen05_kernel (all GEMM launches of one step)",
                    achieved=achieved, peak=peak, unit="TFLOP/s", frac=(achieved / peak if achieved else None),
                    peak_source=f"{peaks['source']} bf16_tflops_sustained (burst {peaks['bf16_tflops']})",
                    traffic=traffic, traffic_unit="bytes per launch", traffic_source=traffic_src,
                    algorithmic_bytes_per_launch=(gemm_bytes / gemm_launches if gemm_launches else None),
                    launches_per_step=gemm_launches, kernel_ms_per_step=gemm_ms,
                    kernel_share_of_step=gemm_ms / prof_ms if prof_ms else None,
                    step_achieved=step_flops / (ms_step / 1e3) / 1e12,
                    step_frac_of_burst=step_flops / (ms_step / 1e3) / 1e12 / peaks["bf16_tflops"],
                    breakdown_ms={k: round(v["ms"], 3) for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])})

    # HBM-bound kernels of the step (north_star: achieved GB/s against the measured copy bandwidth): algorithmic bytes
    # (the launcher's own count) / CUDA-event duration on the launching stream, same profiled step
    roofline_hbm = {}
    for name in ("layernorm_fwd", "layernorm_bwd", "adamw"):
        k = kernels.get(name)
        if k and k["ms"] > 0:
            gbs = k["bytes"] / (k["ms"] / 1e3) / 1e9
            roofline_hbm[name] = dict(bound="hbm", achieved=round(gbs, 1), peak=peaks["hbm_gbs"], unit="GB/s",
                                      frac=round(gbs / peaks["hbm_gbs"], 4), launches_per_step=k["launches"],
                                      ms_per_step=round(k["ms"], 3), algorithmic_bytes_per_launch=k["bytes"] / k["launches"])
    if "adamw" in roofline_hbm:
        roofline_hbm["adamw"]["note"] = ("applied per gradient bucket on a side stream UNDER the backward of earlier layers: "
                                         "the duration is that of a kernel sharing the chip, not of AdamW alone")

    # ---- end to end through the public API: pinned host tensors -> finetune.train_step -> loss.item()
    x_host = torch.randn(B, 3, S, S).pin_memory()
    y_host = torch.randint(0, WORKLOAD["num_labels"], (B,)).pin_memory()
    crit = torch.nn.CrossEntropyLoss()

    def e2e_step():
        if world > 1:
            opt.zero_grad()
            xi = x_host.to(dev, non_blocking=True)
            yi = y_host.to(dev, non_blocking=True)
            return float(trainer.step(xi, yi).item())
        return train_step(model, (x_host, y_host), opt, crit, None)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    e2e = dict(value=PER_GPU_BATCH * world / (e2e_ms / 1e3), unit="img/s", ms_per_step=e2e_ms,
               h2d_bytes_per_step=(x_host.numel() * 4 + y_host.numel() * 8) * world, d2h_bytes_per_step=4 * world,
               api=("touhouimageclassification_b200.finetune.train_step(model, (x_host, y_host), FusedAdamW, CrossEntropyLoss)"
                    if world == 1 else
                    "touhouimageclassification_b200.parallel.DataParallelTrainer.step(x_host.to(dev), y_host.to(dev)).item()"))

    # ---- BASELINE config 3: ntrain's step with the train transform on the device. Pinned uint8 NHWC thumbnails (256x256,
    # the reference's source size) -> H2D -> fused augmentation -> CutMix/MixUp + patchify -> engine step -> loss.item()
    e2e_aug = None
    if S <= 224:
        from touhouimageclassification_b200.augment import GpuAugment
        from touhouimageclassification_b200.ntrain import ViTLModule
        mod = ViTLModule(WORKLOAD["num_labels"], False, "google/vit-base-patch16-224", 1e-5, 0.01, enable_mixup=True,
                         fused_optimizer=True)
        mod.vit = model                      # the module under test is the benchmarked one (no second ViT on the device)
        aug = GpuAugment(seed=1234 + rank, size=S, recipe="full")
        u8_host = torch.randint(0, 256, (B, 256, 256, 3), dtype=torch.uint8).pin_memory()
        torch.manual_seed(99 + rank)         # CutMix / MixUp draws

        def aug_step():
            xi = u8_host.to(dev, non_blocking=True)
            yi = y_host.to(dev, non_blocking=True)
            return float(mod.fused_training_step((xi, yi), opt, grad_sync=trainer._grad_sync if world > 1 else None,
                                                 world_size=world, augment=aug).item())

        for _ in range(2):
            aug_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            aug_step()
        barrier()
        aug_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
        e2e_aug = dict(value=PER_GPU_BATCH * world / (aug_ms / 1e3), unit="img/s", ms_per_step=aug_ms,
                       h2d_bytes_per_step=(u8_host.numel() + y_host.numel() * 8) * world, d2h_bytes_per_step=4 * world,
                       api="touhouimageclassification_b200.ntrain.ViTLModule.fused_training_step((uint8 NHWC 256x256, y), "
                           "FusedAdamW, augment=GpuAugment(recipe='full')) with CutMix/MixUp")
        del mod

    # ---- batched inference (BASELINE config 4: utils/filter + web serve path), one replica per GPU (= per rank)
    inference = None
    if not args.no_inference:
        from touhouimageclassification_b200.augment import IMAGENET_MEAN, IMAGENET_STD
        from touhouimageclassification_b200.serve import predict_batch, predict_batch_u8
        del trainer
        model._workspaces.clear()            # drop the 46 GB training workspace before the batch-1024 forwards
        torch.cuda.empty_cache()
        model.eval()
        fwd_flops = flops_per_image_forward(WORKLOAD)
        inference = dict(unit="img/s", dtype="bf16", replicas=world,
                         note=WORKLOAD_NAME + " forward (engine_forward), inputs resident in HBM, CUDA events; with N ranks "
                         "every rank is one replica working on its own requests (no collective) and img/s is the aggregate "
                         "over the slowest rank", batches={}, e2e={})

        def time_forward(fn, iters):
            for _ in range(3):
                fn()
            barrier()
            a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                fn()
            b2.record()
            barrier()
            return max_over_ranks(a.elapsed_time(b2)) / iters

        def time_wall(fn, iters):
            for _ in range(3):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(iters):
                fn()
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3
            barrier()
            return max_over_ranks(ms) / iters

        big = (1, 8, 64, 256, 1024) if WORKLOAD["image_size"] <= 224 else (1, 8, 64, 256)
        with torch.no_grad():
            for bs in (big if world == 1 else big[-1:]):
                xb = torch.randn(bs, 3, S, S, device=dev, generator=g)
                ms = time_forward(lambda: model.engine_forward(xb, training=False), 20 if bs <= 64 else 5)
                ips = bs * world / (ms / 1e3)
                inference["batches"][str(bs)] = dict(ms=round(ms, 4), img_per_s=round(ips, 1),
                                                     frac_of_burst_peak=round(ips / world * fwd_flops / 1e12 / peaks["bf16_tflops"], 4))
                del xb
            # end to end through the serving API with HOST buffers (the reference: file -> tensor -> forward -> .item(),
            # utils/serve.py:158-230 one image at a time, web/runtime.py:235-251 chunks of <= 64): pinned uint8 256x256
            # thumbnails -> H2D -> resize+normalise+patchify kernel -> engine forward -> softmax/top-1 kernel -> host
            # list of (class, confidence); wall clock around the call, copies inside the timed region
            from_u8 = S <= 224   # the fused resize+normalise+patchify kernel keeps the output image in shared memory: <= 224
            inference["e2e"] = dict(api=("touhouimageclassification_b200.serve.predict_batch_u8(model, pinned uint8 NHWC "
                                         "256x256 thumbnails, mean, std) -> host [(class, confidence)]") if from_u8 else
                                    ("touhouimageclassification_b200.serve.predict_batch(model, pinned fp32 NCHW preprocessed "
                                     "images) -> host [(class, confidence)]"), batches={})
            e2e_sizes = (1, 8, 64, 1024) if S <= 224 else (1, 8, 64, 256)
            for bs in (e2e_sizes if world == 1 else e2e_sizes[-2:]):
                if from_u8:
                    src = torch.randint(0, 256, (bs, 256, 256, 3), dtype=torch.uint8).pin_memory()
                    fn = lambda: predict_batch_u8(model, src, IMAGENET_MEAN, IMAGENET_STD, None, max_batch_size=1024)
                else:
                    src = torch.randn(bs, 3, S, S).pin_memory()
                    fn = lambda: predict_batch(model, src, None, max_batch_size=1024)
                ms = time_wall(fn, 20 if bs <= 64 else 5)
                inference["e2e"]["batches"][str(bs)] = dict(ms=round(ms, 4), img_per_s=round(bs * world / (ms / 1e3), 1),
                                                            h2d_bytes=src.numel() * src.element_size() * world,
                                                            d2h_bytes=bs * 8 * world)
                del src
            if world == 1:
                model.set_precision("fp32")        # the reference's no-autocast serving arithmetic (serve.py:99-101)
                xb = torch.randn(64, 3, S, S, device=dev, generator=g)
                ms = time_forward(lambda: model.engine_forward_f32(xb), 3)
                inference["fp32_mode_batch64"] = dict(ms=round(ms, 3), img_per_s=round(64 / (ms / 1e3), 1))
                model.set_precision("bf16")
                del xb

    # ---- context only (not the reference arm, not a target): what stock PyTorch gives on this same GPU -- the HF class the
    # reference instantiates, under bf16 autocast with SDPA attention and torch's fused AdamW
    torch_gpu = None
    if world == 1 and not args.no_torch_gpu:
        model._workspaces.clear()
        model._graphs.clear()
        torch.cuda.empty_cache()
        torch_gpu = torch_gpu_diagnostic(dev)

    if rank == 0:
        line = base_line(args, n_gpus=world)
        line.update(value=value, ms_per_step=ms_step, e2e=e2e, roofline=roofline, roofline_hbm=roofline_hbm,
                    gpu_launches=int(launches),
                    clocks=clocks.summary(), loss=loss_val)
        if ranks_identical is not None:
            line["ranks_bit_identical_parameters"] = ranks_identical
        if args.diag_no_allreduce:
            line["invalid"] = "diagnostic run without the gradient all-reduce"
        if e2e_aug is not None:
            line["e2e_augmented"] = e2e_aug
        if inference is not None:
            line["inference"] = inference
        if torch_gpu is not None:
            line["torch_gpu_diagnostic"] = torch_gpu
        if world == 1 and not args.no_cpu_baseline:
            v, ms, kind, cores, sample, _ = cpu_train_steps(steps=2, warmup=1, time_budget_s=60.0)
            line["cpu_baseline"] = dict(value=v, unit="img/s", cores=cores, kind=kind, sample=sample)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def protect_stdout():
    """Libraries (NCCL's version banner) may print to fd 1: keep a private copy for the ONE JSON line and point
    fd 1 at stderr for everything else."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


_JSON_OUT = None


def emit(line: dict):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bucket-mb", type=float, default=96.0, help="gradient all-reduce bucket size (N > 1)")
    ap.add_argument("--diag-no-allreduce", action="store_true",
                    help="DIAGNOSTIC ONLY (N > 1): skip the gradient all-reduce to size its cost; the line is marked invalid")
    ap.add_argument("--no-inference", action="store_true", help="skip the batched-inference sweep after the training bench")
    ap.add_argument("--no-torch-gpu", action="store_true", help="skip the stock-PyTorch-on-GPU context measurement (N = 1)")
    ap.add_argument("--workload", default="vitl224", choices=sorted(WORKLOADS),
                    help="vitl224 = the BASELINE.json headline (default); vitb224 / vitl384 = BASELINE configs 2 and 5")
    args = ap.parse_args()
    global WORKLOAD, PER_GPU_BATCH, METRIC, WORKLOAD_NAME
    WORKLOAD, PER_GPU_BATCH, METRIC, WORKLOAD_NAME = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
