"""Full-size parity on the B200 against the fp32 ORACLE (not another bf16 implementation): gradients of ViT-B/16 and
ViT-L/16, the classifier-head kernels on their own, top-1 agreement on 1024 images with the noise-floor control of
SURVEY Appendix D, and a bound on the run-to-run jitter of the split-K weight gradients.

Tolerances are north_star's: logits 2e-2 (bf16) / 1e-4 (fp32 mode), gradients 3e-2, top-1 agreement >= 99.9 %
(`||a-b||_2 / ||b||_2` per tensor). The oracle runs on the same GPU in plain fp32 with TF32 off."""
import ctypes
import json
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import vit_oracle as O

pytestmark = pytest.mark.gpu
dev = "cuda"

BASE = dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, image_size=224, num_labels=120)
LARGE = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096, image_size=224, num_labels=120)
REPORT = {}


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


@pytest.fixture(autouse=True)
def _fp32_ground_truth():
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _write_report():
    """The figures the assertions below were made on, for profiles/ (the test never reads this file back)."""
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.json"), "w") as f:
            json.dump(REPORT, f, indent=1, sort_keys=True)
    except OSError:
        pass


def _model(cfg, seed=1234):
    from touhouimageclassification_b200.model import ViTConfig, ViTForImageClassification
    torch.manual_seed(seed)
    return ViTForImageClassification(ViTConfig(**cfg)).to(dev)


# ------------------------------------------------------------------------------------------------ classifier head
@pytest.mark.parametrize("B,D,C,ldh", [(256, 1024, 120, 1024), (3, 768, 120, 197 * 768), (64, 128, 10, 128), (1, 1024, 1000, 1024)])
def test_head_kernels_against_fp32(B, D, C, ldh):
    """tic_head_fwd / tic_head_bwd (classifier, modeling_vit.py:641-642 [a11]) on their own: logits, dh, dW += and
    db += against fp32 matmuls of the same bf16-rounded operands; ldh > D = the CLS rows read in place from [B, N, D]."""
    from touhouimageclassification_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device=dev).manual_seed(B * 7 + C)
    hbuf = torch.randn(B, ldh, device=dev, generator=g).bfloat16()
    w = (0.05 * torch.randn(C, D, device=dev, generator=g)).bfloat16()
    b = torch.randn(C, device=dev, generator=g)
    h = hbuf[:, :D]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    for rounded in (0, 1):
        logits = torch.empty(B, C, device=dev)
        _lib.check(lib.tic_head_fwd(p(hbuf), ctypes.c_int64(ldh), p(w), p(b), B, D, C, rounded, p(logits), st))
        ref = h.float() @ w.float().t() + b
        if rounded:
            ref = ref.bfloat16().float()
        assert rel(logits, ref) < (4e-3 if rounded else 1e-5), (rounded, rel(logits, ref))
    dl = torch.randn(B, C, device=dev, generator=g) / B
    dh = torch.zeros(B, ldh, device=dev, dtype=torch.bfloat16)
    dW0, db0 = torch.randn(C, D, device=dev, generator=g), torch.randn(C, device=dev, generator=g)
    dW, db = dW0.clone(), db0.clone()
    _lib.check(lib.tic_head_bwd(p(dl), p(hbuf), ctypes.c_int64(ldh), p(w), B, D, C, p(dh), ctypes.c_int64(ldh), p(dW), p(db), st))
    assert rel(dh[:, :D], dl @ w.float()) < 4e-3            # bf16 output
    assert dh[:, D:].abs().max().item() == 0 if ldh > D else True   # only the CLS row segment is written
    assert rel(dW - dW0, dl.t() @ h.float()) < 1e-5          # accumulated into the gradient arena
    assert rel(db - db0, dl.sum(0)) < 1e-5


# ------------------------------------------------------------------------------------------------ gradients
@pytest.mark.parametrize("name,cfg,batch", [("vit_b16_224", BASE, 16), ("vit_l16_224", LARGE, 8)])
def test_full_size_gradients_against_fp32_oracle(name, cfg, batch):
    """Every parameter gradient of a full-size training forward+backward against the fp32 oracle's autograd on the same
    weights and inputs; key.bias gradients (mathematically zero, Appendix D.3) get an absolute bound."""
    m = _model(cfg).train()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.randn(batch, 3, 224, 224, device=dev)
    y = torch.randint(0, 120, (batch,), device=dev)
    out = m(x).logits
    loss = F.cross_entropy(out.float(), y)
    loss.backward()
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref_logits = O.vit_forward(leaves, x, cfg["num_attention_heads"])
    ref_loss = O.cross_entropy(ref_logits, y)
    grads = dict(zip(leaves, torch.autograd.grad(ref_loss, list(leaves.values()))))
    assert rel(out, ref_logits) < 2e-2 and abs(loss.item() - ref_loss.item()) < 2e-2
    qb = grads["vit.encoder.layer.0.attention.attention.query.bias"].abs().max().item()
    num = den = 0.0
    worst = ("", 0.0)
    for n, p in m.named_parameters():
        assert p.grad is not None, n
        if "key.bias" in n:
            assert p.grad.abs().max().item() < 2e-2 * qb + 1e-6, n
            continue
        r = rel(p.grad, grads[n])
        worst = max(worst, (n, r), key=lambda t: t[1])
        num += (p.grad - grads[n]).pow(2).sum().item()
        den += grads[n].pow(2).sum().item()
    REPORT[f"gradients_{name}"] = dict(batch=batch, logits_rel=rel(out, ref_logits), global_rel=(num / den) ** 0.5,
                                       worst_tensor=worst[0], worst_tensor_rel=worst[1], tolerance=3e-2)
    _write_report()
    assert (num / den) ** 0.5 < 3e-2
    assert worst[1] < 3e-2, worst


def test_wgrad_run_to_run_jitter_is_bounded():
    """The weight gradients are split-K sums combined with red.global.add.f32 (gemm_tcgen05.cu): the ORDER of the fp32
    additions differs from run to run, so gradients are reproducible to fp32 rounding of a handful of partial sums, not
    bit for bit. This bounds it: two backward passes of the same batch differ by < 1e-6 relative per tensor -- four
    orders of magnitude below the bf16 error of the same gradients -- and the forward is bit-identical."""
    m = _model(BASE).train()
    x = torch.randn(32, 3, 224, 224, device=dev)
    y = torch.randint(0, 120, (32,), device=dev)
    runs = []
    for _ in range(2):
        m.zero_grad(set_to_none=False)
        m.grad_arena().zero_()
        out = m(x).logits
        F.cross_entropy(out.float(), y).backward()
        runs.append((out.detach().clone(), m.grad_arena().clone()))
    assert torch.equal(runs[0][0], runs[1][0])
    worst = 0.0
    for o, n, (name, _) in zip(m._offsets, m._numels, m.named_parameters()):
        if "key.bias" in name:
            continue
        worst = max(worst, rel(runs[0][1][o:o + n], runs[1][1][o:o + n]))
    REPORT["wgrad_run_to_run_rel"] = worst
    _write_report()
    assert worst < 1e-6, worst


# ------------------------------------------------------------------------------------------------ top-1 agreement
def test_top1_agreement_on_1024_images_with_noise_floor_control():
    """north_star: top-1 agreement >= 99.9 %. On random-init weights the logits are near-tied, so the raw figure of ANY
    bf16 implementation is noise-limited (SURVEY Appendix D consequence 2: torch autocast itself scores ~98 %). Reported
    and asserted on 1024 ViT-L images: fp32 mode raw >= 99.9 %; bf16 raw no worse than the autocast control (- 1 point);
    bf16 margin-filtered (top-2 gap of the fp32 logits > 4x the largest logit error) = 100 %."""
    transformers = pytest.importorskip("transformers")
    m = _model(LARGE).eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    hf = transformers.ViTForImageClassification(transformers.ViTConfig(**LARGE)).to(dev).eval()
    hf.load_state_dict(sd, strict=True)
    g = torch.Generator(device=dev).manual_seed(4321)
    ref, bf16, ctl, f32 = [], [], [], []
    with torch.no_grad():
        for _ in range(1024 // 128):
            x = torch.randn(128, 3, 224, 224, device=dev, generator=g)
            ref.append(O.vit_forward(sd, x, 16))
            m.set_precision("bf16")
            bf16.append(m(x).logits.float())
            m.set_precision("fp32")
            f32.append(m(x).logits.float())
            with torch.autocast("cuda", dtype=torch.bfloat16):
                ctl.append(hf(x).logits.float())
    ref, bf16, ctl, f32 = (torch.cat(t) for t in (ref, bf16, ctl, f32))
    agree = lambda a: (a.argmax(1) == ref.argmax(1)).float().mean().item()
    top2 = ref.topk(2, dim=1).values
    err = (bf16 - ref).abs().max().item()
    decided = (top2[:, 0] - top2[:, 1]) > 4 * err
    filt = (bf16.argmax(1)[decided] == ref.argmax(1)[decided]).float().mean().item()
    REPORT["top1_vit_l16_224_1024_images"] = dict(
        fp32_mode_raw=agree(f32), bf16_raw=agree(bf16), torch_autocast_control_raw=agree(ctl),
        bf16_margin_filtered=filt, margin_filtered_images=int(decided.sum()), max_abs_logit_err_bf16=err,
        logits_rel_bf16=rel(bf16, ref), logits_rel_fp32_mode=rel(f32, ref), logits_rel_autocast_control=rel(ctl, ref))
    _write_report()
    assert rel(f32, ref) < 1e-4 and agree(f32) >= 0.999
    assert rel(bf16, ref) < 2e-2
    assert agree(bf16) >= agree(ctl) - 0.01
    assert int(decided.sum()) >= 64 and filt == 1.0
