"""GPU probe: per-kernel and whole-model parity against torch / HF transformers (run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from touhouimageclassification_b200 import ops
from touhouimageclassification_b200.model import ViTForImageClassification, ViTConfig

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = "cuda"
torch.manual_seed(0)

def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()

def section(name, fn):
    try:
        fn(); torch.cuda.synchronize()
    except Exception as e:
        import traceback; traceback.print_exc()
        print(f"[{name}] FAILED {type(e).__name__}: {e}", flush=True)

def t_ln():
    for D in (768, 1024):
        x = torch.randn(1000, D, device=dev) * 2 + 0.5
        g = torch.randn(D, device=dev); b = torch.randn(D, device=dev)
        y, yf, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-12, out_f32=True)
        xr = x.clone().requires_grad_(True); gr = g.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
        ref = F.layer_norm(xr, (D,), gr, br, 1e-12)
        dy = torch.randn(1000, D, device=dev).bfloat16()
        dres = torch.randn(1000, D, device=dev)
        ref.backward(dy.float())
        dx, dxb, dg, db, _ = ops.layernorm_bwd(dy, x, mean, rstd, g, dres)
        print(f"[ln D={D}] fwd f32 {rel(yf, ref):.2e} bf16 {rel(y, ref):.2e} dx {rel(dx - dres, xr.grad):.2e} dgamma {rel(dg, gr.grad):.2e} dbeta {rel(db, br.grad):.2e}", flush=True)

def t_attn():
    for (B, N, H) in ((2, 197, 12), (3, 577, 4), (2, 64, 2), (1, 1, 1), (2, 130, 3)):
        D = H * 64
        qkv = (torch.randn(B * N, 3 * D, device=dev) * 1.0).bfloat16()
        ctx, lse = ops.attention_fwd(qkv, B, N, H)
        q, k, v = [t.view(B, N, H, 64).transpose(1, 2).float().requires_grad_(True) for t in qkv.float().split(D, dim=1)]
        ref = F.scaled_dot_product_attention(q, k, v, scale=0.125)
        ref_tok = ref.transpose(1, 2).reshape(B * N, D)
        lse_ref = torch.logsumexp(q @ k.transpose(-1, -2) * 0.125, dim=-1)
        dctx = torch.randn(B * N, D, device=dev).bfloat16()
        ref_tok.backward(dctx.float())
        dqkv = ops.attention_bwd(qkv, ctx, dctx, lse, B, N, H)
        dref = torch.cat([t.grad.transpose(1, 2).reshape(B * N, D) for t in (q, k, v)], dim=1)
        dq, dk, dv = dqkv.float().split(D, dim=1); rq, rk, rv = dref.split(D, dim=1)
        print(f"[attn B{B} N{N} H{H}] ctx {rel(ctx, ref_tok):.2e} lse {rel(lse, lse_ref):.2e} dq {rel(dq, rq):.2e} dk {rel(dk, rk):.2e} dv {rel(dv, rv):.2e}", flush=True)

def t_xent():
    B, C = 37, 120
    logits = torch.randn(B, C, device=dev)
    y = torch.randint(0, C, (B,), device=dev)
    lr_ = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(lr_, y); ref.backward()
    loss, dl, correct = ops.softmax_xent(logits, y)
    print(f"[xent hard] loss {abs(loss.item() - ref.item()):.2e} dlogits {rel(dl, lr_.grad):.2e} correct {correct.item()} vs {(logits.argmax(1) == y).sum().item()}", flush=True)
    soft = torch.softmax(torch.randn(B, C, device=dev), 1)
    lr_ = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(lr_, soft); ref.backward()
    loss, dl, _ = ops.softmax_xent(logits, soft)
    print(f"[xent soft] loss {abs(loss.item() - ref.item()):.2e} dlogits {rel(dl, lr_.grad):.2e}", flush=True)

def t_adamw():
    n = 1 << 20
    p = torch.randn(n, device=dev); g = torch.randn(n, device=dev) * 0.1
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-3, weight_decay=0.01)
    m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev); sh = torch.empty(n, device=dev, dtype=torch.bfloat16)
    for step in range(1, 4):
        pr.grad = g.clone() * step
        opt.step()
        ops.adamw_step(p, g * step, m, v, sh, 1e-3, 0.9, 0.999, 1e-8, 0.01, step)
    print(f"[adamw] p {rel(p, pr):.2e} max|d| {(p - pr).abs().max().item():.2e} shadow {rel(sh, p):.2e}", flush=True)

def t_misc():
    x = torch.randn(3, 3, 224, 224, device=dev)
    pt = ops.patchify_f32(x)
    ref = x.unfold(2, 16, 16).unfold(3, 16, 16).permute(0, 2, 3, 1, 4, 5).reshape(3 * 196, 768).bfloat16()
    print(f"[patchify] exact={torch.equal(pt, ref)}", flush=True)
    dy = torch.randn(5000, 3072, device=dev).bfloat16()
    print(f"[colsum] {rel(ops.colsum_bf16(dy), dy.float().sum(0)):.2e}", flush=True)

def hf_model(cfg):
    from transformers import ViTForImageClassification as HF, ViTConfig as HFC
    return HF(HFC(hidden_size=cfg.hidden_size, num_hidden_layers=cfg.num_hidden_layers,
                  num_attention_heads=cfg.num_attention_heads, intermediate_size=cfg.intermediate_size,
                  image_size=cfg.image_size, num_labels=cfg.num_labels))

def t_model(cfg, B, tag):
    torch.manual_seed(1234)
    hf = hf_model(cfg).to(dev)
    # non-trivial biases / LN params so every gradient path is exercised
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if n.endswith("bias"): p.normal_(0, 0.02)
            if "layernorm" in n and n.endswith("weight"): p.normal_(1.0, 0.05)
    m = ViTForImageClassification(cfg).to(dev)
    m.load_state_dict(hf.state_dict(), strict=True)
    x = torch.randn(B, 3, cfg.image_size, cfg.image_size, device=dev)
    y = torch.randint(0, cfg.num_labels, (B,), device=dev)
    hf.train(); m.train()
    # fp32 ground truth
    out32 = hf(x).logits; loss32 = F.cross_entropy(out32, y); hf.zero_grad(); loss32.backward()
    g32 = {n: p.grad.clone() for n, p in hf.named_parameters()}
    # autocast calibration
    hf.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outac = hf(x).logits; lossac = F.cross_entropy(outac, y)
    lossac.backward()
    gac = {n: p.grad.clone() for n, p in hf.named_parameters()}
    # ours through autograd
    out = m(x).logits
    loss = F.cross_entropy(out.float(), y)
    loss.backward()
    print(f"[{tag}] logits rel vs fp32: ours {rel(out, out32):.3e} | torch-autocast {rel(outac, out32):.3e}; ours vs autocast {rel(out, outac):.3e}", flush=True)
    print(f"[{tag}] loss fp32 {loss32.item():.5f} autocast {lossac.item():.5f} ours {loss.item():.5f}", flush=True)
    print(f"[{tag}] top1 agree vs fp32: ours {(out.argmax(1) == out32.argmax(1)).float().mean().item():.3f} autocast {(outac.argmax(1) == out32.argmax(1)).float().mean().item():.3f}", flush=True)
    worst = []
    tot_n = tot_d = tot_na = 0.0
    for n, p in m.named_parameters():
        if p.grad is None:
            print("  missing grad", n); continue
        r = rel(p.grad, g32[n]); ra = rel(gac[n], g32[n])
        tot_n += (p.grad.float() - g32[n]).pow(2).sum().item(); tot_d += g32[n].pow(2).sum().item()
        tot_na += (gac[n].float() - g32[n]).pow(2).sum().item()
        worst.append((r, ra, n))
    worst.sort(reverse=True)
    print(f"[{tag}] grads global rel: ours {(tot_n / tot_d) ** 0.5:.3e} autocast {(tot_na / tot_d) ** 0.5:.3e}", flush=True)
    for r, ra, n in worst[:8]:
        print(f"    {n}: ours {r:.3e} autocast {ra:.3e}", flush=True)
    nonkey = [w for w in worst if "key.bias" not in w[2]]
    print(f"[{tag}] worst non-key.bias: {nonkey[0][0]:.3e} ({nonkey[0][2]})", flush=True)

def t_speed(cfg, B, tag, iters=5):
    m = ViTForImageClassification(cfg).to(dev)
    x = torch.randn(B, 3, cfg.image_size, cfg.image_size, device=dev)
    y = torch.randint(0, cfg.num_labels, (B,), device=dev)
    m.train()
    def step():
        logits = m.engine_forward(x, training=True)
        loss, dl, _ = ops.softmax_xent(logits, y, round_grad=True)
        m.grad_arena().zero_()
        m.engine_backward(dl, B)
        return loss
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
    tf = tb = 0.0
    for _ in range(iters):
        e0.record(); logits = m.engine_forward(x, training=True); e1.record()
        loss, dl, _ = ops.softmax_xent(logits, y, round_grad=True); m.grad_arena().zero_(); m.engine_backward(dl, B); e2.record()
        torch.cuda.synchronize(); tf += e0.elapsed_time(e1); tb += e1.elapsed_time(e2)
    tf /= iters; tb /= iters
    N = (cfg.image_size // 16) ** 2 + 1; D = cfg.hidden_size; L = cfg.num_hidden_layers
    fl = L * (24 * N * D * D + 4 * N * N * D) + 2 * (N - 1) * 768 * D + 2 * D * cfg.num_labels
    print(f"[{tag}] B={B} fwd {tf:.2f} ms ({fl * B / tf / 1e9:.0f} TF) bwd {tb:.2f} ms ({2 * fl * B / tb / 1e9:.0f} TF) fwd+bwd {B / (tf + tb) * 1e3:.0f} img/s = {3 * fl * B / (tf + tb) / 1e9:.0f} TFLOP/s; mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    m.eval()
    with torch.no_grad():
        for _ in range(2): m.engine_forward(x, training=False)
        torch.cuda.synchronize(); e0.record()
        for _ in range(iters): m.engine_forward(x, training=False)
        e1.record(); torch.cuda.synchronize()
    ti = e0.elapsed_time(e1) / iters
    print(f"[{tag}] inference B={B}: {ti:.2f} ms {B / ti * 1e3:.0f} img/s ({fl * B / ti / 1e9:.0f} TFLOP/s)", flush=True)

if __name__ == "__main__":
    which = sys.argv[1:] or ["kernels", "model", "speed"]
    print(torch.cuda.get_device_name(0), flush=True)
    if "kernels" in which:
        section("ln", t_ln); section("attn", t_attn); section("xent", t_xent); section("adamw", t_adamw); section("misc", t_misc)
    if "model" in which:
        tiny = ViTConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512, image_size=64, num_labels=120)
        section("tiny", lambda: t_model(tiny, 4, "tiny"))
        section("vitb", lambda: t_model(ViTConfig(), 16, "ViT-B/16"))
    if "speed" in which:
        section("speedB", lambda: t_speed(ViTConfig(), 256, "ViT-B/16"))
        L = ViTConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096)
        section("speedL", lambda: t_speed(L, 256, "ViT-L/16"))
