"""GPU probe: ViT-L/16 224 inference forward latency per batch size (CUDA graph path up to 64, eager above), CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from touhouimageclassification_b200.model import ViT
m = ViT(120, False, "google/vit-large-patch16-224").cuda().eval()
out = []
with torch.no_grad():
    for bs in (1, 8, 64, 256, 1024):
        x = torch.randn(bs, 3, 224, 224, device="cuda")
        for _ in range(3):
            m.engine_forward(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20 if bs <= 64 else 5
        e0.record()
        for _ in range(n):
            m.engine_forward(x)
        e1.record(); torch.cuda.synchronize()
        out.append(f"b{bs}: {e0.elapsed_time(e1) / n:.3f} ms")
print("PDL off" if os.environ.get("TIC_NO_PDL") else "PDL on ", " | ".join(out))
