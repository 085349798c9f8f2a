"""For ncu: a few eager ViT-L/16 224 inference forwards at one batch size (argv[1]) -- per-kernel device times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from touhouimageclassification_b200.model import ViT
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = ViT(120, False, "google/vit-large-patch16-224").cuda().eval()
m.graph_max_batch = 0
x = torch.randn(bs, 3, 224, 224, device="cuda")
with torch.no_grad():
    for _ in range(4):
        m.engine_forward(x)
torch.cuda.synchronize()
