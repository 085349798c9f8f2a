import sys; sys.path.insert(0, "/root/repo")
import torch
from oracle import vit_oracle as O
from touhouimageclassification_b200.model import ViTConfig, ViTForImageClassification
from touhouimageclassification_b200 import serve as S
cfg = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, image_size=32, num_labels=10)
def mk(dev):
    m = ViTForImageClassification(ViTConfig(**cfg)); m.load_state_dict(O.deterministic_state_dict(cfg, 0.05), strict=True); return m.to(dev).eval()
m = mk("cuda:0")
u8 = torch.randint(0, 256, (37, 48, 40, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(5))
mean, std = (0.5, 0.4, 0.3), (0.2, 0.25, 0.3)
def logits(model, x):
    dev = model._arena.device
    with torch.no_grad(), torch.cuda.device(dev):
        p = S.preprocess_u8(x.to(dev), mean, std, 32)
        return model.engine_forward(patches=p).float().cpu(), p.float().cpu()
la, pa = logits(m, u8)
lb, pb = logits(m, u8[19:])
print("same device, batch 37 vs chunk [19:37]: patches equal", torch.equal(pa[19*4:], pb), "logits max diff", (la[19:] - lb).abs().max().item(), "argmax equal", torch.equal(la[19:].argmax(1), lb.argmax(1)))
top2 = la.topk(2, 1).values; print("min top-2 gap", (top2[:,0]-top2[:,1]).min().item())
if torch.cuda.device_count() > 1:
    m1 = mk("cuda:1")
    lc, pc = logits(m1, u8[19:])
    print("cuda:1 chunk vs cuda:0 chunk: patches equal", torch.equal(pb, pc), "logits max diff", (lb - lc).abs().max().item())
    ld, pd = logits(m1, u8)
    print("cuda:1 full vs cuda:0 full: patches equal", torch.equal(pa, pd), "logits max diff", (la - ld).abs().max().item())
