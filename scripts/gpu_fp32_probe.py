"""GPU probe: accuracy and per-kernel time of the fp32 inference mode (ViT-L/16 224, batch 64)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import vit_oracle as O
from touhouimageclassification_b200 import _lib
from touhouimageclassification_b200.model import ViT
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
lib = _lib.load()
lib.tic_prof_collect.restype = ctypes.c_int64
torch.manual_seed(1234)
m = ViT(120, False, "google/vit-large-patch16-224").cuda().eval().set_precision("fp32")
sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
x = torch.randn(64, 3, 224, 224, device="cuda")
with torch.no_grad():
    ours = m(x).logits
    ref = O.vit_forward(sd, x, 16)
    ref64 = O.vit_forward({k: v.double() for k, v in sd.items()}, x.double(), 16).float()
rel = lambda a, b: float((a - b).norm() / b.norm())
print(f"fp32 mode vs torch fp32: {rel(ours, ref):.3e}; vs torch fp64: {rel(ours, ref64):.3e}; torch fp32 vs fp64: {rel(ref, ref64):.3e}")
with torch.no_grad():
    lib.tic_prof_enable(1)
    m(x)
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(1 << 16)
    n = lib.tic_prof_collect(buf, ctypes.c_int64(len(buf)))
    lib.tic_prof_enable(0)
tot = 0
for ln in buf.raw[:n].decode().splitlines():
    name, cnt, ms, fl, by = ln.split("\t")
    tot += float(ms)
    print(f"{name:24s} x{cnt:>4s} {float(ms):8.3f} ms")
print(f"sum {tot:.3f} ms")
