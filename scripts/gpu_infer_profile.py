"""GPU probe: per-kernel breakdown of one inference forward (tic_prof_enable), ViT-L/16 224."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from touhouimageclassification_b200 import _lib
from touhouimageclassification_b200.model import ViT
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
lib = _lib.load()
lib.tic_prof_collect.restype = ctypes.c_int64
m = ViT(120, False, "google/vit-large-patch16-224").cuda().eval()
m.graph_max_batch = 0
x = torch.randn(bs, 3, 224, 224, device="cuda")
with torch.no_grad():
    for _ in range(3):
        m.engine_forward(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m.engine_forward(x)
    e1.record(); torch.cuda.synchronize()
    print(f"batch {bs}: {e0.elapsed_time(e1)/5:.3f} ms per forward")
    lib.tic_prof_enable(1)
    m.engine_forward(x)
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(1 << 16)
    n = lib.tic_prof_collect(buf, ctypes.c_int64(len(buf)))
    lib.tic_prof_enable(0)
tot = 0
for ln in buf.raw[:n].decode().splitlines():
    name, cnt, ms, fl, by = ln.split("\t")
    tot += float(ms)
    print(f"{name:24s} x{cnt:>4s} {float(ms):8.3f} ms  {float(fl)/float(ms)/1e9 if float(fl) else 0:8.0f} TFLOP/s  {float(by)/float(ms)/1e6 if float(by) else 0:8.0f} GB/s")
print(f"sum {tot:.3f} ms")
