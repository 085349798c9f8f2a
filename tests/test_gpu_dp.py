"""Data-parallel parity on real GPUs: needs two or more GPUs on the box, skipped otherwise."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_step_matches_one_gpu_step_and_ranks_stay_identical():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "gpu_dp_parity.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "data-parallel parity ok" in out.stdout


def _tiny_model(device):
    from oracle import vit_oracle as O
    from touhouimageclassification_b200.model import ViTConfig, ViTForImageClassification
    cfg = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, image_size=32, num_labels=10)
    m = ViTForImageClassification(ViTConfig(**cfg))
    m.load_state_dict(O.deterministic_state_dict(cfg, 0.05), strict=True)
    return m.to(device).eval()


@pytest.mark.gpu
def test_replica_pool_round_robins_and_keeps_request_order():
    """serve.ReplicaPool (one replica per GPU, requests round-robined, no collective): results equal the
    single-model answers in request order, whatever the chunking; with one device the pool is just that model."""
    from touhouimageclassification_b200 import serve as S
    devices = [f"cuda:{i}" for i in range(min(2, torch.cuda.device_count()))]
    m = _tiny_model(devices[0])
    u8 = torch.randint(0, 256, (37, 48, 40, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(5))
    mean, std = (0.5, 0.4, 0.3), (0.2, 0.25, 0.3)
    want = S.predict_batch_u8(m, u8, mean, std, None, max_batch_size=64)
    pool = S.ReplicaPool(m, devices)
    try:
        assert len(pool) == len(devices)
        for chunk in (None, 5, 16):
            got = pool.predict_u8(u8.pin_memory(), mean, std, None, chunk=chunk)
            assert [g[0] for g in got] == [w[0] for w in want]
            assert max(abs(g[1] - w[1]) for g, w in zip(got, want)) < 1e-5
    finally:
        pool.close()


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_model_on_second_device_without_set_device():
    """Every C-ABI launch is made on the device its buffers live on (device guard), not on the process's current device:
    a model on cuda:1 gives the cuda:0 logits while the current device stays 0, forward and backward."""
    import torch.nn.functional as F
    from oracle import vit_oracle as O
    assert torch.cuda.current_device() == 0
    a, b = _tiny_model("cuda:0").train(), _tiny_model("cuda:1").train()
    x = O.deterministic_images(3, 32, seed=1)
    y = torch.tensor([0, 3, 7])
    la = F.cross_entropy(a(x.to("cuda:0")).logits.float(), y.to("cuda:0"))
    lb = F.cross_entropy(b(x.to("cuda:1")).logits.float(), y.to("cuda:1"))
    la.backward(); lb.backward()
    assert torch.cuda.current_device() == 0
    assert abs(la.item() - lb.item()) < 1e-6
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        if "key.bias" in n:
            continue
        assert torch.allclose(p.grad.cpu(), q.grad.cpu(), rtol=1e-4, atol=1e-7), n
