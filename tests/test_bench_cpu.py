"""Host-side logic of bench.py that the reported numbers hang on (no GPU): the FLOP model is SURVEY section 8(d)'s, and the
ncu DRAM-traffic figure is only quoted from a capture of the current GEMM sources."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_flop_model_matches_the_survey():
    # forward GFLOP per image -- B/16-224 35.126, L/16-224 123.108, L/16-384 382.131; a train step is 3x
    want = {"vitb224": 35.126e9, "vitl224": 123.108e9, "vitl384": 382.131e9}
    for name, gflop in want.items():
        cfg = bench.WORKLOADS[name][0]
        assert abs(bench.flops_per_image_forward(cfg) - gflop) / gflop < 1e-4, name


def test_traffic_is_only_quoted_from_a_capture_of_the_current_gemm_sources(tmp_path, monkeypatch):
    from scripts.ncu_launch_summary import gemm_sources_sha
    sha = gemm_sources_sha()
    value, src = bench.committed_gemm_traffic("vitl224")
    assert (value is None and (src is None or src.startswith("null:"))) or (value > 0 and sha in src)
    # a summary stamped with another hash is refused, one stamped with the current hash is used
    prof = tmp_path / "profiles"
    prof.mkdir()
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    (prof / "r09_launches_vitl224_v1_summary.json").write_text(json.dumps(
        {"gemm_sources_sha": "0" * 16, "gemm": {"dram_bytes_per_launch": 1.0}}))
    value, src = bench.committed_gemm_traffic("vitl224")
    assert value is None and "other GEMM sources" in src
    (prof / "r09_launches_vitl224_v2_summary.json").write_text(json.dumps(
        {"gemm_sources_sha": sha, "gemm": {"dram_bytes_per_launch": 5.5e8}}))
    value, src = bench.committed_gemm_traffic("vitl224")
    assert value == 5.5e8 and "v2_summary" in src
    assert bench.committed_gemm_traffic("vitl384") == (None, None)   # no capture of that workload in this tree


def test_launch_summary_reproduces_the_committed_figures(tmp_path):
    """scripts/ncu_launch_summary.py on the committed ncu launch list of the final tree: the per-kernel shares and the DRAM
    bytes per GEMM launch that DESIGN.md and bench.py quote come out of the raw CSV, not out of a hand-edited file."""
    import subprocess
    csv = os.path.join(ROOT, "profiles", "r02_launches_vitl224_v3.csv")
    out = tmp_path / "s.json"
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_launch_summary.py"), csv, "--json", str(out)],
                   check=True, capture_output=True)
    got = json.loads(out.read_text())
    with open(os.path.join(ROOT, "profiles", "r02_launches_vitl224_v3_summary.json")) as f:
        committed = json.load(f)
    assert got["launches"] == committed["launches"] == 1000
    assert abs(got["gemm"]["dram_bytes_per_launch"] - committed["gemm"]["dram_bytes_per_launch"]) < 1.0
    assert 0.70 < got["gemm"]["share"] < 0.80                      # GEMMs: three quarters of the device time of a step
    names = list(got["kernels"])
    assert any(n.startswith("attn_bwd_fused_kernel") for n in names) and any(n.startswith("ln_bwd_kernel") for n in names)
    assert not any("elementwise" in n and got["kernels"][n]["share"] > 0.01 for n in names)   # no library kernel above 1 %
