"""bench.py contract pieces that can be checked without a GPU: the byte model of SURVEY §8d, the
defaults, and the reference arm (the oracle port of the reference's CPU path) end to end."""
import json
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))


def test_spmm_byte_model_matches_survey():
    import bench
    # cfg-2 fp32 0.575 GB/layer (nnz 2 M, N 71 k), bf16 0.310 GB; cfg-3 fp32 534.5 GB
    assert abs(bench.spmm_layer_bytes(2_000_000, 71_000, 64, 4) / 1e9 - 0.575) < 0.005
    assert abs(bench.spmm_layer_bytes(2_000_000, 71_000, 64, 2) / 1e9 - 0.310) < 0.005
    assert abs(bench.spmm_layer_bytes(1_000_000_000, 12_000_000, 128, 4) / 1e9 - 534.5) < 1.0


def test_defaults_follow_the_timing_rules(monkeypatch):
    import bench
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse()
    assert a.gpus == 1 and a.warmup >= 3 and a.steps >= 1 and a.impl == "ours" and a.workload == "cfg2"
    assert bench.workload_name("cfg-2", 1, 3, 64, 2048, 30000, 39771, 2015064) == \
        bench.workload_name("cfg-2", 1, 3, 64, 2048, 30000, 39771, 2015064)
    assert " x8:" in bench.workload_name("cfg-2", 8, 3, 64, 2048, 1, 1, 1)


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    j = json.loads(out.stdout.strip().splitlines()[-1])
    assert j["impl"] == "reference" and j["unit"] == "edges*layers/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["config"]["workload"].startswith("cfg-2: LightGCN 3-layer d=64 BPR B=2048")
