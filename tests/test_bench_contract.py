"""The bench.py JSON contract, checked on the committed bench lines under profiles/ (produced on B200 boxes by the very
command the driver runs): every key the task statement names is present and self-consistent.  CPU test, no GPU needed."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r01_*.json")))
LINES_R02 = sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r02_*.json")))


def _load(path):
    return json.loads(open(path, encoding="utf-8").read().strip().splitlines()[-1])


def test_bench_lines_exist_for_every_workload_and_gpu_count():
    names = {os.path.basename(p) for p in LINES}
    for want in ("bench_r01_quartznet15x5.json", "bench_r01_citrinet1024.json", "bench_r01_features.json",
                 "bench_r01_quartznet15x5_train.json", "bench_r01_quartznet15x5_reference_arm.json",
                 "bench_r01_quartznet15x5_2gpu.json", "bench_r01_quartznet15x5_4gpu.json", "bench_r01_quartznet15x5_8gpu.json",
                 "bench_r01_quartznet15x5_train_2gpu.json", "bench_r01_quartznet15x5_train_4gpu.json"):
        assert want in names, want


@pytest.mark.parametrize("path", LINES, ids=[os.path.basename(p) for p in LINES])
def test_bench_line_contract(path):
    d = _load(path)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e"):
        assert k in d, k
    assert ("cpu_baseline" in d) == (d["n_gpus"] == 1)          # timed on rank 0 at N = 1 only
    assert d["metric"] == "audio-sec/sec" and d["unit"] == "audio-s/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 2 and d["steps"] >= 1 and d["value"] > 0 and d["ms_per_step"] > 0
    e = d["e2e"]
    assert set(e) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and e["unit"] == d["unit"]
    c = d.get("cpu_baseline")
    if c is not None:
        assert set(c) >= {"value", "unit", "cores", "kind", "sample"} and c["kind"] in ("port", "reference") and c["cores"] >= 1
    if d.get("impl") == "reference":
        assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0 and e["value"] == d["value"] == c["value"]
        return
    # our arm: host-buffer e2e really moves bytes and cannot beat the device-only number; kernels were launched
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] <= d["value"] * 1.001
    assert d["gpu_launches"] > 0
    assert d["dtype"] in ("bf16", "f32")
    clk = d["clocks"]
    assert clk["sm_mhz"] > 0 and clk["sm_max_mhz"] >= clk["sm_mhz"]
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(clk["reasons"])
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == ("GB/s" if r["bound"] == "hbm" else "TFLOP/s")
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-6) and 0 < r["frac"] < 1
    assert "traffic" in r
    # value is the whole-job aggregate: audio seconds of all ranks / step time
    per_gpu = d["config"]["per_gpu_batch"] * d["config"]["seconds"]
    assert d["value"] == pytest.approx(per_gpu * d["n_gpus"] / (d["ms_per_step"] * 1e-3), rel=1e-3)


@pytest.mark.parametrize("path", LINES_R02, ids=[os.path.basename(p) for p in LINES_R02])
def test_bench_line_contract_round2(path):
    """Round-2 lines: strong scaling by default (the configuration's batch is SHARDED over the GPUs), the Citrinet-1024 leg
    on the default line, reference arms that say what they ran."""
    d = _load(path)
    if d.get("impl") == "reference_cuda":
        assert d["config"]["note"].startswith("unmodified reference modules") and d["value"] > 0
        assert "autocast_bf16" in d
        return
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e"):
        assert k in d, k
    assert d["metric"] == "audio-sec/sec" and d["unit"] == "audio-s/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    if d.get("impl") == "reference":
        c = d["cpu_baseline"]
        assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] == d["value"] == d["e2e"]["value"]
        assert d["config"]["sample_batch_per_step"] < d["config"]["global_batch"] and "bounded sample" in d["config"]["note"]
        return
    cfg = d["config"]
    assert d["scaling"] in ("strong", "weak")
    if d["scaling"] == "strong":      # per-GPU work shrinks with N: global batch fixed
        assert cfg["per_gpu_batch"] * d["n_gpus"] >= cfg["global_batch"] > cfg["per_gpu_batch"] * (d["n_gpus"] - 1)
    else:
        assert cfg["global_batch"] == cfg["per_gpu_batch"] * d["n_gpus"]
    assert d["value"] == pytest.approx(cfg["global_batch"] * cfg["seconds"] / (d["ms_per_step"] * 1e-3), rel=1e-3)
    e = d["e2e"]
    # (the pipelined e2e loop of the training step may beat the synchronous device loop by timing noise: 2 %)
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.02
    assert d["gpu_launches"] > 0 and d["dtype"] in ("bf16", "f16", "f32")
    clk = d["clocks"]
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(clk["reasons"])
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-6) and 0 < r["frac"] < 1
    if "also" in d:                   # both networks of BASELINE.json's metric on one line
        a = d["also"]["citrinet1024"]
        assert a["value"] > 0 and a["e2e"]["value"] > 0 and a["roofline"]["frac"] > 0 and "Citrinet-1024" in a["config"]["workload"]
    assert ("cpu_baseline" in d) == (d["n_gpus"] == 1 and "also" in d or d["n_gpus"] == 1 and "cpu_baseline" in d)
