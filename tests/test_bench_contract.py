"""bench.py's output contract, checked without a GPU: the reference arm (CPU, the oracle port) is run live on a tiny
sample, and the measured lines committed under profiles/ are checked for the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches")


def _baseline():
    return json.load(open(os.path.join(ROOT, "BASELINE.json")))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""), timeout=900)
    assert out.returncode == 0, out.stderr[-800:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in BASE + ("impl", "cpu_baseline"):
        assert k in d, k
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["higher_is_better"] is True
    assert d["unit"] == "songs/s" and d["value"] > 0 and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["metric"] == _baseline()["metric"]


def test_committed_bench_lines_keep_the_contract():
    metric = _baseline()["metric"]
    for name, n in (("bench_r02_1gpu.json", 1), ("bench_r02_2gpu.json", 2), ("bench_r02_8gpu.json", 8)):
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        for k in BASE + ("roofline", "cpu_baseline", "clocks"):
            assert k in d, (name, k)
        assert d["metric"] == metric and d["n_gpus"] == n and d["warmup"] >= 3 and d["gpu_launches"] > 0
        assert d["scaling"] == "weak" and d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None
        assert abs(d["value"] - n * d["config"]["songs_per_gpu"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
        r = d["roofline"]
        assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        e = d["e2e"]
        assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
        assert e["devices"] == n and e["bitwise_equal_to_device_path"] is True  # ONE call of one process over all n GPUs
        assert 0.5 < e["frac_of_h2d_ceiling"] <= 1.05
        assert 0 < r["step_read_frac"] < 1 and "static" in r["traffic_source"]
        c = d["clocks"]
        assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert c["sm_mhz"] >= 0.9 * c["sm_max_mhz"]
    one = json.load(open(os.path.join(ROOT, "profiles", "bench_r02_1gpu.json")))
    cb = one["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["parity_within_1e-4"] is True and cb["note"]
    sm = one["stft_microbench"]  # BASELINE.json configs[2]: >= 10 000 tracks, looped-resident stated, fractions of the measured peak
    assert sm["tracks"] >= 10000 and sm["looped_resident"] and sm["frac_read"] >= 0.25 and sm["frac_read_write"] > sm["frac_read"]
    for name in ("config4_r02_8gpu_100k.json", "config5_r02_8gpu_20k.json"):
        c = json.load(open(os.path.join(ROOT, "profiles", name)))
        assert c["n_gpus"] == 8
    c4 = json.load(open(os.path.join(ROOT, "profiles", "config4_r02_8gpu_100k.json")))
    assert c4["songs"] == 100000 and c4["checks_ok"] is True and c4["exchange"] == "p2p-fused"
    c5 = json.load(open(os.path.join(ROOT, "profiles", "config5_r02_8gpu_20k.json")))
    assert c5["songs"] >= 20000 and c5["all_ok"] and c5["parity_subset"]["within_1e-4"] and c5["parity_subset"]["songs"] >= 2000
    assert c5["playlist_from_seed"]["first_k_exact"] and c5["playlist_from_seed"]["kendall_tau"] > 0.999
