"""bench.py's contract, checked without a GPU: the committed evidence under profiles/ belongs to the kernel sources in the
tree (otherwise bench.py reports roofline.traffic = null), and the committed bench lines carry every key the driver reads."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_traffic_capture_matches_the_kernel_sources():
    import bench
    cap = json.load(open(os.path.join(ROOT, "profiles", "r02_stream_traffic.json")))
    assert cap["csrc_sha"] == bench._csrc_sha(), "re-capture profiles/r02_stream_traffic.json (scripts/gpu_round.sh ... ncu; scripts/ncu_traffic.py)"
    assert 0.9 * 17_104_896 < cap["dram_bytes_per_launch"] < 1.1 * 17_104_896       # traffic = algorithmic bytes
    assert all(l["kernel"].startswith("void dm_k_stream<0, 0>") and l["registers"] <= 64 for l in cap["launches"])


def test_committed_bench_lines_follow_the_contract():
    ours = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_n1.json")))
    ref = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_reference_n1.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in ours, k
    assert ours["config"] == ref["config"] and ref["impl"] == "reference" and ours["metric"] == ref["metric"]
    r = ours["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"]
    assert ours["e2e"]["h2d_bytes_per_step"] > 0 and ours["e2e"]["d2h_bytes_per_step"] > 0 and ours["gpu_launches"] > 0
    assert {"config4_windows", "config5_varlen", "config3_nng_pipeline"} <= set(ours["extra"]["configs"])
    assert ours["cpu_baseline"]["kind"] == "port" and "reference_engine" in ours["cpu_baseline"]
    for n in (2, 4, 8):
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_scaling", f"bench_n{n}.json")))
        assert d["n_gpus"] == n and d["value"] > 0.95 * n * ours["value"]            # weak-scaling efficiency >= 0.95
