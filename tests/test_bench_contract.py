"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the
agreed keys, and ranks other than 0 stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None, args=()):
    env = dict(os.environ)
    env.update(env_extra or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--steps", "3", "--warmup", "1", *args], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr
    return p.stdout.strip()


def test_reference_arm_json_line():
    out = _run()
    lines = [x for x in out.splitlines() if x.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "vb_entries_K_per_s" and d["unit"] == "entries*K/s"
    assert d["vs_baseline"] is None and d["higher_is_better"] is True
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 1e5 and "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, ("--gpus", "2"))
    assert out == ""


def test_committed_bench_lines_follow_the_contract():
    """The lines of record under profiles/ (what `python bench.py` printed on the B200) carry every key of the
    contract, with consistent arithmetic: value = entries*K per iteration / time, roofline.frac = achieved / peak,
    the e2e figure counted from the bytes it copies, and a bounded CPU baseline beside it."""
    import glob
    lines = sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_bench_c5_*.json")))
    assert any(p.endswith("r2_bench_c5_default.json") for p in lines)
    for path in lines:
        d = json.loads(open(path).read())
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "clocks", "gpu_launches", "roofline"):
            assert k in d, (path, k)
        assert d["metric"] == "vb_entries_K_per_s" and d["unit"] == "entries*K/s" and d["higher_is_better"] is True
        assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
        assert d["warmup"] >= 3 and d["gpu_launches"] > 0
        c = d["config"]
        work = float(c["F"]) * float(c["N"]) * float(c["K"])
        assert abs(d["value"] - work / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
        r = d["roofline"]
        assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s"
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.3 < r["frac"] < 1.0
        # 12 flop per entry*K over the two contraction launches of an iteration
        # (this rank's columns; at N > 1 the time is the maximum over the ranks, the two pass times printed are rank 0's)
        rank0 = 12.0 * work / d["n_gpus"] / ((r["pass_rows_ms"] + r["pass_cols_ms"]) * 1e-3) / 1e12
        assert 0.95 * rank0 < r["achieved"] <= rank0 * (1 + 1e-6) if d["n_gpus"] > 1 else abs(r["achieved"] - rank0) < 1e-6 * rank0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if d.get("e2e"):
            e = d["e2e"]
            assert e["unit"] == d["unit"] and e["value"] < d["value"]
            assert e["h2d_bytes_per_step"] >= float(c["F"]) * float(c["N"]) / 8 and e["d2h_bytes_per_step"] > 0
        if path.endswith("_default.json") or path.endswith("_auto.json"):
            b = d["cpu_baseline"]
            assert b["kind"] in ("reference", "port") and b["cores"] == 1 and b["unit"] == d["unit"] and b["sample"]
            assert d["e2e"] and d["n_gpus"] == 1
