"""bench.py's driver contract on the CPU-runnable arm: `--impl reference` prints exactly ONE JSON line on stdout with the
keys the driver reads (metric, value, unit, n_gpus, steps, warmup, ms_per_step, impl, cpu_baseline, e2e)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "3000"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "columns/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["metric"].startswith("MAF columns/sec") and d["config"]["phylogenetic_model"] == "58mammals"
    # the kind is "reference" exactly when the shim-built reference binary is there
    assert (d["cpu_baseline"]["kind"] == "reference") == os.path.exists(os.path.join(ROOT, "oracle", "_ref", "phylocsf_ref"))


def test_other_ranks_of_the_reference_arm_exit_quietly():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_committed_bench_line_has_every_key_the_contract_names():
    """The closing bench line of the round (profiles/bench_r2_final.json, written by `python bench.py` on a B200): the driver's keys,
    the roofline / cpu_baseline / e2e objects, and the three legs beyond config 3.  A schema check — it keeps bench.py's output and
    the committed evidence from drifting apart silently (bench.py itself needs a GPU)."""
    d = json.load(open(os.path.join(ROOT, "profiles", "bench_r2_final.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "stages_ms", "hbm_passes", "config4", "config5", "cli"):
        assert k in d, k
    assert d["unit"] == "columns/s" and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["vs_baseline"] is None and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["unit"] == "columns/s"
    e = d["e2e"]
    assert e["unit"] == "columns/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0.5 * d["value"] < e["value"] <= 1.02 * d["value"]
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"]))
    assert d["config4"]["scaling"] == "strong" and d["config4"]["columns"] == 250_000_000 and d["config4"]["columns_per_s"] > 0
    c5 = d["config5"]
    assert c5["alignments"] >= 1_000_000 and c5["finite_scores_in_last_batch"] > 0 and {"k_mle_expm", "k_prune<true>"} <= set(c5["roofline"])
    assert d["cli"]["columns"] == 100_000_000 and d["cli"]["process_seconds"] >= d["cli"]["tool_seconds"] > 0
