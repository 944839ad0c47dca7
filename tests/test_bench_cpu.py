"""bench.py contract pieces that run without a GPU: the reference arm (oracle port on the host cores) prints one
JSON line with the agreed keys and stays inside its wall-clock budget; the seedb200 arm refuses to run on the CPU."""
import json
import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py")] + args, capture_output=True, text=True,
                          timeout=timeout, env=env, cwd=REPO)


def test_reference_arm_json_contract_and_budget():
    t0 = time.time()
    p = run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-seconds", "6"])
    dt = time.time() - t0
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "images/sec SEED encode+VQ" and line["unit"] == "images/s"
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "images/step" in cb["sample"]
    assert dt < 240, f"reference arm took {dt:.0f} s with a 6 s budget"


def test_seedb200_arm_has_no_cpu_fallback():
    p = run(["--steps", "1", "--warmup", "1", "--no-cpu"], timeout=300)
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)


def test_both_arms_name_the_same_workload():
    """the driver compares the two arms' `config`: it must be the same object key for key at every N"""
    import argparse
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(REPO, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for wl in ("encode", "llama_prefill", "llama_decode", "pipeline", "preprocess"):
        a = argparse.Namespace(workload=wl, batch=256, vq="fp16", seq=2048, prompt=256, new_tokens=128)
        for world in (1, 8):
            c1, c2 = bench.workload_config(a, world), bench.workload_config(a, world)
            assert c1 == c2 and c1["workload"] and c1["global_batch"] >= world
    src = open(os.path.join(REPO, "bench.py")).read()
    assert src.count('"config": workload_config(') >= 6      # every arm, product and reference, goes through it
