"""CPU: the reference arm of bench.py (`--impl reference`) runs without a GPU and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env_extra=None):
    env = dict(os.environ, FLOCK_BENCH_REF_PAIRS_PER_CORE="4e5")
    env.update(env_extra or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=env,
                         timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()


def test_reference_arm_prints_one_contract_line():
    lines = _run(["--impl", "reference", "--gpus", "1", "--steps", "3", "--warmup", "1"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == "pair-interactions/s" and d["value"] > 0 and d["steps"] == 3
    assert d["config"]["workload"].startswith("C2: FlockingRelative-v0 4096 envs x 100 agents")
    cb = d["cpu_baseline"]
    # "reference": the unmodified reference classes ran (baseline/_ref or /root/reference present); "port": the restatement
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert "envs x 100 agents" in cb["sample"] and cb["what"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    lines = _run(["--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"], {"RANK": "1", "WORLD_SIZE": "2",
                                                                                          "LOCAL_RANK": "1"})
    assert lines == []


def test_traffic_figures_are_keyed_to_the_kernels_sass():
    """profiles/traffic.json: every config names the SASS hash of the kernel its capture was taken from, and bench.py reports a
    traffic figure only for a matching kernel in the built library (a rebuilt, different kernel reads as None, not as stale)."""
    sys.path.insert(0, ROOT)
    import bench
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
        t = json.load(f)
    for cfg, mangled in bench.TRAFFIC_KERNELS.items():
        assert cfg in t and len(t[cfg]["sass_sha"]) == 16 and t[cfg]["bytes_per_step"] > 0
        sha = bench.kernel_sass_sha(cfg)                 # cuobjdump on the in-tree library (built by build(); no GPU needed)
        assert sha is not None and len(sha) == 16, (cfg, mangled)
        got = bench.recorded_traffic(cfg, 20)
        assert got == (t[cfg]["bytes_per_step"] * 20 if sha == t[cfg]["sass_sha"] else None)
    assert bench.kernel_sass_sha("C2", lib_path="/nonexistent.so") is None
