"""The reference arm of bench.py (`--impl reference`) runs on host cores only, so its JSON contract can be checked
without a GPU: one line, the keys the driver reads, `impl: reference`, a cpu_baseline block describing the run and an
e2e block that repeats the value with zero transfer bytes.  With the Python reference present (/root/reference in the
build container, oracle/_ref on the GPU box) the line's value is the unmodified reference (kind "reference") and the
C++ restatement is reported beside it (`port`); without it (--no-pyref) the port is the value (kind "port")."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--gpus", "1", "--steps", "1", "--warmup", "1", "--columns", "256", "--nsteps", "240", "--sites", "4",
         "--cpu-sample-columns", "32"]


def _run(extra, env=None, timeout=900):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *extra],
                         capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.startswith("{")]


def _common(d):
    assert d["impl"] == "reference" and d["metric"].startswith("column-timesteps/sec")
    assert d["unit"] == "column-timesteps/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and "workload" in d["config"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1


def test_reference_arm_port_only():
    lines = _run(SMALL + ["--no-pyref"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    _common(d)
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and "32 columns" in cb["sample"] and "alive column-steps" in cb["sample"]
    assert d["fwd_grad"]["value"] > 0


def test_reference_arm_python_reference():
    have = any(os.path.isdir(os.path.join(p, "dpLGAR")) for p in ("/root/reference", os.path.join(ROOT, "oracle", "_ref")))
    lines = _run(SMALL + ["--pyref-rows", "24", "--pyref-procs", "2"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    _common(d)
    if have:
        assert d["cpu_baseline"]["kind"] == "reference" and "unmodified Python reference" in d["cpu_baseline"]["sample"]
        assert d["port"]["kind"] == "port" and d["port"]["value"] > d["value"]
        assert d["fwd_grad"]["value"] > 0
    else:
        assert d["cpu_baseline"]["kind"] == "port"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_segments_cover_the_record_once():
    sys.path.insert(0, ROOT)
    import bench
    for T, K in ((8760, 20), (8760, 7), (100, 20), (5, 20), (64, 1)):
        segs = bench.segments(T, K)
        assert segs[0][0] == 0 and segs[-1][1] == T and len(segs) <= K
        assert all(a[1] == b[0] for a, b in zip(segs, segs[1:])) and all(t1 > t0 for t0, t1 in segs)
