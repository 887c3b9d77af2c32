"""The reference arm of bench.py (`--impl reference`): one JSON line with the contract's keys; under torchrun (N > 1) rank 0
alone runs and prints it, the other ranks exit 0 without work.  CPU tier (the arm never touches a GPU)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def _check(line, gpus):
    assert line["impl"] == "reference" and line["n_gpus"] == gpus and line["higher_is_better"] is True
    assert line["metric"] == "eskf_frames_per_sec" and line["unit"] == "frames/s" and line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    e = line["e2e"]
    assert e["value"] == line["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]


def test_reference_arm_single_process():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "T1", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _lines(r.stdout)
    assert len(lines) == 1
    _check(lines[0], 1)


def test_reference_arm_under_torchrun_rank0_only():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "T1",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _lines(r.stdout)
    assert len(lines) == 1, r.stdout[-2000:]
    _check(lines[0], 2)
