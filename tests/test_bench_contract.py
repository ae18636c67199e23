"""bench.py's output contract on the host side (no GPU): one JSON line on stdout, also when C code in the process
(NCCL's banner under torchrun) writes to file descriptor 1; the reference arm's line carries the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_json_line_is_alone_on_stdout_when_descriptor_one_is_noisy():
    code = (
        "import os, sys; sys.path.insert(0, %r); import bench\n"
        "sys.stdout.flush(); bench._REAL_STDOUT = os.dup(1); os.dup2(2, 1)\n"      # what run_native does under torchrun
        "os.write(1, b'NCCL version 2.x (banner written from C on descriptor 1)\\n')\n"
        "print('python-level print')\n"
        "bench.emit_line({'metric': bench.METRIC, 'value': 1.0})\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = [x for x in r.stdout.splitlines() if x.strip()]
    assert len(lines) == 1 and json.loads(lines[0])["metric"] == "audio_seconds_decoded_per_second"
    assert "NCCL version" in r.stderr and "python-level print" in r.stderr


def test_reference_arm_other_ranks_exit_without_work():
    """Under torchrun only rank 0 runs the reference arm; the other ranks print nothing and exit 0."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
