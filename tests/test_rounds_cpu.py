"""Builds and runs the CPU unit test of the batching-rounds barrier (csrc/tn_rounds.h) that synchronises the QJMC ensemble's
worker threads in TN_QJMC_BATCH=1 mode: staggered exits, early leavers, a failing round -- all under a time limit, so that a
dead-lock is a test failure here and not a hung GPU box."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rounds_barrier_never_deadlocks():
    src = os.path.join(ROOT, "tests", "cpu", "test_rounds.cpp")
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "test_rounds")
        subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-pthread", "-o", exe, src], check=True, timeout=120)
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "rounds ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
