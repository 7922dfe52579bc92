"""CPU: the boost-free kflow stand-in the C++ stage classes compile against (bwaflow_b200/host/kflow_shim.h) delivers every record
exactly once through multi-worker stages over small bounded queues, and terminates."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_kflow_shim_selftest(tmp_path):
    exe = str(tmp_path / "shim_selftest")
    host = os.path.join(ROOT, "bwaflow_b200", "host")
    subprocess.run(["g++", "-O2", "-std=c++11", "-pthread", "-Wall", "-o", exe, os.path.join(host, "shim_selftest.cpp")], check=True)
    r = subprocess.run([exe, "20000"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.startswith("ok")
