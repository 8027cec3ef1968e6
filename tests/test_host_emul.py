"""CPU: the CUDA NTT / LDE tile code (csrc/ntt.cuh) compiled with g++ and run thread-by-thread on the host
against textbook transforms -- catches index / twiddle / planner bugs without a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
BUILD = os.path.join(ROOT, "tests", "host_emul", "build")


def _build(name):
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(ROOT, "tests", "host_emul", name + ".cpp")
    exe = os.path.join(BUILD, name)
    deps = [src, os.path.join(ROOT, "stark_perpetual_b200", "csrc", "ntt.cuh"),
            os.path.join(ROOT, "stark_perpetual_b200", "csrc", "fp.cuh")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, src])
    return exe


@pytest.mark.parametrize("args", ["5 0 0", "11 1 0", "12 0 1", "13 1 1", "14 0 1 3", "9 0 1 5"])
def test_emulated_ntt_passes(args):
    exe = _build("emul_ntt")
    out = subprocess.run([exe] + args.split(), capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("args", ["3 3", "10 1", "12 3"])
def test_emulated_lde(args):
    exe = _build("emul_lde")
    out = subprocess.run([exe] + args.split(), capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
