"""tests/c_driver/fake_athena.c: a C stand-in for athena's Fortran host side that drives the C
ABI in the call order of network%train (SURVEY.md sections 3.1 - 3.5) and checks the results
against the oracle.  The CPU suite builds it with gcc and checks that it fails loudly without
a device (no CPU fallback); the GPU suite runs it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_driver", "fake_athena.c")
LIBDIR = os.path.join(ROOT, "athena_b200", "lib")
ORADIR = os.path.join(ROOT, "oracle", "_build")


def _build(tmp_path):
    from oracle.oracle import build as build_oracle
    build_oracle()
    exe = str(tmp_path / "fake_athena")
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           "-o", exe, SRC, "-L", LIBDIR, "-lathena_cuda", "-L", ORADIR, "-l:liboracle_f32.so",
           f"-Wl,-rpath,{LIBDIR}", f"-Wl,-rpath,{ORADIR}", "-lm"]
    subprocess.check_call(cmd)
    return exe


def test_c_driver_builds_and_fails_loudly_without_a_device(tmp_path):
    exe = _build(tmp_path)
    import athena_b200 as ab
    if ab.lib().athena_cuda_init(-1) == 0:
        pytest.skip("a device is present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_driver_runs_the_train_loop_call_order_against_the_oracle(cuda, tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    assert "parity ok" in r.stdout
