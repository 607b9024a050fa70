"""The memory-safety substitute (compute-sanitizer is closed on this pool): rtfs_core.cuh's device functions compiled
for the host with -fsanitize=address,undefined (csrc/host_debug/), driven by a scalar version of the frame, checked
against the oracle — both trees, every scene family.  A deliberate walk off the traversal stack must kill the process."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ray_tracing_fsharp_b200", "csrc")
WORKER = os.path.join(ROOT, "tests", "host_debug_worker.py")


@pytest.fixture(scope="module")
def asan_env():
    subprocess.check_call(["make", "-C", CSRC, "-s", "host-debug"])
    libasan = subprocess.check_output(["gcc", "-print-file-name=libasan.so"], text=True).strip()
    if not os.path.exists(libasan):
        pytest.skip("libasan.so not found")
    env = dict(os.environ)
    env["LD_PRELOAD"] = libasan
    env["ASAN_OPTIONS"] = "detect_leaks=0:abort_on_error=1:halt_on_error=1"
    env["UBSAN_OPTIONS"] = "halt_on_error=1:print_stacktrace=1"
    return env


def run(env, *args, timeout=900):
    return subprocess.run([sys.executable, WORKER, *args], env=env, capture_output=True, text=True, timeout=timeout)


def test_frames_of_every_scene_family_under_asan_match_the_oracle(asan_env):
    res = run(asan_env, "frames")
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-6000:]
    assert "ERROR: AddressSanitizer" not in res.stderr and "runtime error" not in res.stderr, res.stderr[-6000:]
    assert res.stdout.count("byte-identical") == 12


def test_closest_hits_of_both_trees_under_asan_equal_the_oracle(asan_env):
    res = run(asan_env, "hits")
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-6000:]
    assert "ERROR: AddressSanitizer" not in res.stderr and "runtime error" not in res.stderr, res.stderr[-6000:]


def test_an_out_of_range_stack_push_is_caught(asan_env):
    ok = run(asan_env, "overflow", "64")
    assert ok.returncode == 0 and "returned" in ok.stdout, ok.stderr[-3000:]
    bad = run(asan_env, "overflow", "65")
    assert bad.returncode != 0 and "returned" not in bad.stdout
    assert "rtfs bounds check failed" in bad.stderr or "AddressSanitizer" in bad.stderr, bad.stderr[-3000:]
