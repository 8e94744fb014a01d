"""csrc/host_workers.cuh without CUDA: the shard workers and the claim-based staging of pageable depth frames
(the threads behind dbot_b200_set_observation and the multi-GPU handle) compiled into a small C++ program."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_workers_and_claim_based_staging(tmp_path):
    exe = str(tmp_path / "stage_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-pthread",
                    os.path.join(ROOT, "tests", "stubs", "stage_check.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok:"), r.stdout
