"""include/dbot_b200_adapter.hpp (the C++ dbot::RbSensor<State> a maintainer would add)
compiles against stand-ins of dbot's headers, links with libdbot_b200.so and -- on a machine without a
GPU -- turns the C-ABI's error code into a C++ exception instead of falling back to anything."""
import os
import subprocess

from dbot_ros_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_adapter_compiles_links_and_reports_errors(tmp_path):
    exe = str(tmp_path / "adapter_check")
    libdir = os.path.dirname(abi.LIB_PATH)
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(ROOT, "tests", "stubs"), os.path.join(ROOT, "tests", "stubs", "adapter_check.cpp"),
                    "-L", libdir, "-ldbot_b200", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "created" in r.stdout or "no CPU fallback" in r.stdout, r.stdout
