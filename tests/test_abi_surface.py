"""The drop-in boundary without a GPU: libdbot_b200.so loads, exports every symbol that
include/dbot_b200.h declares, the ctypes stub types every one of them, the struct layouts agree,
and the product path fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from dbot_ros_b200 import (CameraData, DbotB200Error, ObjectModel, Parameters, RbSensorBuilder, abi,
                           synth)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dbot_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dbot_b200_[a-z0-9_]+)\s*\(", text)))


def has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_is_built_in_tree():
    assert os.path.exists(abi.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.dirname(abi.LIB_PATH) == os.path.join(ROOT, "dbot_ros_b200")


def test_every_declared_symbol_is_exported_and_typed():
    names = declared_symbols()
    assert len(names) >= 20
    lib = abi.load_library()
    for nm in names:
        assert hasattr(lib, nm), "libdbot_b200.so does not export %s" % nm
        assert nm in abi.SIGNATURES, "abi.py has no signature for %s" % nm
    assert sorted(abi.SIGNATURES) == names  # and the stub binds nothing the header does not declare
    assert lib.dbot_b200_abi_version() == 2


def test_no_cpp_symbols_leak_as_entry_points():
    """only extern "C" names start with dbot_b200_ in the dynamic symbol table"""
    out = subprocess.run(["nm", "-D", "--defined-only", abi.LIB_PATH], capture_output=True, text=True,
                         check=True).stdout
    exported = sorted(ln.split()[-1] for ln in out.splitlines() if " T " in ln and "dbot_b200_" in ln
                      and not ln.split()[-1].startswith("_Z"))
    assert exported == declared_symbols()


def test_library_holds_sm100a_code_and_tma():
    """the shipped binary is sm_100a SASS with TMA bulk copies (UBLKCP) -- not a generic PTX build"""
    sass = subprocess.run(["cuobjdump", "-sass", abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass            # cp.async.bulk (TMA 1-D) mesh staging
    assert "LDG.E.ENL2.256" in sass or ".256" in sass  # 256-bit streaming loads


def test_config_struct_layout_matches_header():
    """compile a 5-line C program against the header and compare sizeof/offsetof with ctypes"""
    src = r"""
    #include <stdio.h>
    #include <stddef.h>
    #include "dbot_b200.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(dbot_b200_config), offsetof(dbot_b200_config, K),
             offsetof(dbot_b200_config, max_particles), offsetof(dbot_b200_config, vertices),
             offsetof(dbot_b200_config, tail_weight), offsetof(dbot_b200_config, delta_time),
             offsetof(dbot_b200_config, device_id));
      return 0; }
    """
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "t.c"), "w") as f:
            f.write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o",
                        os.path.join(d, "t")], check=True)
        got = [int(x) for x in subprocess.run([os.path.join(d, "t")], capture_output=True, text=True,
                                               check=True).stdout.split()]
    want = [C.sizeof(abi.Config), abi.Config.K.offset, abi.Config.max_particles.offset,
            abi.Config.vertices.offset, abi.Config.tail_weight.offset, abi.Config.delta_time.offset,
            abi.Config.device_id.offset]
    assert got == want


def test_argument_errors_need_no_device():
    lib = abi.load_library()
    h = C.c_void_p()
    assert lib.dbot_b200_create(None, C.byref(h)) == -1
    assert b"null" in lib.dbot_b200_last_error()
    V, T = synth.mesh_set("ico80")
    kw = synth.default_parameters()
    cfg, keep = abi.make_config(0, 80, synth.camera_matrix(8), 4, V, T, **kw)
    assert lib.dbot_b200_create(C.byref(cfg), C.byref(h)) == -2       # bad resolution
    K = synth.camera_matrix(8)
    K[1, 0] = 0.5
    cfg, keep = abi.make_config(60, 80, K, 4, V, T, **kw)
    assert lib.dbot_b200_create(C.byref(cfg), C.byref(h)) == -3       # K not upper triangular
    bad = dict(kw, p_occluded_occluded=0.1)
    cfg, keep = abi.make_config(60, 80, synth.camera_matrix(8), 4, V, T, **bad)
    assert lib.dbot_b200_create(C.byref(cfg), C.byref(h)) == -4       # c = p_oo - p_ov must be in (0,1)
    many = synth.mesh_set("ico80")
    cfg, keep = abi.make_config(60, 80, synth.camera_matrix(8), 4, many[0] * 9, many[1] * 9, **kw)
    assert lib.dbot_b200_create(C.byref(cfg), C.byref(h)) == -2       # more than DBOT_B200_MAX_OBJECTS
    cfg, keep = abi.make_config(60, 80, synth.camera_matrix(8), 0, V, T, **kw)
    assert lib.dbot_b200_create(C.byref(cfg), C.byref(h)) == -2       # max_particles must be positive
    cfg, keep = abi.make_config(60, 80, synth.camera_matrix(8), 4, V, T, **dict(kw), occlusion_mode=1)
    assert lib.dbot_b200_create(C.byref(cfg), C.byref(h)) == -2       # reserved mode
    assert h.value is None
    f = C.c_void_p()
    fc = abi.FilterConfig()
    assert lib.dbot_b200_filter_create(None, C.byref(fc), C.byref(f)) == -1 and f.value is None
    assert lib.dbot_b200_filter_track(None, None, None, None) == -1
    assert lib.dbot_b200_export_slot_device(None, 0, None) == -1
    assert lib.dbot_b200_slot_floats(None) == -1
    assert lib.dbot_b200_reset(None) == -1
    assert lib.dbot_b200_loglikes(None, None, None, 0, 0, None) == -1
    assert lib.dbot_b200_n_slots(None) == -1


@pytest.mark.skipif(has_gpu(), reason="checks the no-device failure mode")
def test_no_cpu_fallback_without_a_device():
    """without a CUDA device building the sensor raises -- there is no CPU path to fall back to"""
    V, T = synth.mesh_set("ico80")
    rows, cols = synth.resolution(8)
    b = RbSensorBuilder(ObjectModel(V, T), CameraData(synth.camera_matrix(8), rows, cols),
                        Parameters(sample_count=4))
    with pytest.raises(DbotB200Error) as e:
        b.build()
    assert e.value.code > 0 and "no CPU fallback" in str(e.value)


def test_missing_library_is_a_hard_error(tmp_path):
    with pytest.raises(ImportError):
        abi.load_library(str(tmp_path / "libdbot_b200.so"))


def test_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under dbot_ros_b200/ may name it"""
    pkg = os.path.join(ROOT, "dbot_ros_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle_lib" not in text and "libdbot_oracle" not in text, fn
                assert "import oracle" not in text, fn
