"""The C-ABI library loads on a machine without a GPU, exports every symbol the header
declares, and fails loudly (no CPU fallback) when asked to decode without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from dint_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "dint_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dint_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = _header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/dint_b200.h but not exported"
    assert sorted(capi.SYMBOLS) == names
    assert L.dint_abi_version() == 7


def test_struct_layouts_match_header():
    assert capi.SEGMENT_DTYPE.itemsize == 32
    assert capi.BLOCK_DTYPE.itemsize == 32
    assert capi.SEGMENT_DTYPE.fields["n"][1] == 16 and capi.BLOCK_DTYPE.fields["block_max"][1] == 24


def test_no_silent_cpu_fallback(golden):
    """without a CUDA device the decode entry points must refuse, not emulate"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    data = golden.bytes(golden.dict_name("single_packed_dint", "docs"))
    with pytest.raises(capi.DintError) as e:
        capi.Dictionary(1, data, device=0)
    assert e.value.code == capi.E_NODEVICE
    d = capi.Dictionary(1, data, device=-1)  # host tables only: fine for the pre-passes
    out = np.zeros(4, dtype=np.uint32)
    used = C.c_size_t(0)
    rc = capi.lib().dint_decode_list_host(d.handle, out.ctypes.data, 16, out.ctypes.data, 1, 256, C.byref(used))
    assert rc == capi.E_NODEVICE
    rc = capi.lib().dint_decode_segments(d.handle, 1, 1, 1, 1, 256, 1, 0, None)
    assert rc == capi.E_NODEVICE
    from dint_b200 import driver
    with pytest.raises(capi.DintError):
        driver.VroomFile("single_packed_dint", b"", data)
    # the round-2 entry points refuse as well: device-side discovery, cold host ctx, device index plan, AND queries
    L = capi.lib()
    h = C.c_void_p()
    buf = np.zeros(64, dtype=np.uint8)
    assert L.dint_vroom_discover(d.handle, buf.ctypes.data, 64, 256, None, 0, 0, None, C.byref(h)) == capi.E_NODEVICE
    assert L.dint_vroom_ctx_create_discover(d.handle, buf.ctypes.data, 64, 256, None, 0, 0, None, C.byref(h)) == capi.E_NODEVICE
    h2 = C.c_void_p()
    assert L.dint_index_discover(d.handle, d.handle, buf.ctypes.data, 64, buf.ctypes.data, 1, 256, None, C.byref(h), C.byref(h2)) \
        == capi.E_NODEVICE
    idx = golden.bytes(golden.index_name("single_packed_dint"))
    a = np.frombuffer(idx, dtype=np.uint8)
    assert L.dint_index_device_plan_create(1, a.ctypes.data, len(a), a.ctypes.data, 0, 256, None, C.byref(h)) == capi.E_NODEVICE
    rc = L.dint_and_queries(buf.ctypes.data, buf.ctypes.data, 1, buf.ctypes.data, buf.ctypes.data, 1, buf.ctypes.data, None)
    assert rc in (capi.E_NODEVICE, -3)  # no device to ask about the pointer: a CUDA error, never a CPU answer
    for kind in (3, 4, 5):  # the dictionary-less block codecs: host memory is refused
        rc = L.dint_decode_index_blocks(kind, buf.ctypes.data, 64, buf.ctypes.data, 1, 256, buf.ctypes.data, buf.ctypes.data, None)
        assert rc in (capi.E_NODEVICE, -3)
    assert L.dint_decode_index_blocks(1, buf.ctypes.data, 64, buf.ctypes.data, 1, 256, buf.ctypes.data, buf.ctypes.data, None) == -1
    assert not buf.any()


def test_product_does_not_reference_the_oracle():
    """the product sources never include, link or import anything under oracle/"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dint_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower().replace("no cpu fallback", ""), os.path.join(dirpath, f)
    assert "oracle" not in open(os.path.join(ROOT, "include", "dint_b200.h")).read().lower()
