"""ctypes bindings of the CPU checkers (TEST INFRASTRUCTURE).

* `Oracle`    -> oracle/libdint_oracle.so, the plain-C restatement (always available:
                 built on demand with gcc).
* `Reference` -> oracle/_ref/libdint_ref.so, the UNMODIFIED reference compiled from
                 /root/reference (present only if `make -C oracle ref` ran where the
                 reference is mounted; travels to the GPU box as a prebuilt file).
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
KINDS = {"single_rect_dint": 0, "single_packed_dint": 1, "multi_packed_dint": 2}
DICT_NAMES = {"single_rect_dint": "rectangular", "single_packed_dint": "single_packed",
              "multi_packed_dint": "multi_packed"}


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class IndexInfo(C.Structure):
    _fields_ = [("n_lists", C.c_uint64), ("num_docs", C.c_uint64), ("endpoints_bits", C.c_uint64),
                ("endpoints_words_off", C.c_size_t), ("lists_off", C.c_size_t), ("lists_len", C.c_size_t),
                ("docs_dict_off", C.c_size_t), ("freqs_dict_off", C.c_size_t), ("end_off", C.c_size_t)]


def build_oracle():
    so = os.path.join(ORACLE_DIR, "libdint_oracle.so")
    src = os.path.join(ORACLE_DIR, "dint_oracle.c")
    hdr = os.path.join(ORACLE_DIR, "dint_oracle.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, stdout=subprocess.DEVNULL)
    return so


class Oracle:
    def __init__(self):
        L = self.lib = C.CDLL(build_oracle())
        L.oracle_dict_parse.restype = C.c_void_p
        L.oracle_dict_parse.argtypes = [C.c_int, C.c_void_p, C.c_size_t]
        L.oracle_dict_parse_mapped.restype = C.c_void_p
        L.oracle_dict_parse_mapped.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.oracle_dict_free.argtypes = [C.c_void_p]
        L.oracle_dict_copy.restype = C.c_uint32
        L.oracle_dict_copy.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.oracle_vbyte_decode.restype = C.c_size_t
        L.oracle_vbyte_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.oracle_interpolative_decode.restype = C.c_size_t
        L.oracle_interpolative_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_size_t]
        L.oracle_vroom_decode.restype = C.c_size_t
        L.oracle_vroom_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32]
        L.oracle_block_decode.restype = C.c_size_t
        L.oracle_block_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_size_t, C.c_uint32]
        L.oracle_vroom_decode_file.restype = C.c_long
        L.oracle_vroom_decode_file.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p,
                                               C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
        L.oracle_index_parse.restype = C.c_int
        L.oracle_index_parse.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.POINTER(IndexInfo)]
        L.oracle_index_endpoints.restype = C.c_int
        L.oracle_index_endpoints.argtypes = [C.c_void_p, C.POINTER(IndexInfo), C.c_void_p]
        L.oracle_index_decode.restype = C.c_long
        L.oracle_index_decode.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p,
                                          C.c_size_t, C.c_void_p, C.c_size_t]

    def dict_parse(self, kind, file_bytes):
        a = np.frombuffer(file_bytes, dtype=np.uint8)
        h = self.lib.oracle_dict_parse(kind, _ptr(a), len(a))
        if not h:
            raise ValueError("oracle: malformed dictionary file")
        return h

    def dict_free(self, h):
        self.lib.oracle_dict_free(h)

    def dict_copy(self, h, cls, i):
        out = np.zeros(16, dtype=np.uint32)
        n = self.lib.oracle_dict_copy(h, cls, i, _ptr(out))
        return n, out

    def vroom_decode_list(self, h, data, n, block_size=256):
        """decode one list from `data` (uint8 array); returns (values[n], bytes_consumed)"""
        out = np.zeros(n + 512 + 16, dtype=np.uint32)
        used = self.lib.oracle_vroom_decode(h, _ptr(data), _ptr(out), n, block_size)
        return out[:n].copy(), used

    def block_decode(self, h, data, sum_of_values, n, block_size=256):
        out = np.zeros(2 * block_size + 512, dtype=np.uint32)
        used = self.lib.oracle_block_decode(h, _ptr(data), _ptr(out), sum_of_values & 0xFFFFFFFF, n, block_size)
        return out[:n].copy(), used

    def vroom_decode_file(self, h, bin_bytes, n_ints_cap, n_lists_cap, block_size=256):
        a = np.frombuffer(bin_bytes, dtype=np.uint8)
        out = np.zeros(n_ints_cap + 16, dtype=np.uint32)
        ln = np.zeros(n_lists_cap, dtype=np.uint32)
        tot = C.c_uint64(0)
        nl = self.lib.oracle_vroom_decode_file(h, _ptr(a), len(a), block_size, _ptr(out), n_ints_cap, _ptr(ln),
                                               n_lists_cap, C.byref(tot))
        if nl < 0:
            raise ValueError("oracle: output capacity exceeded")
        return out[:tot.value], ln[:nl]

    def index_parse(self, kind, bin_bytes):
        a = np.frombuffer(bin_bytes, dtype=np.uint8)
        info = IndexInfo()
        rc = self.lib.oracle_index_parse(kind, _ptr(a), len(a), C.byref(info))
        if rc != 0:
            raise ValueError(f"oracle: index parse failed rc={rc}")
        return info

    def index_endpoints(self, kind, bin_bytes):
        a = np.frombuffer(bin_bytes, dtype=np.uint8)
        info = self.index_parse(kind, bin_bytes)
        ep = np.zeros(info.n_lists, dtype=np.uint64)
        if self.lib.oracle_index_endpoints(_ptr(a), C.byref(info), _ptr(ep)) != 0:
            raise ValueError("oracle: EF decode failed")
        return ep, info

    def index_decode(self, kind, bin_bytes, n_cap, block_size=256):
        a = np.frombuffer(bin_bytes, dtype=np.uint8)
        info = self.index_parse(kind, bin_bytes)
        d = np.zeros(n_cap, dtype=np.uint32)
        f = np.zeros(n_cap, dtype=np.uint32)
        ln = np.zeros(info.n_lists, dtype=np.uint32)
        n = self.lib.oracle_index_decode(kind, _ptr(a), len(a), block_size, _ptr(d), _ptr(f), n_cap, _ptr(ln),
                                         len(ln))
        if n < 0:
            raise ValueError("oracle: index decode failed")
        return d[:n], f[:n], ln


def reference_so(block_size=256):
    sub = "_ref" if block_size == 256 else os.path.join("_ref", f"bs{block_size}")
    p = os.path.join(ORACLE_DIR, sub, "libdint_ref.so")
    return p if os.path.exists(p) else None


class Reference:
    """The real reference decoders (oracle/ref_drivers/ref_capi.cpp)."""

    def __init__(self, block_size=256):
        so = reference_so(block_size)
        if so is None:
            raise FileNotFoundError("oracle/_ref/libdint_ref.so not built (make -C oracle ref)")
        L = self.lib = C.CDLL(so)
        L.ref_dict_load.restype = C.c_void_p
        L.ref_dict_load.argtypes = [C.c_int, C.c_char_p]
        L.ref_dict_free.argtypes = [C.c_void_p]
        L.ref_vroom_decode.restype = C.c_size_t
        L.ref_vroom_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.ref_block_decode.restype = C.c_size_t
        L.ref_block_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.ref_vbyte_decode.restype = C.c_size_t
        L.ref_vbyte_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.ref_vroom_decode_file.restype = C.c_long
        L.ref_vroom_decode_file.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                            C.POINTER(C.c_uint64)]
        L.ref_vroom_time_file.restype = C.c_double
        L.ref_vroom_time_file.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_uint64), C.c_void_p]
        assert L.ref_block_size() == block_size

    def dict_load(self, kind, path):
        h = self.lib.ref_dict_load(kind, path.encode())
        if not h:
            raise ValueError("reference: cannot load dictionary " + path)
        return h

    def vroom_decode_list(self, h, data, n):
        out = np.zeros(n + 512 + 16, dtype=np.uint32)
        used = self.lib.ref_vroom_decode(h, _ptr(data), _ptr(out), n)
        return out[:n].copy(), used

    def block_decode(self, h, data, sum_of_values, n):
        out = np.zeros(1024, dtype=np.uint32)
        used = self.lib.ref_block_decode(h, _ptr(data), _ptr(out), sum_of_values & 0xFFFFFFFF, n)
        return out[:n].copy(), used

    def vroom_decode_file(self, h, bin_bytes, n_ints_cap):
        a = np.frombuffer(bin_bytes, dtype=np.uint8)
        out = np.zeros(n_ints_cap + 16, dtype=np.uint32)
        tot = C.c_uint64(0)
        nl = self.lib.ref_vroom_decode_file(h, _ptr(a), len(a), _ptr(out), n_ints_cap, C.byref(tot))
        if nl < 0:
            raise ValueError("reference: output capacity exceeded")
        return out[:tot.value], nl

    def vroom_time(self, h, a, repeat=1):
        tot = C.c_uint64(0)
        secs = self.lib.ref_vroom_time_file(h, _ptr(a), len(a), repeat, C.byref(tot), None)
        return secs, tot.value
