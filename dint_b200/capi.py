"""ctypes binding of the C ABI in include/dint_b200.h (libdint_b200.so).

This is the stub a maintainer of a Python host would write; the library itself has no
Python or torch dependency. Every call raises DintError on a non-zero return code —
there is no CPU fallback behind any of them.
"""
import ctypes as C
import os

import numpy as np

from . import _build

DICT_RECT, DICT_PACKED, DICT_MULTI = 0, 1, 2
INDEX_BLOCK_INTERPOLATIVE = 3  # index files only: block_freq_index<interpolative_block>, no dictionaries
INDEX_BLOCK_VBYTE = 4  # block_freq_index<vbyte_block>
INDEX_BLOCK_U32 = 5  # block_freq_index<u32_block>
MODE_RAW, MODE_DOCIDS, MODE_FREQS = 0, 1, 2
E_NODEVICE = -5

KIND_OF_TYPE = {"single_rect_dint": DICT_RECT, "single_packed_dint": DICT_PACKED, "multi_packed_dint": DICT_MULTI,
                "block_interpolative": INDEX_BLOCK_INTERPOLATIVE, "block_vbyte": INDEX_BLOCK_VBYTE, "block_u32": INDEX_BLOCK_U32}

SEGMENT_DTYPE = np.dtype([("in_off", "<u8"), ("out_off", "<u8"), ("n", "<u4"), ("doc_base", "<u4"),
                          ("in_len", "<u4"), ("doc_sum", "<u4")])
BLOCK_DTYPE = np.dtype([("in_off", "<u8"), ("out_off", "<u8"), ("size", "<u4"), ("cur_base", "<u4"),
                        ("block_max", "<u4"), ("reserved", "<u4")])
STRIP_UNIT_DTYPE = np.dtype([("in_off", "<u8"), ("out_off", "<u8"), ("doc_base", "<u4"), ("list_starts", "<u4"),
                             ("in_len", "<u2"), ("n_ints", "<u2"), ("n_strips", "u1"), ("flags", "u1"), ("reserved", "u1", (2,)),
                             ("strip", "<u4", (32,))])
assert SEGMENT_DTYPE.itemsize == 32 and BLOCK_DTYPE.itemsize == 32 and STRIP_UNIT_DTYPE.itemsize == 160

# every symbol include/dint_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "dint_last_error", "dint_abi_version", "dint_kernel_launches", "dint_strip_limits",
    "dint_vroom_plan_num_strip_units", "dint_vroom_plan_codeword_stats", "dint_vroom_plan_strip_units", "dint_decode_strips",
    "dint_dict_create", "dint_dict_create_mapped", "dint_dict_destroy", "dint_dict_kind", "dint_dict_num_rows",
    "dint_dict_device_bytes", "dint_dict_copy",
    "dint_vroom_plan_create", "dint_vroom_plan_destroy", "dint_vroom_plan_num_lists", "dint_vroom_plan_num_ints",
    "dint_vroom_plan_num_segments", "dint_vroom_plan_segments", "dint_vroom_plan_num_units",
    "dint_vroom_plan_unit_first", "dint_vroom_plan_list_offsets", "dint_vroom_plan_list_byte_offsets",
    "dint_index_plan_create", "dint_index_plan_destroy", "dint_index_plan_num_lists", "dint_index_plan_num_docs",
    "dint_index_plan_num_postings", "dint_index_plan_num_blocks", "dint_index_plan_blocks",
    "dint_index_plan_list_offsets", "dint_index_plan_endpoints", "dint_index_plan_layout",
    "dint_decode_segments", "dint_decode_units", "dint_decode_index", "dint_checksum_lists", "dint_decode_list_host",
    "dint_vroom_ctx_create", "dint_vroom_ctx_destroy", "dint_vroom_ctx_device_output", "dint_vroom_decode_host",
    "dint_vroom_decode_resident", "dint_vroom_ctx_sync",
    "dint_vroom_discover", "dint_index_discover", "dint_device_plan_destroy", "dint_device_plan_num_lists",
    "dint_device_plan_num_ints", "dint_device_plan_num_units", "dint_device_plan_num_chains",
    "dint_device_plan_num_sync_found", "dint_device_plan_num_skimmed", "dint_device_plan_retries", "dint_device_plan_discover_ms", "dint_device_plan_phase_ms",
    "dint_device_plan_units", "dint_device_plan_list_offsets", "dint_device_plan_list_byte_offsets",
    "dint_decode_index_planned", "dint_vroom_ctx_create_discover", "dint_vroom_ctx_num_lists",
    "dint_vroom_ctx_num_ints", "dint_vroom_ctx_device_plan", "dint_vroom_ctx_results", "dint_and_queries",
    "dint_index_device_plan_create", "dint_device_index_destroy", "dint_device_index_num_lists", "dint_device_index_num_docs",
    "dint_device_index_num_postings", "dint_device_index_num_blocks", "dint_device_index_build_ms", "dint_device_index_blocks",
    "dint_device_index_list_offsets", "dint_device_index_endpoints", "dint_device_index_layout",
    "dint_decode_index_interpolative", "dint_decode_index_blocks",
]


class DintError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dint_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.environ.get("DINT_B200_LIB") or _build.build()  # override: A/B builds of the same ABI (profiles/tools)
        L = C.CDLL(path)
        vp, sz, u32, u64, i32 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_int
        L.dint_last_error.restype = C.c_char_p
        L.dint_abi_version.restype = i32
        L.dint_kernel_launches.restype = u64
        L.dint_dict_create.argtypes = [i32, vp, sz, i32, C.POINTER(vp)]
        L.dint_dict_create_mapped.argtypes = [i32, vp, sz, i32, C.POINTER(vp), C.POINTER(sz)]
        L.dint_dict_destroy.argtypes = [vp]
        L.dint_dict_destroy.restype = None
        L.dint_dict_kind.argtypes = [vp]
        L.dint_dict_num_rows.argtypes = [vp]
        L.dint_dict_num_rows.restype = u32
        L.dint_dict_device_bytes.argtypes = [vp]
        L.dint_dict_device_bytes.restype = sz
        L.dint_dict_copy.argtypes = [vp, u32, u32, vp, C.POINTER(u32)]
        L.dint_vroom_plan_create.argtypes = [vp, vp, sz, u32, u32, C.POINTER(vp)]
        L.dint_vroom_plan_destroy.argtypes = [vp]
        L.dint_vroom_plan_destroy.restype = None
        L.dint_vroom_plan_codeword_stats.argtypes = [vp, C.POINTER(u64 * 16)]
        L.dint_vroom_plan_codeword_stats.restype = None
        L.dint_strip_limits.argtypes = [C.POINTER(u32)] * 3
        L.dint_strip_limits.restype = None
        L.dint_decode_strips.argtypes = [vp, vp, sz, vp, sz, vp, i32, vp]
        for f in ("num_lists", "num_ints", "num_segments", "num_units", "num_strip_units"):
            getattr(L, "dint_vroom_plan_" + f).argtypes = [vp]
            getattr(L, "dint_vroom_plan_" + f).restype = u64
        for f in ("segments", "unit_first", "list_offsets", "list_byte_offsets", "strip_units"):
            getattr(L, "dint_vroom_plan_" + f).argtypes = [vp]
            getattr(L, "dint_vroom_plan_" + f).restype = vp
        L.dint_index_plan_create.argtypes = [i32, vp, sz, u32, C.POINTER(vp)]
        L.dint_index_plan_destroy.argtypes = [vp]
        L.dint_index_plan_destroy.restype = None
        for f in ("num_lists", "num_docs", "num_postings", "num_blocks"):
            getattr(L, "dint_index_plan_" + f).argtypes = [vp]
            getattr(L, "dint_index_plan_" + f).restype = u64
        for f in ("blocks", "list_offsets", "endpoints"):
            getattr(L, "dint_index_plan_" + f).argtypes = [vp]
            getattr(L, "dint_index_plan_" + f).restype = vp
        L.dint_index_plan_layout.argtypes = [vp] + [C.POINTER(u64)] * 5
        L.dint_decode_segments.argtypes = [vp, vp, sz, vp, sz, u32, vp, i32, vp]
        L.dint_decode_units.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, i32, vp]
        L.dint_decode_index.argtypes = [vp, vp, vp, sz, vp, sz, u32, vp, vp, vp]
        L.dint_checksum_lists.argtypes = [vp, vp, sz, vp, vp]
        L.dint_decode_list_host.argtypes = [vp, vp, sz, vp, u32, u32, C.POINTER(sz)]
        L.dint_vroom_ctx_create.argtypes = [vp, vp, u32, C.POINTER(vp)]
        L.dint_vroom_ctx_destroy.argtypes = [vp]
        L.dint_vroom_ctx_destroy.restype = None
        L.dint_vroom_ctx_device_output.argtypes = [vp]
        L.dint_vroom_ctx_device_output.restype = vp
        L.dint_vroom_decode_host.argtypes = [vp, vp, sz, i32, vp, vp, vp]
        L.dint_vroom_decode_resident.argtypes = [vp, i32, vp]
        L.dint_vroom_ctx_sync.argtypes = [vp, vp]
        L.dint_vroom_discover.argtypes = [vp, vp, sz, u32, vp, sz, u64, vp, C.POINTER(vp)]
        L.dint_index_discover.argtypes = [vp, vp, vp, sz, vp, sz, u32, vp, C.POINTER(vp), C.POINTER(vp)]
        L.dint_device_plan_destroy.argtypes = [vp]
        L.dint_device_plan_destroy.restype = None
        for f in ("num_lists", "num_ints", "num_units", "num_chains", "num_sync_found", "num_skimmed"):
            getattr(L, "dint_device_plan_" + f).argtypes = [vp]
            getattr(L, "dint_device_plan_" + f).restype = u64
        L.dint_device_plan_retries.argtypes = [vp]
        L.dint_device_plan_retries.restype = u32
        L.dint_device_plan_discover_ms.argtypes = [vp]
        L.dint_device_plan_discover_ms.restype = C.c_double
        L.dint_device_plan_phase_ms.argtypes = [vp, i32]
        L.dint_device_plan_phase_ms.restype = C.c_double
        for f in ("units", "list_offsets", "list_byte_offsets"):
            getattr(L, "dint_device_plan_" + f).argtypes = [vp]
            getattr(L, "dint_device_plan_" + f).restype = vp
        L.dint_decode_index_planned.argtypes = [vp, vp, vp, sz, vp, vp, vp, sz, u32, vp, vp, vp]
        L.dint_vroom_ctx_create_discover.argtypes = [vp, vp, sz, u32, vp, sz, u64, vp, C.POINTER(vp)]
        L.dint_vroom_ctx_num_lists.argtypes = [vp]
        L.dint_vroom_ctx_num_lists.restype = u64
        L.dint_vroom_ctx_num_ints.argtypes = [vp]
        L.dint_vroom_ctx_num_ints.restype = u64
        L.dint_vroom_ctx_device_plan.argtypes = [vp]
        L.dint_vroom_ctx_device_plan.restype = vp
        L.dint_vroom_ctx_results.argtypes = [vp, vp, vp, vp]
        L.dint_and_queries.argtypes = [vp, vp, sz, vp, vp, sz, vp, vp]
        L.dint_index_device_plan_create.argtypes = [i32, vp, sz, vp, i32, u32, vp, C.POINTER(vp)]
        L.dint_device_index_destroy.argtypes = [vp]
        L.dint_device_index_destroy.restype = None
        for f in ("num_lists", "num_docs", "num_postings", "num_blocks"):
            getattr(L, "dint_device_index_" + f).argtypes = [vp]
            getattr(L, "dint_device_index_" + f).restype = u64
        L.dint_device_index_build_ms.argtypes = [vp]
        L.dint_device_index_build_ms.restype = C.c_double
        for f in ("blocks", "list_offsets", "endpoints"):
            getattr(L, "dint_device_index_" + f).argtypes = [vp]
            getattr(L, "dint_device_index_" + f).restype = vp
        L.dint_device_index_layout.argtypes = [vp] + [C.POINTER(u64)] * 5
        L.dint_decode_index_interpolative.argtypes = [vp, sz, vp, sz, u32, vp, vp, vp]
        L.dint_decode_index_blocks.argtypes = [i32, vp, sz, vp, sz, u32, vp, vp, vp]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise DintError(rc, lib().dint_last_error().decode(errors="replace"))


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _view(ptr, count, dtype):
    """copy `count` items of `dtype` out of library-owned memory"""
    if count == 0:
        return np.zeros(0, dtype=dtype)
    nbytes = count * np.dtype(dtype).itemsize
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count).copy()


class Dictionary:
    """A reference dictionary file, parsed and (device >= 0) resident in HBM."""

    def __init__(self, kind, data, device=-1, mapped=False):
        self.kind = kind
        self.device = device
        a = np.frombuffer(data, dtype=np.uint8)
        h = C.c_void_p()
        self.consumed = len(a)
        if mapped:
            used = C.c_size_t(0)
            check(lib().dint_dict_create_mapped(kind, _np_ptr(a), len(a), device, C.byref(h), C.byref(used)))
            self.consumed = used.value
        else:
            check(lib().dint_dict_create(kind, _np_ptr(a), len(a), device, C.byref(h)))
        self.handle = h

    @classmethod
    def from_file(cls, kind, path, device=-1):
        with open(path, "rb") as f:
            return cls(kind, f.read(), device)

    @property
    def num_rows(self):
        return lib().dint_dict_num_rows(self.handle)

    @property
    def device_bytes(self):
        return lib().dint_dict_device_bytes(self.handle)

    def copy(self, cls_, i):
        out = np.zeros(16, dtype=np.uint32)
        ln = C.c_uint32(0)
        check(lib().dint_dict_copy(self.handle, cls_, i, _np_ptr(out), C.byref(ln)))
        return ln.value, out

    def close(self):
        if self.handle:
            lib().dint_dict_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class VroomPlan:
    """Segment table of a vroom .bin (host pre-pass)."""

    def __init__(self, dictionary, data, block_size=256, chunk_ints=0):
        a = np.frombuffer(data, dtype=np.uint8)
        h = C.c_void_p()
        check(lib().dint_vroom_plan_create(dictionary.handle, _np_ptr(a), len(a), block_size, chunk_ints, C.byref(h)))
        L = lib()
        self.n_lists = L.dint_vroom_plan_num_lists(h)
        self.n_ints = L.dint_vroom_plan_num_ints(h)
        self.n_segments = L.dint_vroom_plan_num_segments(h)
        self.segments = _view(L.dint_vroom_plan_segments(h), self.n_segments, SEGMENT_DTYPE)
        self.n_units = L.dint_vroom_plan_num_units(h)
        self.unit_first = _view(L.dint_vroom_plan_unit_first(h), self.n_units + 1, np.uint32)
        self.n_strip_units = L.dint_vroom_plan_num_strip_units(h)
        self.strip_units = _view(L.dint_vroom_plan_strip_units(h), self.n_strip_units, STRIP_UNIT_DTYPE)
        self.list_offsets = _view(L.dint_vroom_plan_list_offsets(h), self.n_lists + 1, np.uint64)
        self.list_byte_offsets = _view(L.dint_vroom_plan_list_byte_offsets(h), self.n_lists + 1, np.uint64)
        st = (C.c_uint64 * 16)()
        L.dint_vroom_plan_codeword_stats(h, C.byref(st))
        self.codeword_stats = dict(zip(("exc16", "exc32", "runs", "len1", "len2", "len4", "len8", "len16", "u8_codewords",
                                        "u16_codewords", "run_ints", "unit_trips", "strips"), (int(x) for x in st)))
        self.block_size = block_size
        self.in_bytes = len(a)
        self.handle = h
        self.dictionary = dictionary

    def close(self):
        if getattr(self, "handle", None):
            lib().dint_vroom_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


NO_SYNC_SEARCH = 0xFFFFFFFFFFFFFFFF  # sync_stride: use only the caller's list starts


def _sync_args(sync_offsets):
    if sync_offsets is None or len(sync_offsets) == 0:
        return None, 0, None
    a = np.ascontiguousarray(sync_offsets, dtype=np.uint64)
    return _np_ptr(a), len(a), a


class DevicePlan:
    """Strip table (+ list offsets) the GPU found itself; everything stays in HBM (dint_device_plan)."""

    def __init__(self, handle, owned=True):
        L = lib()
        self.handle = handle
        self.owned = owned
        self.n_lists = L.dint_device_plan_num_lists(handle)
        self.n_ints = L.dint_device_plan_num_ints(handle)
        self.n_units = L.dint_device_plan_num_units(handle)
        self.n_chains = L.dint_device_plan_num_chains(handle)
        self.n_sync_found = L.dint_device_plan_num_sync_found(handle)
        self.n_skimmed = L.dint_device_plan_num_skimmed(handle)
        self.retries = L.dint_device_plan_retries(handle)
        self.discover_ms = L.dint_device_plan_discover_ms(handle)
        self.phase_ms = [L.dint_device_plan_phase_ms(handle, i) for i in range(3)]
        self.units_ptr = L.dint_device_plan_units(handle)
        self.list_offsets_ptr = L.dint_device_plan_list_offsets(handle)
        self.list_byte_offsets_ptr = L.dint_device_plan_list_byte_offsets(handle)

    @classmethod
    def discover_vroom(cls, dictionary, d_in_ptr, in_bytes, block_size=256, sync_offsets=None, sync_stride=0, stream=0):
        """dint_vroom_discover: restart points of a vroom .bin that is already in HBM (no host plan)."""
        h = C.c_void_p()
        p, n, keep = _sync_args(sync_offsets)
        check(lib().dint_vroom_discover(dictionary.handle, d_in_ptr, in_bytes, block_size, p, n, sync_stride, stream,
                                        C.byref(h)))
        return cls(h)

    @classmethod
    def discover_index(cls, docs_dict, freqs_dict, d_lists_ptr, lists_bytes, d_blocks_ptr, n_blocks, block_size=256, stream=0):
        """dint_index_discover: (docs plan, freqs plan) for the full blocks of a frozen index"""
        hd, hf = C.c_void_p(), C.c_void_p()
        check(lib().dint_index_discover(docs_dict.handle, freqs_dict.handle, d_lists_ptr, lists_bytes, d_blocks_ptr, n_blocks,
                                        block_size, stream, C.byref(hd), C.byref(hf)))
        return cls(hd), cls(hf)

    def close(self):
        if getattr(self, "handle", None) and self.owned:
            lib().dint_device_plan_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class VroomHostDecoder:
    """Host-buffer batch decode of whole vroom .bin files that share one plan
    (vroom_env/decode.cpp:139-150 as one call): H2D, decode, checksums / output D2H."""

    def __init__(self, plan, block_size=256):
        self.plan = plan
        h = C.c_void_p()
        check(lib().dint_vroom_ctx_create(plan.dictionary.handle, plan.handle, block_size, C.byref(h)))
        self.handle = h

    @classmethod
    def discover(cls, dictionary, h_bin_ptr, nbytes, block_size=256, sync_offsets=None, sync_stride=0, stream=0):
        """cold path: no host plan — H2D of the file + device-side discovery (dint_vroom_ctx_create_discover)"""
        self = cls.__new__(cls)
        self.plan = None
        self.dictionary = dictionary
        h = C.c_void_p()
        p, n, keep = _sync_args(sync_offsets)
        check(lib().dint_vroom_ctx_create_discover(dictionary.handle, h_bin_ptr, nbytes, block_size, p, n, sync_stride,
                                                   stream, C.byref(h)))
        self.handle = h
        self.device_plan = DevicePlan(lib().dint_vroom_ctx_device_plan(h), owned=False)
        return self

    @property
    def n_lists(self):
        return lib().dint_vroom_ctx_num_lists(self.handle)

    @property
    def n_ints(self):
        return lib().dint_vroom_ctx_num_ints(self.handle)

    def results(self, h_out_ptr=None, h_sums_ptr=None, stream=0):
        check(lib().dint_vroom_ctx_results(self.handle, h_out_ptr, h_sums_ptr, stream))

    def decode(self, h_bin_ptr, nbytes, mode=MODE_RAW, h_out_ptr=None, h_sums_ptr=None, stream=0):
        check(lib().dint_vroom_decode_host(self.handle, h_bin_ptr, nbytes, mode, h_out_ptr, h_sums_ptr, stream))

    def decode_resident(self, mode=MODE_RAW, stream=0):
        check(lib().dint_vroom_decode_resident(self.handle, mode, stream))

    def sync(self, stream=0):
        check(lib().dint_vroom_ctx_sync(self.handle, stream))

    @property
    def device_output(self):
        return lib().dint_vroom_ctx_device_output(self.handle)

    def close(self):
        if getattr(self, "handle", None):
            lib().dint_vroom_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class IndexPlan:
    """Block table of a frozen dict_freq_index .bin (host pre-pass)."""

    def __init__(self, kind, data, block_size=256):
        a = np.frombuffer(data, dtype=np.uint8)
        h = C.c_void_p()
        check(lib().dint_index_plan_create(kind, _np_ptr(a), len(a), block_size, C.byref(h)))
        L = lib()
        self.kind = kind
        self.n_lists = L.dint_index_plan_num_lists(h)
        self.num_docs = L.dint_index_plan_num_docs(h)
        self.n_postings = L.dint_index_plan_num_postings(h)
        self.n_blocks = L.dint_index_plan_num_blocks(h)
        self.blocks = _view(L.dint_index_plan_blocks(h), self.n_blocks, BLOCK_DTYPE)
        self.list_offsets = _view(L.dint_index_plan_list_offsets(h), self.n_lists + 1, np.uint64)
        self.endpoints = _view(L.dint_index_plan_endpoints(h), self.n_lists + 1, np.uint64)
        v = [C.c_uint64(0) for _ in range(5)]
        check(L.dint_index_plan_layout(h, *[C.byref(x) for x in v]))
        self.lists_off, self.lists_len, self.docs_dict_off, self.freqs_dict_off, self.end_off = [x.value for x in v]
        self.block_size = block_size
        L.dint_index_plan_destroy(h)


class DeviceIndexPlan:
    """Block table of a frozen dict_freq_index built ON THE DEVICE (dint_index_device_plan_create): the host reads
    only the fixed-size header of the file; endpoints, list offsets and blocks stay in HBM."""

    def __init__(self, kind, data, d_bin_ptr, device=0, block_size=256, stream=0):
        a = np.frombuffer(data, dtype=np.uint8)
        h = C.c_void_p()
        check(lib().dint_index_device_plan_create(kind, _np_ptr(a), len(a), d_bin_ptr, device, block_size, stream, C.byref(h)))
        L = lib()
        self.handle = h
        self.kind = kind
        self.block_size = block_size
        self.n_lists = L.dint_device_index_num_lists(h)
        self.num_docs = L.dint_device_index_num_docs(h)
        self.n_postings = L.dint_device_index_num_postings(h)
        self.n_blocks = L.dint_device_index_num_blocks(h)
        self.build_ms = L.dint_device_index_build_ms(h)
        self.blocks_ptr = L.dint_device_index_blocks(h)
        self.list_offsets_ptr = L.dint_device_index_list_offsets(h)
        self.endpoints_ptr = L.dint_device_index_endpoints(h)
        v = [C.c_uint64(0) for _ in range(5)]
        check(L.dint_device_index_layout(h, *[C.byref(x) for x in v]))
        self.lists_off, self.lists_len, self.docs_dict_off, self.freqs_dict_off, self.end_off = [x.value for x in v]

    def close(self):
        if getattr(self, "handle", None):
            lib().dint_device_index_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def decode_segments(dictionary, d_in_ptr, in_bytes, d_segs_ptr, n_segs, d_out_ptr, mode, stream=0, block_size=256):
    check(lib().dint_decode_segments(dictionary.handle, d_in_ptr, in_bytes, d_segs_ptr, n_segs, block_size, d_out_ptr,
                                     mode, stream))


def decode_units(dictionary, d_in_ptr, in_bytes, d_segs_ptr, n_segs, d_unit_first_ptr, n_units, d_out_ptr, mode, stream=0):
    check(lib().dint_decode_units(dictionary.handle, d_in_ptr, in_bytes, d_segs_ptr, n_segs, d_unit_first_ptr, n_units,
                                  d_out_ptr, mode, stream))


def decode_strips(dictionary, d_in_ptr, in_bytes, d_units_ptr, n_units, d_out_ptr, mode, stream=0):
    check(lib().dint_decode_strips(dictionary.handle, d_in_ptr, in_bytes, d_units_ptr, n_units, d_out_ptr, mode, stream))


def strip_limits():
    a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
    lib().dint_strip_limits(C.byref(a), C.byref(b), C.byref(c))
    return {"unit_ints": a.value, "unit_bytes": b.value, "strip_codewords": c.value}


def decode_index(docs_dict, freqs_dict, d_lists_ptr, lists_bytes, d_blocks_ptr, n_blocks, block_size, d_docids_ptr,
                 d_freqs_ptr, stream=0):
    check(lib().dint_decode_index(docs_dict.handle, freqs_dict.handle, d_lists_ptr, lists_bytes, d_blocks_ptr, n_blocks,
                                  block_size, d_docids_ptr, d_freqs_ptr, stream))


def decode_index_planned(docs_dict, freqs_dict, d_lists_ptr, lists_bytes, docs_plan, freqs_plan, d_blocks_ptr, n_blocks,
                         block_size, d_docids_ptr, d_freqs_ptr, stream=0):
    check(lib().dint_decode_index_planned(docs_dict.handle, freqs_dict.handle, d_lists_ptr, lists_bytes, docs_plan.handle,
                                          freqs_plan.handle, d_blocks_ptr, n_blocks, block_size, d_docids_ptr, d_freqs_ptr,
                                          stream))


def decode_index_blocks(kind, d_lists_ptr, lists_bytes, d_blocks_ptr, n_blocks, block_size, d_docids_ptr, d_freqs_ptr, stream=0):
    """block_freq_index<interpolative_block | vbyte_block | u32_block>: every block of the table by that codec"""
    check(lib().dint_decode_index_blocks(kind, d_lists_ptr, lists_bytes, d_blocks_ptr, n_blocks, block_size, d_docids_ptr,
                                         d_freqs_ptr, stream))


def decode_index_interpolative(d_lists_ptr, lists_bytes, d_blocks_ptr, n_blocks, block_size, d_docids_ptr, d_freqs_ptr, stream=0):
    """block_freq_index<interpolative_block>: every block of the table is binary-interpolative coded"""
    check(lib().dint_decode_index_interpolative(d_lists_ptr, lists_bytes, d_blocks_ptr, n_blocks, block_size, d_docids_ptr,
                                                d_freqs_ptr, stream))


def and_queries(d_docids_ptr, d_list_offsets_ptr, n_lists, d_query_offsets_ptr, d_query_terms_ptr, n_queries, d_counts_ptr, stream=0):
    """ds2i::and_query over decoded docID lists in HBM: counts[q] = documents holding every term of query q"""
    check(lib().dint_and_queries(d_docids_ptr, d_list_offsets_ptr, n_lists, d_query_offsets_ptr, d_query_terms_ptr, n_queries,
                                 d_counts_ptr, stream))


def checksum_lists(d_values_ptr, d_list_offsets_ptr, n_lists, d_sums_ptr, stream=0):
    check(lib().dint_checksum_lists(d_values_ptr, d_list_offsets_ptr, n_lists, d_sums_ptr, stream))


def decode_list_host(dictionary, data, n, block_size=256):
    """One Decoder::decode(dict, in, out, universe, n) call on host buffers. Returns (values, consumed)."""
    a = np.frombuffer(data, dtype=np.uint8)
    out = np.zeros(max(n, 1), dtype=np.uint32)
    used = C.c_size_t(0)
    check(lib().dint_decode_list_host(dictionary.handle, _np_ptr(a), len(a), _np_ptr(out), n, block_size, C.byref(used)))
    return out[:n], used.value


def kernel_launches():
    return lib().dint_kernel_launches()
