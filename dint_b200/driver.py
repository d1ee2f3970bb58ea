"""Host-side mirror of the reference's decode drivers, on top of the C ABI.

  * `VroomFile`  + `decode_dint`   <->  vroom_env/decode.cpp:95-155 (decode_dint<Decoder, Dictionary>)
                                       and vroom_env/statistics.hpp:7-34 (same JSON keys)
  * `IndexFile`  + `scan_index`    <->  succinct::mapper::map(dict_freq_index) + a full
                                       document_enumerator scan (dict_freq_index.hpp:174-181,
                                       dict_posting_list.hpp:113-124, 164-169)

PyTorch is used only as plumbing: device buffers, streams, events. All decoding happens in
libdint_b200.so; if the library or a CUDA device is missing these calls raise.
"""
import json

import numpy as np
import torch

from . import capi, shard

TYPES = ("single_rect_dint", "single_packed_dint", "multi_packed_dint")
INDEX_TYPES = TYPES + ("block_interpolative", "block_vbyte", "block_u32")  # frozen index files: + block_freq_index<...>


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise capi.DintError(capi.E_NODEVICE, "no CUDA device: dint_b200 has no CPU fallback")
    return torch.device("cuda", device)


def _to_device_bytes(data, dev, pinned=False):
    if isinstance(data, torch.Tensor):
        t = data
    elif len(data) == 0:
        t = torch.empty(0, dtype=torch.uint8)
    else:
        t = torch.frombuffer(bytearray(data), dtype=torch.uint8)
    if pinned:
        t = t.pin_memory()
    return t.to(dev, non_blocking=pinned)


def _structured_to_device(arr, dev):
    raw = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1))
    return raw.to(dev)


class _DeviceArray:
    """zero-copy torch view of library-owned device memory (__cuda_array_interface__)"""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_view(ptr, count, typestr, dev):
    if not count:
        return torch.empty(0, dtype={"<i8": torch.int64, "|u1": torch.uint8}[typestr], device=dev)
    return torch.as_tensor(_DeviceArray(ptr, count, typestr), device=dev)


class VroomFile:
    """A vroom-encoded .bin + its dictionary, resident on one GPU, ready for batch decode.

    plan="host": the restart points come from the serial host walk (dint_vroom_plan_create);
    plan="device": no host plan at all — the GPU finds them in the bytes it already holds
    (dint_vroom_discover; sync_offsets / sync_stride as documented in include/dint_b200.h)."""

    def __init__(self, type_, bin_bytes, dict_bytes, device=0, block_size=256, chunk_ints=0, engine="strips", plan="host",
                 sync_offsets=None, sync_stride=0):
        if type_ not in capi.KIND_OF_TYPE:
            raise ValueError(f"unknown type '{type_}'")  # decode.cpp:259 logs "ERROR: unknown type"
        self.type = type_
        self.dev = _require_cuda(device)
        self.block_size = block_size
        self.dict = capi.Dictionary(capi.KIND_OF_TYPE[type_], dict_bytes, device=device)
        if plan == "device":
            if engine != "strips":
                raise ValueError("plan='device' produces strip tables: engine must be 'strips'")
            self.engine = engine
            self.plan = None
            self.in_bytes = len(bin_bytes)
            self.d_in = _to_device_bytes(bin_bytes, self.dev)
            with torch.cuda.device(self.dev):
                st = torch.cuda.current_stream(self.dev).cuda_stream
                self.dplan = capi.DevicePlan.discover_vroom(self.dict, self.d_in.data_ptr(), self.in_bytes, block_size,
                                                            sync_offsets, sync_stride, stream=st)
            self.n_ints = int(self.dplan.n_ints)
            self.n_lists = int(self.dplan.n_lists)
            self.n_strip_units = int(self.dplan.n_units)
            self.units_ptr = self.dplan.units_ptr
            self.d_list_offsets = device_view(self.dplan.list_offsets_ptr, self.n_lists + 1, "<i8", self.dev)
            self.d_list_byte_offsets = device_view(self.dplan.list_byte_offsets_ptr, self.n_lists + 1, "<i8", self.dev)
            return
        if plan != "host":
            raise ValueError(f"unknown plan '{plan}'")
        self.dplan = None
        self.in_bytes = len(bin_bytes)
        self.plan = capi.VroomPlan(self.dict, bin_bytes, block_size=block_size, chunk_ints=chunk_ints)
        # engine: "strips" = lane-sequential decoder over the plan's strip table (dint_decode_strips, the
        # fast path); "units" = warp-per-work-unit tile decoder over the segment table (dint_decode_units)
        if engine not in ("strips", "units"):
            raise ValueError(f"unknown engine '{engine}'")
        self.engine = engine
        self.d_in = _to_device_bytes(bin_bytes, self.dev)
        self.d_strip_units = _structured_to_device(self.plan.strip_units, self.dev)
        self.d_segs = _structured_to_device(self.plan.segments, self.dev)
        self.d_unit_first = torch.from_numpy(self.plan.unit_first.astype(np.int64)).to(torch.int32).to(self.dev)
        self.d_list_offsets = torch.from_numpy(self.plan.list_offsets.astype(np.int64)).to(self.dev)
        self.n_ints = int(self.plan.n_ints)
        self.n_lists = int(self.plan.n_lists)
        self.n_strip_units = int(self.plan.n_strip_units)
        self.units_ptr = self.d_strip_units.data_ptr()

    def decode(self, out=None, mode=capi.MODE_RAW, stream=None):
        """Decode every list of the file into one packed int32 tensor of n_ints values."""
        if out is None:
            out = torch.empty(self.n_ints, dtype=torch.int32, device=self.dev)
        st = (stream or torch.cuda.current_stream(self.dev)).cuda_stream
        if self.engine == "strips":
            capi.decode_strips(self.dict, self.d_in.data_ptr(), self.in_bytes, self.units_ptr, self.n_strip_units,
                               out.data_ptr(), mode, stream=st)
            return out
        capi.decode_units(self.dict, self.d_in.data_ptr(), self.plan.in_bytes, self.d_segs.data_ptr(),
                          int(self.plan.n_segments), self.d_unit_first.data_ptr(), int(self.plan.n_units),
                          out.data_ptr(), mode, stream=st)
        return out

    def checksums(self, values, stream=None):
        sums = torch.empty(self.n_lists, dtype=torch.int64, device=self.dev)
        st = (stream or torch.cuda.current_stream(self.dev)).cuda_stream
        capi.checksum_lists(values.data_ptr(), self.d_list_offsets.data_ptr(), self.n_lists, sums.data_ptr(), stream=st)
        return sums


class ShardedVroomFile(VroomFile):
    """This rank's term-range shard of a vroom .bin (one process per GPU, dictionary replicated,
    no traffic on the hot path). `totals()` is the job's only collective."""

    def __init__(self, type_, bin_bytes, dict_bytes, rank, world, device=0, block_size=256):
        host_dict = capi.Dictionary(capi.KIND_OF_TYPE[type_], dict_bytes)  # host tables: enough to walk the file
        whole = capi.VroomPlan(host_dict, bin_bytes, block_size=block_size)
        lengths = np.diff(whole.list_offsets.astype(np.int64))
        mine, self.first_list, self.last_list = shard.shard_vroom(bin_bytes, lengths, whole.list_byte_offsets, rank, world)
        self.rank, self.world = rank, world
        super().__init__(type_, mine, dict_bytes, device=device, block_size=block_size)

    def totals(self, values):
        """(integers decoded, sum of per-list checksums mod 2^63) over all ranks"""
        sums = self.checksums(values)
        local = int(sums.view(torch.int64).sum().item()) if self.n_lists else 0
        return shard.reduce_totals(self.n_ints, local, device=self.dev)


class IndexFile:
    """A frozen dict_freq_index .bin resident on one GPU (dictionaries are read from the file)."""

    def __init__(self, type_, bin_bytes, device=0, block_size=256, plan="host"):
        """plan="host": header, Elias-Fano endpoints and block tables expanded on the host (dint_index_plan_create);
        plan="device": the host reads the fixed-size header only, the GPU builds the block table from the bytes in HBM
        (dint_index_device_plan_create)"""
        if type_ not in capi.KIND_OF_TYPE:
            raise ValueError(f"unknown type '{type_}'")
        self.type = type_
        self.dev = _require_cuda(device)
        kind = capi.KIND_OF_TYPE[type_]
        self.block_size = block_size
        self.docs_plan = self.freqs_plan = None
        if plan == "device":
            self.d_bin = _to_device_bytes(bin_bytes, self.dev)
            with torch.cuda.device(self.dev):
                st = torch.cuda.current_stream(self.dev).cuda_stream
                self.plan = capi.DeviceIndexPlan(kind, bin_bytes, self.d_bin.data_ptr(), device=device, block_size=block_size, stream=st)
            p = self.plan
            self.docs_dict = self.freqs_dict = None
            if kind < capi.INDEX_BLOCK_INTERPOLATIVE:
                self.docs_dict = capi.Dictionary(kind, bin_bytes[p.docs_dict_off:p.freqs_dict_off], device=device, mapped=True)
                self.freqs_dict = capi.Dictionary(kind, bin_bytes[p.freqs_dict_off:p.end_off], device=device, mapped=True)
            self.n_postings, self.n_lists = int(p.n_postings), int(p.n_lists)
            self.d_blocks = device_view(p.blocks_ptr, int(p.n_blocks) * 32, "|u1", self.dev)
            self.d_list_offsets = device_view(p.list_offsets_ptr, self.n_lists + 1, "<i8", self.dev)
            return
        self.plan = capi.IndexPlan(kind, bin_bytes, block_size=block_size)
        p = self.plan
        self.docs_dict = self.freqs_dict = None
        if kind < capi.INDEX_BLOCK_INTERPOLATIVE:
            self.docs_dict = capi.Dictionary(kind, bin_bytes[p.docs_dict_off:p.freqs_dict_off], device=device, mapped=True)
            self.freqs_dict = capi.Dictionary(kind, bin_bytes[p.freqs_dict_off:p.end_off], device=device, mapped=True)
        self.d_bin = _to_device_bytes(bin_bytes, self.dev)
        self.d_blocks = _structured_to_device(p.blocks, self.dev)
        self.d_list_offsets = torch.from_numpy(p.list_offsets.astype(np.int64)).to(self.dev)
        self.n_postings = int(p.n_postings)
        self.n_lists = int(p.n_lists)
        self.docs_plan = self.freqs_plan = None

    def discover(self, stream=None):
        """strip tables of the full blocks, found on the device from the file's own block endpoints
        (dint_index_discover); scan(engine="strips") then runs the lane-sequential decoder"""
        if self.docs_plan is None:
            st = (stream or torch.cuda.current_stream(self.dev)).cuda_stream
            p = self.plan
            with torch.cuda.device(self.dev):
                self.docs_plan, self.freqs_plan = capi.DevicePlan.discover_index(
                    self.docs_dict, self.freqs_dict, self.d_bin.data_ptr() + p.lists_off, p.lists_len, self.d_blocks.data_ptr(),
                    int(p.n_blocks), self.block_size, stream=st)
        return self.docs_plan, self.freqs_plan

    def scan(self, docids=None, freqs=None, want_docs=True, want_freqs=True, stream=None, engine="tiles"):
        """docid() and freq() of every posting of every list, packed list after list.
        engine="tiles": warp-per-block tile decoder (dint_decode_index); "strips": dint_decode_index_planned."""
        if want_docs and docids is None:
            docids = torch.empty(self.n_postings, dtype=torch.int32, device=self.dev)
        if want_freqs and freqs is None:
            freqs = torch.empty(self.n_postings, dtype=torch.int32, device=self.dev)
        st = (stream or torch.cuda.current_stream(self.dev)).cuda_stream
        p = self.plan
        if self.docs_dict is None:  # block_freq_index<interpolative | vbyte | u32 block>: every block by that codec
            capi.decode_index_blocks(capi.KIND_OF_TYPE[self.type], self.d_bin.data_ptr() + p.lists_off, p.lists_len,
                                     self.d_blocks.data_ptr(), int(p.n_blocks), self.block_size,
                                     docids.data_ptr() if want_docs else None, freqs.data_ptr() if want_freqs else None, stream=st)
            return docids, freqs
        if engine == "strips":
            dp, fp = self.discover(stream)
            capi.decode_index_planned(self.docs_dict, self.freqs_dict, self.d_bin.data_ptr() + p.lists_off, p.lists_len, dp, fp,
                                      self.d_blocks.data_ptr(), int(p.n_blocks), self.block_size,
                                      docids.data_ptr() if want_docs else None, freqs.data_ptr() if want_freqs else None,
                                      stream=st)
            return docids, freqs
        capi.decode_index(self.docs_dict, self.freqs_dict, self.d_bin.data_ptr() + p.lists_off, p.lists_len,
                          self.d_blocks.data_ptr(), int(p.n_blocks), self.block_size,
                          docids.data_ptr() if want_docs else None, freqs.data_ptr() if want_freqs else None, stream=st)
        return docids, freqs


def decode_dint(type_, encoded_data_filename, dictionary_filename, device=0, repeat=1, block_size=256):
    """vroom_env/decode.cpp:95-155 on the GPU: load, decode every list, report the reference's stats line."""
    if not dictionary_filename:
        raise RuntimeError("dictionary_filename must be specified")  # decode.cpp:98-100
    try:
        with open(encoded_data_filename, "rb") as f:
            bin_bytes = f.read()
    except OSError:
        raise RuntimeError("Error opening index file")  # decode.cpp:104-106
    with open(dictionary_filename, "rb") as f:
        dict_bytes = f.read()
    vf = VroomFile(type_, bin_bytes, dict_bytes, device=device, block_size=block_size)
    out = torch.empty(vf.n_ints, dtype=torch.int32, device=vf.dev)
    vf.decode(out)  # warm-up
    torch.cuda.synchronize(vf.dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(repeat):
        vf.decode(out)
    e1.record()
    torch.cuda.synchronize(vf.dev)
    tot_elapsed = e0.elapsed_time(e1) / 1e3 / repeat
    ns_x_int = tot_elapsed * 1e9 / max(vf.n_ints, 1)
    stats = {"filename": encoded_data_filename, "num_sequences": str(vf.n_lists), "num_integers": str(vf.n_ints),
             "type": type_, "tot_elapsed_time": str(tot_elapsed), "ns_x_int": str(ns_x_int),
             "ints_x_sec": str(int(1e9 / ns_x_int)) if ns_x_int > 0 else "0"}
    return stats, out


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="GPU counterpart of vroom_env/decode")
    ap.add_argument("type")
    ap.add_argument("encoded_data_filename")
    ap.add_argument("--dict", dest="dictionary_filename")
    ap.add_argument("--repeat", type=int, default=10)
    a = ap.parse_args(argv)
    stats, _ = decode_dint(a.type, a.encoded_data_filename, a.dictionary_filename, repeat=a.repeat)
    print(json.dumps(stats))


if __name__ == "__main__":
    main()
