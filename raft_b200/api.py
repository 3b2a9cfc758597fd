"""Host-side mirror of the reference's interface for the fragmentation path, on top of the C ABI.

Reference seam (SURVEY.md §8 row b): `break_long_reads(reads, paf, unused, algoParams&)` (chop.hpp:331)
and its stages loadFASTA / create_pileup / repeat_annotate / break_reads.  Names and argument meaning
follow the reference; errors the reference reports by exit(1) or by crashing surface as RaftError.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib

OUT_COVERAGE, OUT_LONG_REPEATS, OUT_BED, OUT_READS_FASTA = 0, 1, 2, 3
OUT_SPLIT_NAIVE = 4
OPT_DEFER_SEQ_UPLOAD = 1
OUT_SUFFIX = {OUT_COVERAGE: "coverage.txt", OUT_LONG_REPEATS: "long_repeats.txt", OUT_BED: "long_repeats.bed",
              OUT_READS_FASTA: "reads.fasta"}
(TAB_QID, TAB_TID, TAB_QS, TAB_QE, TAB_TS, TAB_TE, TAB_STRAND, TAB_BIN_OFF, TAB_COV, TAB_REP_OFF, TAB_REP,
 TAB_FRAG) = range(12)
_TAB_DTYPE = {TAB_STRAND: np.uint8, TAB_BIN_OFF: np.int64, TAB_REP_OFF: np.int64}


class RaftError(RuntimeError):
    def __init__(self, status, detail="", index=-1):
        self.status, self.index = status, index
        msg = _lib.lib().raftgpu_strerror(status).decode()
        super().__init__(f"[{status}] {msg}" + (f": {detail}" if detail else ""))


@dataclass
class AlgoParams:
    """algoParams (param.hpp:4-31).  -p sets both repeat_length and interval_length (main.cpp:44-47)."""
    reso: int = 50
    est_cov: int = 0
    cov_mul: float = 1.5
    repeat_length: int = 10000
    interval_length: int = None
    read_length: int = 20000
    overlap_length: int = 500
    flanking_length: int = 1000
    outputfilename: str = "raft"

    def c_struct(self):
        il = self.repeat_length if self.interval_length is None else self.interval_length
        return _lib.Params(self.reso, self.est_cov, self.cov_mul, self.repeat_length, il, self.read_length,
                           self.overlap_length, self.flanking_length)

    @classmethod
    def from_args(cls, args):
        """Parse reference CLI flags (main.cpp:28-59), including the -v -> -o fallthrough."""
        p = cls()
        it = iter(args)
        for flag in it:
            val = next(it)
            if flag == "-r": p.reso = int(val)
            elif flag == "-e": p.est_cov = int(val)
            elif flag == "-m": p.cov_mul = float(val)
            elif flag == "-l": p.read_length = int(val)
            elif flag == "-p": p.repeat_length = int(val); p.interval_length = int(val)
            elif flag == "-f": p.flanking_length = int(val)
            elif flag == "-v": p.overlap_length = int(val); p.outputfilename = val
            elif flag == "-o": p.outputfilename = val
            else: raise ValueError(flag)
        return p


def _ptr(a):
    """Host numpy array / bytes / torch tensor (host or CUDA) / int address -> void* value."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data if a.size else None
    if isinstance(a, bytes):
        return C.cast(C.c_char_p(a), C.c_void_p).value if len(a) else None  # the bytes object itself (kept alive by the caller)
    if isinstance(a, bytearray):
        return C.addressof((C.c_char * len(a)).from_buffer(a)) if len(a) else None  # in place, no temporary copy
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


class Context:
    """One GPU's state for the path (raftgpu_ctx)."""

    def __init__(self, params: AlgoParams, device: int = 0):
        self.L = _lib.lib()
        self.params = params
        self._h = C.c_void_p()
        self._keep = []
        st = self.L.raftgpu_create(C.byref(params.c_struct()), device, C.byref(self._h))
        if st:
            raise RaftError(st, "raftgpu_create (is there a B200? the library has no CPU fallback)")

    def close(self):
        if self._h:
            self.L.raftgpu_destroy(self._h)
            self._h = C.c_void_p()
        self._keep = []  # borrowed inputs may be released now

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, st):
        if st:
            raise RaftError(st, self.L.raftgpu_last_error(self._h).decode(), self.L.raftgpu_error_index(self._h))

    def set_option(self, option, value):
        self._ck(self.L.raftgpu_set_option(self._h, option, int(value)))

    # ---- a0 loadFASTA (chop.hpp:88-131), after tokenisation
    def set_reads(self, seq_off, seq, name_off, names):
        # the previous inputs stay referenced until the C call has returned: it starts with raftgpu_reset, which waits for
        # a deferred upload that may still be reading the old host arena
        old, new = self._keep, [seq_off, seq, name_off, names]
        n = len(seq_off) - 1
        try:
            self._ck(self.L.raftgpu_set_reads(self._h, n, _ptr(seq_off), _ptr(seq), _ptr(name_off), _ptr(names)))
        finally:
            self._keep = new
            del old

    def ingest_fasta(self, text, nbytes=None, last=True, total_hint=0):
        """Device FASTA tokenizer (loadFASTA's record grammar, chop.hpp:88-131); raises RaftError(-12) for FASTQ / CRLF text."""
        if nbytes is None:
            nbytes = len(text) if not hasattr(text, "numel") else text.numel()
        self._keep.append(text)
        self._ck(self.L.raftgpu_ingest_fasta(self._h, _ptr(text), nbytes, 1 if last else 0, total_hint))

    def set_reads_sharded(self, n, lengths, name_off, names, own_first, own_count, own_seq_off, own_seq):
        old, new = self._keep, [lengths, name_off, names, own_seq_off, own_seq]
        try:
            self._ck(self.L.raftgpu_set_reads_sharded(self._h, n, _ptr(lengths), _ptr(name_off), _ptr(names), own_first,
                                                      own_count, _ptr(own_seq_off), _ptr(own_seq)))
        finally:
            self._keep = new
            del old

    def split_naive(self, subread_length: int):
        """split_naive.cpp:10-44 on the device; bytes through fetch(OUT_SPLIT_NAIVE)."""
        self._ck(self.L.raftgpu_split_naive(self._h, subread_length))

    # ---- a1/a2 create_pileup (chop.hpp:133-191)
    def ingest_paf(self, text, nbytes=None, last=True):
        if nbytes is None:
            nbytes = len(text) if not hasattr(text, "numel") else text.numel()
        self._keep.append(text)
        self._ck(self.L.raftgpu_ingest_paf(self._h, _ptr(text), nbytes, 1 if last else 0))

    # ---- a3-a5 repeat_annotate + break_reads arithmetic
    def run(self):
        s = _lib.Stats()
        self._ck(self.L.raftgpu_run(self._h, C.byref(s)))
        return s

    def stats(self):
        s = _lib.Stats()
        self._ck(self.L.raftgpu_get_stats(self._h, C.byref(s)))
        return s

    def finalize(self):
        s = _lib.Stats()
        self._ck(self.L.raftgpu_finalize(self._h, C.byref(s)))
        return s

    def output_size(self, which):
        n = C.c_uint64()
        self._ck(self.L.raftgpu_output_size(self._h, which, C.byref(n)))
        return n.value

    def fetch(self, which, off=0, n=None) -> bytes:
        if n is None:
            n = self.output_size(which) - off
        buf = np.empty(max(n, 1), dtype=np.uint8)
        self._ck(self.L.raftgpu_fetch(self._h, which, off, buf.ctypes.data, n))
        return buf[:n].tobytes()

    def fetch_into(self, which, off, dst, n):
        self._ck(self.L.raftgpu_fetch(self._h, which, off, _ptr(dst), n))

    def fetch_async(self, which, off, dst, n):
        """Queue the emitter for a device destination; pair with sync()."""
        self._ck(self.L.raftgpu_fetch_async(self._h, which, off, _ptr(dst), n))

    def reset(self):
        """Forget reads, PAF and results; device buffers stay allocated for the next run."""
        self._ck(self.L.raftgpu_reset(self._h))
        self._keep = []

    def sync(self):
        self._ck(self.L.raftgpu_sync(self._h))

    def digest(self, which, stream_base=0) -> int:
        """64-bit digest of a whole output stream, computed on the device; `stream_base` = file offset of this
        context's slice when the run is sharded (per-rank digests then add up to the file's)."""
        d = C.c_uint64()
        self._ck(self.L.raftgpu_digest_at(self._h, which, stream_base, C.byref(d)))
        return d.value

    def table(self, tab) -> np.ndarray:
        n = C.c_size_t()
        self._ck(self.L.raftgpu_fetch_table(self._h, tab, None, 0, C.byref(n)))
        out = np.empty(n.value, dtype=_TAB_DTYPE.get(tab, np.int32))
        if n.value:
            self._ck(self.L.raftgpu_fetch_table(self._h, tab, out.ctypes.data, out.nbytes, C.byref(n)))
        return out

    # ---- multi-GPU inside the library (NCCL): raftgpu_comm_init + raftgpu_run_sharded
    def comm_init(self, nranks, rank, comm_id: bytes):
        """Collective over the ranks: the same 128-byte id (comm_unique_id() of one rank) everywhere."""
        buf = (C.c_uint8 * _lib.COMM_ID_BYTES).from_buffer_copy(comm_id)
        self._ck(self.L.raftgpu_comm_init(self._h, nranks, rank, C.addressof(buf)))

    def comm_destroy(self):
        self._ck(self.L.raftgpu_comm_destroy(self._h))

    def run_sharded(self, bounds: np.ndarray, text, nbytes=None):
        """raftgpu_ingest_paf + raftgpu_run for one rank of a sharded run; returns (stats, shard_info)."""
        if nbytes is None:
            nbytes = len(text) if not hasattr(text, "numel") else text.numel()
        self._keep.append(text)
        bounds = np.ascontiguousarray(bounds, dtype=np.int64)
        s, info = _lib.Stats(), _lib.ShardInfo()
        self._ck(self.L.raftgpu_run_sharded(self._h, bounds.ctypes.data, _ptr(text), nbytes, C.byref(s), C.byref(info)))
        return s, info

    # ---- multi-GPU building blocks (the caller runs the exchange)
    def peek_first_record(self, text, nbytes):
        rec, found = (C.c_int32 * 6)(), C.c_int32()
        self._ck(self.L.raftgpu_peek_first_record(self._h, _ptr(text), nbytes, C.byref(rec), C.byref(found)))
        return list(rec), bool(found.value)

    def set_first_record(self, rec, is_local):
        arr = (C.c_int32 * 6)(*rec) if rec is not None else None
        self._ck(self.L.raftgpu_set_first_record(self._h, C.byref(arr) if arr is not None else None, 1 if is_local else 0))

    def get_symmetric(self):
        f = C.c_int32()
        self._ck(self.L.raftgpu_get_symmetric(self._h, C.byref(f)))
        return int(f.value)

    def set_symmetric(self, flag):
        self._ck(self.L.raftgpu_set_symmetric(self._h, int(flag)))

    def route_count(self, bounds: np.ndarray) -> np.ndarray:
        nr = len(bounds) - 1
        counts = np.zeros(nr, dtype=np.int64)
        self._ck(self.L.raftgpu_route_count(self._h, nr, bounds.ctypes.data, counts.ctypes.data))
        return counts

    def route_pack(self, bounds: np.ndarray, counts: np.ndarray, sendbuf):
        self._ck(self.L.raftgpu_route_pack(self._h, len(counts), bounds.ctypes.data, counts.ctypes.data, _ptr(sendbuf)))

    def accumulate_local(self):
        self._ck(self.L.raftgpu_accumulate_local(self._h))

    def accumulate_endpoints(self, ep, count):
        self._ck(self.L.raftgpu_accumulate_endpoints(self._h, _ptr(ep), count))

    def set_output_base(self, first_read_num):
        self._ck(self.L.raftgpu_set_output_base(self._h, first_read_num))


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library: created by one rank, handed to every rank's Context.comm_init."""
    buf = (C.c_uint8 * _lib.COMM_ID_BYTES)()
    st = _lib.lib().raftgpu_comm_unique_id(C.addressof(buf))
    if st:
        raise RaftError(st, "raftgpu_comm_unique_id")
    return bytes(buf)


def partition_reads(lengths, reso, nranks) -> np.ndarray:
    """Read-id range boundaries (int64[nranks+1]) balanced by coverage slots (raftgpu_partition_reads)."""
    lengths = np.ascontiguousarray(lengths, dtype=np.int64)
    bounds = np.zeros(nranks + 1, dtype=np.int64)
    st = _lib.lib().raftgpu_partition_reads(lengths.ctypes.data, len(lengths), reso, nranks, bounds.ctypes.data)
    if st:
        raise RaftError(st, "raftgpu_partition_reads")
    return bounds


def load_fasta(path):
    """loadFASTA's record reader (chop.hpp:88-131 / kseq.h:240-298) -> (seq_off, seq, name_off, names)."""
    L = _lib.lib()
    n = C.c_int64()
    so, sq, no, nm = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    st = L.raftgpu_load_fasta(path.encode(), C.byref(n), C.byref(so), C.byref(sq), C.byref(no), C.byref(nm))
    if st:
        raise RaftError(st, path)
    try:
        seq_off = np.ctypeslib.as_array(C.cast(so, C.POINTER(C.c_int64)), shape=(n.value + 1,)).copy()
        name_off = np.ctypeslib.as_array(C.cast(no, C.POINTER(C.c_int64)), shape=(n.value + 1,)).copy()
        ns, nn = int(seq_off[-1]), int(name_off[-1])
        seq = np.ctypeslib.as_array(C.cast(sq, C.POINTER(C.c_uint8)), shape=(max(ns, 1),))[:ns].copy()
        names = np.ctypeslib.as_array(C.cast(nm, C.POINTER(C.c_uint8)), shape=(max(nn, 1),))[:nn].copy()
    finally:
        for p in (so, sq, no, nm):
            L.raftgpu_free_host(p)
    return seq_off, seq, name_off, names


def break_long_reads(readfilename: str, paffilename, params: AlgoParams, device=0):
    """Drop-in for break_long_reads (chop.hpp:331-373): files in, prefix.* files out.

    `paffilename` may be a list of paths: they are ingested back to back as `cat` would join them
    (README.md:35-36 merges hifiasm's *.0.ovlp.paf and *.1.ovlp.paf before calling raft).
    `device` may be a list of GPU indices: the run is then sharded over them (raftgpu_break_long_reads_mgpu)."""
    L = _lib.lib()
    s = _lib.Stats()
    pafs = [paffilename] if isinstance(paffilename, (str, bytes)) else list(paffilename)
    arr = (C.c_char_p * len(pafs))(*[q.encode() if isinstance(q, str) else q for q in pafs])
    if isinstance(device, (list, tuple)):
        devs = (C.c_int * len(device))(*device)
        st = L.raftgpu_break_long_reads_mgpu(readfilename.encode(), len(pafs), arr, C.byref(params.c_struct()),
                                             params.outputfilename.encode(), len(device), devs, C.byref(s))
        if st:
            raise RaftError(st)
        return s
    st = L.raftgpu_break_long_reads_multi(readfilename.encode(), len(pafs), arr, C.byref(params.c_struct()),
                                          params.outputfilename.encode(), device, C.byref(s))
    if st:
        raise RaftError(st)
    return s
