"""ctypes wrapper around oracle/libraft_oracle.so and the compiled reference binary oracle/_ref/raft.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never by the product package raft_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libraft_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "raft")


class OrcParams(C.Structure):
    _fields_ = [("reso", C.c_int32), ("est_cov", C.c_int32), ("cov_mul", C.c_double),
                ("repeat_length", C.c_int32), ("interval_length", C.c_int32),
                ("read_length", C.c_int32), ("overlap_length", C.c_int32),
                ("flanking_length", C.c_int32)]


class OrcReads(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("seq_off", C.c_void_p), ("seq", C.c_void_p),
                ("name_off", C.c_void_p), ("names", C.c_void_p)]


_P32 = C.POINTER(C.c_int32)
_P64 = C.POINTER(C.c_int64)
_P8 = C.POINTER(C.c_uint8)


class OrcResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("bad_index", C.c_int64),
                ("n_rec", C.c_int64),
                ("qid", _P32), ("tid", _P32), ("qs", _P32), ("qe", _P32), ("ts", _P32), ("te", _P32),
                ("strand", _P8),
                ("symmetric", C.c_int32), ("high_cov", C.c_int32), ("real_reads", C.c_int32),
                ("bin_off", _P64), ("cov", _P32),
                ("rep_off", _P64), ("rep_s", _P32), ("rep_e", _P32),
                ("n_frag", C.c_int64), ("frag_read", _P32), ("frag_a", _P32), ("frag_b", _P32),
                ("total_cov", C.c_int64), ("total_windows", C.c_int32),
                ("total_repeat_len", C.c_int64), ("total_read_len", C.c_int64),
                ("cov_txt", _P8), ("cov_txt_len", C.c_int64),
                ("rep_txt", _P8), ("rep_txt_len", C.c_int64),
                ("bed_txt", _P8), ("bed_txt_len", C.c_int64),
                ("fasta", _P8), ("fasta_len", C.c_int64)]


class OrcDigestResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("bad_index", C.c_int64), ("n_rec", C.c_int64),
                ("symmetric", C.c_int32), ("high_cov", C.c_int32), ("real_reads", C.c_int32),
                ("n_frag", C.c_int64), ("n_rep", C.c_int64), ("total_cov", C.c_int64), ("total_windows", C.c_int32),
                ("total_repeat_len", C.c_int64), ("total_read_len", C.c_int64),
                ("digest", C.c_uint64 * 4), ("bytes", C.c_int64 * 4)]


class OrcFasta(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("seq_off", _P64), ("seq", _P8), ("name_off", _P64), ("names", _P8)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.check_call(["make", "-C", HERE, "libraft_oracle.so"])
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_run.restype = C.c_int
        _lib.orc_run.argtypes = [C.POINTER(OrcReads), C.c_void_p, C.c_int64, C.POINTER(OrcParams), C.c_int,
                                 C.POINTER(OrcResult)]
        _lib.orc_free.argtypes = [C.POINTER(OrcResult)]
        _lib.orc_run_digest.restype = C.c_int
        _lib.orc_run_digest.argtypes = [C.POINTER(OrcReads), C.c_void_p, C.c_int64, C.POINTER(OrcParams), C.c_int,
                                        C.POINTER(OrcDigestResult)]
        _lib.orc_parse_fasta.restype = C.c_int64
        _lib.orc_parse_fasta.argtypes = [C.c_void_p, C.c_int64, C.POINTER(OrcFasta)]
        _lib.orc_free_fasta.argtypes = [C.POINTER(OrcFasta)]
        _lib.orc_split_naive.restype = C.c_int64
        _lib.orc_split_naive.argtypes = [C.POINTER(OrcReads), C.c_int32, C.POINTER(_P8)]
        _lib.orc_digest.restype = C.c_uint64
        _lib.orc_digest.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
    return _lib


def _arr(ptr, n, dtype):
    if n <= 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype, copy=True)


def _bytes(ptr, n):
    return C.string_at(ptr, int(n)) if n > 0 else b""


def make_params(reso=50, est_cov=0, cov_mul=1.5, repeat_length=10000, interval_length=None,
                read_length=20000, overlap_length=500, flanking_length=1000):
    if interval_length is None:
        interval_length = repeat_length  # main.cpp:44-47: -p sets both
    return OrcParams(reso, est_cov, cov_mul, repeat_length, interval_length, read_length,
                     overlap_length, flanking_length)


class Result:
    """Python-side copy of orc_result_t (numpy arrays + bytes)."""


def run(reads, paf: bytes, params: OrcParams, text=True):
    """reads: object with n, seq_off(int64[n+1]), seq(uint8), name_off(int64[n+1]), names(uint8)."""
    L = lib()
    seq_off = np.ascontiguousarray(reads.seq_off, dtype=np.int64)
    name_off = np.ascontiguousarray(reads.name_off, dtype=np.int64)
    seq = np.ascontiguousarray(reads.seq, dtype=np.uint8)
    names = np.ascontiguousarray(reads.names, dtype=np.uint8)
    n = len(seq_off) - 1
    rd = OrcReads(n, seq_off.ctypes.data, seq.ctypes.data if seq.size else None,
                  name_off.ctypes.data, names.ctypes.data if names.size else None)
    pafb = np.frombuffer(paf, dtype=np.uint8) if len(paf) else np.zeros(0, np.uint8)
    res = OrcResult()
    st = L.orc_run(C.byref(rd), pafb.ctypes.data if pafb.size else None, len(paf), C.byref(params),
                   0 if text else 1, C.byref(res))
    out = Result()
    out.status = st
    out.bad_index = res.bad_index
    if st == 0:
        N = res.n_rec
        out.n_rec = N
        for k in ("qid", "tid", "qs", "qe", "ts", "te"):
            setattr(out, k, _arr(getattr(res, k), N, np.int32))
        out.strand = _arr(res.strand, N, np.uint8)
        out.symmetric, out.high_cov, out.real_reads = res.symmetric, res.high_cov, res.real_reads
        out.bin_off = _arr(res.bin_off, n + 1, np.int64)
        B = int(out.bin_off[-1]) if n >= 0 else 0
        out.cov = _arr(res.cov, B, np.int32)
        out.rep_off = _arr(res.rep_off, n + 1, np.int64)
        R = int(out.rep_off[-1])
        out.rep_s, out.rep_e = _arr(res.rep_s, R, np.int32), _arr(res.rep_e, R, np.int32)
        G = res.n_frag
        out.n_frag = G
        out.frag_read, out.frag_a, out.frag_b = (_arr(res.frag_read, G, np.int32), _arr(res.frag_a, G, np.int32),
                                                  _arr(res.frag_b, G, np.int32))
        out.total_cov, out.total_windows = res.total_cov, res.total_windows
        out.total_repeat_len, out.total_read_len = res.total_repeat_len, res.total_read_len
        if text:
            out.cov_txt = _bytes(res.cov_txt, res.cov_txt_len)
            out.rep_txt = _bytes(res.rep_txt, res.rep_txt_len)
            out.bed_txt = _bytes(res.bed_txt, res.bed_txt_len)
            out.fasta = _bytes(res.fasta, res.fasta_len)
    L.orc_free(C.byref(res))
    return out


def run_digest(reads, paf, params: OrcParams, threads=None):
    """orc_run_digest: the whole path with `threads` host threads, outputs reported as (bytes, digest) per file.
    `paf` may be bytes or a uint8 numpy array; nothing is copied.  Returns the OrcDigestResult structure
    (digest[k] / bytes[k] indexed coverage.txt, long_repeats.txt, long_repeats.bed, reads.fasta)."""
    L = lib()
    seq_off = np.ascontiguousarray(reads.seq_off, dtype=np.int64)
    name_off = np.ascontiguousarray(reads.name_off, dtype=np.int64)
    seq = np.ascontiguousarray(reads.seq, dtype=np.uint8)
    names = np.ascontiguousarray(reads.names, dtype=np.uint8)
    rd = OrcReads(len(seq_off) - 1, seq_off.ctypes.data, seq.ctypes.data if seq.size else None,
                  name_off.ctypes.data, names.ctypes.data if names.size else None)
    pafb = np.frombuffer(paf, dtype=np.uint8) if isinstance(paf, (bytes, bytearray)) else np.ascontiguousarray(paf, dtype=np.uint8)
    res = OrcDigestResult()
    L.orc_run_digest(C.byref(rd), pafb.ctypes.data if pafb.size else None, pafb.size, C.byref(params),
                     int(threads or os.cpu_count() or 1), C.byref(res))
    return res


class Reads:
    def __init__(self, seq_off, seq, name_off, names):
        self.seq_off, self.seq, self.name_off, self.names = seq_off, seq, name_off, names
        self.n = len(seq_off) - 1


def parse_fasta(text: bytes) -> Reads:
    L = lib()
    buf = np.frombuffer(text, dtype=np.uint8) if len(text) else np.zeros(0, np.uint8)
    f = OrcFasta()
    n = L.orc_parse_fasta(buf.ctypes.data if buf.size else None, len(text), C.byref(f))
    if n < 0:
        raise RuntimeError(f"orc_parse_fasta failed: {n}")
    seq_off = _arr(f.seq_off, n + 1, np.int64)
    name_off = _arr(f.name_off, n + 1, np.int64)
    seq = _arr(f.seq, int(seq_off[-1]), np.uint8)
    names = _arr(f.names, int(name_off[-1]), np.uint8)
    L.orc_free_fasta(C.byref(f))
    return Reads(seq_off, seq, name_off, names)


REF_SPLIT_BIN = os.path.join(HERE, "_ref", "split_naive")


def split_naive(reads, sublen: int) -> bytes:
    L = lib()
    seq_off = np.ascontiguousarray(reads.seq_off, dtype=np.int64)
    name_off = np.ascontiguousarray(reads.name_off, dtype=np.int64)
    seq = np.ascontiguousarray(reads.seq, dtype=np.uint8)
    names = np.ascontiguousarray(reads.names, dtype=np.uint8)
    rd = OrcReads(len(seq_off) - 1, seq_off.ctypes.data, seq.ctypes.data if seq.size else None, name_off.ctypes.data,
                  names.ctypes.data if names.size else None)
    out = _P8()
    n = L.orc_split_naive(C.byref(rd), sublen, C.byref(out))
    if n < 0:
        raise RuntimeError(f"orc_split_naive: {n}")
    data = _bytes(out, n)
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    libc.free(C.cast(out, C.c_void_p))
    return data


def digest(data: bytes, abs_offset=0) -> int:
    buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(0, np.uint8)
    return int(lib().orc_digest(buf.ctypes.data if buf.size else None, len(data), abs_offset))


def have_ref() -> bool:
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def run_ref(reads_path, paf_path, workdir, args, prefix_arg=True, timeout=600):
    """Run the unmodified reference binary; returns (returncode, stdout, {suffix: bytes}).

    args: list of CLI flags WITHOUT -o; outputs go to workdir/out.* (the -o is appended last so the
    -v fallthrough quirk, main.cpp:51-55, cannot rename them).
    """
    prefix = os.path.join(workdir, "out")
    cmd = [REF_BIN] + list(args) + (["-o", prefix] if prefix_arg else []) + [reads_path, paf_path]
    p = subprocess.run(cmd, cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout)
    outs = {}
    for suf in ("coverage.txt", "long_repeats.txt", "long_repeats.bed", "reads.fasta"):
        fp = f"{prefix}.{suf}"
        if os.path.exists(fp):
            with open(fp, "rb") as fh:
                outs[suf] = fh.read()
    return p.returncode, p.stdout.decode(errors="replace"), outs
