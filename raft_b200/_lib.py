"""ctypes binding of libraft_b200.so (include/raft_b200.h).  Fails loudly when the library is missing:
there is no CPU fallback for the fragmentation path."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RAFT_B200_LIB") or os.path.join(HERE, "libraft_b200.so")  # the variable selects an experimental build

# every symbol include/raft_b200.h declares
SYMBOLS = [
    "raftgpu_default_params", "raftgpu_create", "raftgpu_destroy", "raftgpu_reset", "raftgpu_strerror",
    "raftgpu_last_error", "raftgpu_error_index", "raftgpu_set_option", "raftgpu_set_reads", "raftgpu_ingest_fasta", "raftgpu_load_fasta", "raftgpu_free_host",
    "raftgpu_split_naive", "raftgpu_ingest_paf", "raftgpu_run", "raftgpu_get_stats", "raftgpu_output_size", "raftgpu_fetch", "raftgpu_fetch_async", "raftgpu_sync", "raftgpu_digest", "raftgpu_digest_at",
    "raftgpu_fetch_table", "raftgpu_set_reads_sharded", "raftgpu_peek_first_record", "raftgpu_set_first_record",
    "raftgpu_get_symmetric", "raftgpu_set_symmetric", "raftgpu_route_count", "raftgpu_route_pack",
    "raftgpu_accumulate_local", "raftgpu_accumulate_endpoints", "raftgpu_finalize", "raftgpu_set_output_base", "raftgpu_break_long_reads",
    "raftgpu_break_long_reads_multi", "raftgpu_comm_unique_id", "raftgpu_comm_init", "raftgpu_comm_destroy", "raftgpu_run_sharded",
    "raftgpu_partition_reads", "raftgpu_break_long_reads_mgpu", "raftgpu_sharded_begin", "raftgpu_sharded_finish",
    "raftgpu_reads_info", "raftgpu_reads_copy", "raftgpu_reads_device",
]


class Params(C.Structure):
    """raftgpu_params == algoParams (param.hpp:4-16)."""
    _fields_ = [("reso", C.c_int32), ("est_cov", C.c_int32), ("cov_mul", C.c_double),
                ("repeat_length", C.c_int32), ("interval_length", C.c_int32), ("read_length", C.c_int32),
                ("overlap_length", C.c_int32), ("flanking_length", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_records", C.c_int64), ("symmetric", C.c_int32), ("high_cov", C.c_int32),
                ("real_reads", C.c_int32), ("total_windows", C.c_int32), ("total_cov", C.c_int64),
                ("total_repeat_len", C.c_int64), ("total_read_len", C.c_int64), ("n_bins", C.c_int64),
                ("n_repeats", C.c_int64), ("n_fragments", C.c_int64), ("out_bytes", C.c_uint64 * 4),
                ("ms_tokenize", C.c_float), ("ms_scatter", C.c_float), ("ms_scan", C.c_float),
                ("ms_repeat_cut", C.c_float), ("ms_layout", C.c_float), ("ms_total", C.c_float),
                ("kernel_launches", C.c_int32), ("reserved", C.c_int32),
                ("ms_emit", C.c_float * 4), ("emit_launches", C.c_int32 * 4), ("emit_bytes", C.c_uint64 * 4),
                ("ms_set_reads", C.c_float), ("reserved2", C.c_int32)]


class ShardInfo(C.Structure):
    """raftgpu_shard_info"""
    _fields_ = [("nranks", C.c_int32), ("rank", C.c_int32), ("symmetric", C.c_int32), ("reserved", C.c_int32),
                ("n_records_total", C.c_int64), ("n_fragments_total", C.c_int64), ("first_read_num", C.c_int64),
                ("endpoints_sent", C.c_int64), ("endpoints_received", C.c_int64),
                ("stream_base", C.c_uint64 * 4), ("stream_total", C.c_uint64 * 4),
                ("total_cov", C.c_int64), ("total_repeat_len", C.c_int64), ("total_read_len", C.c_int64), ("n_bins_total", C.c_int64),
                ("ms_exchange", C.c_float), ("peek_retries", C.c_int32)]


COMM_ID_BYTES = 128
_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()); "
                           "raft_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, u64, sz = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_size_t
    PP, PS = C.POINTER(Params), C.POINTER(Stats)
    sig = {
        "raftgpu_default_params": (None, [PP]),
        "raftgpu_create": (C.c_int, [PP, C.c_int, C.POINTER(vp)]),
        "raftgpu_destroy": (C.c_int, [vp]),
        "raftgpu_reset": (C.c_int, [vp]),
        "raftgpu_strerror": (C.c_char_p, [C.c_int]),
        "raftgpu_last_error": (C.c_char_p, [vp]),
        "raftgpu_error_index": (i64, [vp]),
        "raftgpu_set_option": (C.c_int, [vp, C.c_int, i64]),
        "raftgpu_set_reads": (C.c_int, [vp, i64, vp, vp, vp, vp]),
        "raftgpu_ingest_fasta": (C.c_int, [vp, vp, sz, C.c_int, u64]),
        "raftgpu_load_fasta": (C.c_int, [C.c_char_p, C.POINTER(i64), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
        "raftgpu_free_host": (None, [vp]),
        "raftgpu_split_naive": (C.c_int, [vp, i32]),
        "raftgpu_ingest_paf": (C.c_int, [vp, vp, sz, C.c_int]),
        "raftgpu_run": (C.c_int, [vp, PS]),
        "raftgpu_get_stats": (C.c_int, [vp, PS]),
        "raftgpu_output_size": (C.c_int, [vp, C.c_int, C.POINTER(u64)]),
        "raftgpu_fetch": (C.c_int, [vp, C.c_int, u64, vp, sz]),
        "raftgpu_fetch_async": (C.c_int, [vp, C.c_int, u64, vp, sz]),
        "raftgpu_sync": (C.c_int, [vp]),
        "raftgpu_digest": (C.c_int, [vp, C.c_int, C.POINTER(u64)]),
        "raftgpu_digest_at": (C.c_int, [vp, C.c_int, u64, C.POINTER(u64)]),
        "raftgpu_fetch_table": (C.c_int, [vp, C.c_int, vp, sz, C.POINTER(sz)]),
        "raftgpu_set_reads_sharded": (C.c_int, [vp, i64, vp, vp, vp, i64, i64, vp, vp]),
        "raftgpu_peek_first_record": (C.c_int, [vp, vp, sz, C.POINTER(i32 * 6), C.POINTER(i32)]),
        "raftgpu_set_first_record": (C.c_int, [vp, C.POINTER(i32 * 6), i32]),
        "raftgpu_get_symmetric": (C.c_int, [vp, C.POINTER(i32)]),
        "raftgpu_set_symmetric": (C.c_int, [vp, i32]),
        "raftgpu_route_count": (C.c_int, [vp, C.c_int, vp, vp]),
        "raftgpu_route_pack": (C.c_int, [vp, C.c_int, vp, vp, vp]),
        "raftgpu_accumulate_local": (C.c_int, [vp]),
        "raftgpu_accumulate_endpoints": (C.c_int, [vp, vp, i64]),
        "raftgpu_finalize": (C.c_int, [vp, PS]),
        "raftgpu_set_output_base": (C.c_int, [vp, i64]),
        "raftgpu_break_long_reads": (C.c_int, [C.c_char_p, C.c_char_p, PP, C.c_char_p, C.c_int, PS]),
        "raftgpu_break_long_reads_multi": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), PP, C.c_char_p, C.c_int, PS]),
        "raftgpu_comm_unique_id": (C.c_int, [vp]),
        "raftgpu_comm_init": (C.c_int, [vp, C.c_int, C.c_int, vp]),
        "raftgpu_comm_destroy": (C.c_int, [vp]),
        "raftgpu_run_sharded": (C.c_int, [vp, vp, vp, sz, PS, C.POINTER(ShardInfo)]),
        "raftgpu_partition_reads": (C.c_int, [vp, i64, i32, C.c_int, vp]),
        "raftgpu_sharded_begin": (C.c_int, [vp, vp, vp, sz, C.c_int]),
        "raftgpu_sharded_finish": (C.c_int, [vp, PS, C.POINTER(ShardInfo)]),
        "raftgpu_reads_info": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
        "raftgpu_reads_copy": (C.c_int, [vp, vp, vp, vp]),
        "raftgpu_reads_device": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp)]),
        "raftgpu_break_long_reads_mgpu": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), PP, C.c_char_p, C.c_int, C.POINTER(C.c_int), PS]),
    }
    for name in SYMBOLS:
        fn = getattr(L, name)  # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = sig[name]
    _lib = L
    return L
