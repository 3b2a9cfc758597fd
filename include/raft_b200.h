/*
 * raft_b200.h — C ABI of the B200-native RAFT fragmentation path (libraft_b200.so).
 *
 * The reference (at-cg/RAFT) has no plugin/FFI layer: its seam is the header-level function
 * boundary inside one translation unit.  Each entry point below names the reference function
 * (file:line under /root/reference) it replaces.  Plain pointers and sizes only; no C++ or torch
 * types cross this boundary.  All functions return 0 on success or a negative raftgpu_status.
 *
 * Pointer arguments marked [host|device] may be host memory (pageable or pinned) or device memory
 * of the context's GPU; the library detects which (cudaPointerGetAttributes).  Device inputs are
 * used in place (zero copy) and must stay valid until raftgpu_destroy / the next raftgpu_reset.
 *
 * There is NO CPU fallback: every compute entry point fails with RAFTGPU_E_CUDA when no sm_100
 * device is usable.
 */
#ifndef RAFT_B200_H
#define RAFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct raftgpu_ctx raftgpu_ctx;

/* Mirrors algoParams (param.hpp:4-16; defaults param.hpp:18-31). */
typedef struct {
    int32_t reso;            /* -r  (main.cpp:32-34) */
    int32_t est_cov;         /* -e  (main.cpp:35-37) */
    double  cov_mul;         /* -m  (main.cpp:38-40) */
    int32_t repeat_length;   /* -p  (main.cpp:44-47 sets both) */
    int32_t interval_length; /* -p */
    int32_t read_length;     /* -l  (main.cpp:41-43) */
    int32_t overlap_length;  /* -v  (main.cpp:51-52) */
    int32_t flanking_length; /* -f  (main.cpp:48-50) */
} raftgpu_params;

typedef enum {
    RAFTGPU_OK = 0,
    RAFTGPU_E_PARAM = -1,        /* reso<1, p<1, l<p: the reference divides by zero (chop.hpp:248,270; repeat.hpp:32) */
    RAFTGPU_E_UNKNOWN_NAME = -2, /* PAF name absent from the reads: reference indexes out of range (chop.hpp:162-168) */
    RAFTGPU_E_DUP_NAME = -3,     /* duplicate read name: reference aliases ids (chop.hpp:73-85) */
    RAFTGPU_E_RANGE = -4,        /* interval past the last bin: reference writes out of bounds (repeat.hpp:69-73) */
    RAFTGPU_E_NEG_START = -5,    /* fragment start < 0 (-v larger than a star): reference throws (chop.hpp:318) */
    RAFTGPU_E_NOMEM = -6,
    RAFTGPU_E_FASTQ = -7,        /* truncated FASTQ quality (kseq.h:290-296) */
    RAFTGPU_E_CUDA = -8,         /* CUDA runtime error / no usable device; see raftgpu_last_error */
    RAFTGPU_E_STATE = -9,        /* call order violated */
    RAFTGPU_E_IO = -10,          /* input file missing or empty (chop.hpp:336-349) / write failure */
    RAFTGPU_E_ARG = -11,         /* bad argument */
    RAFTGPU_E_UNSUPPORTED = -12,
    RAFTGPU_E_SIM_NAME = -13,    /* simulated-read mode (chop.hpp:99-106) but a name lacks the fields chop.hpp:14-70 dereference */
    RAFTGPU_E_PEER = -14         /* sharded run: this rank is fine, another rank reported an error */
} raftgpu_status;

/* Output streams (files the reference writes: chop.hpp:333, repeat.hpp:85-87). */
enum { RAFTGPU_OUT_COVERAGE = 0, RAFTGPU_OUT_LONG_REPEATS = 1, RAFTGPU_OUT_BED = 2, RAFTGPU_OUT_READS_FASTA = 3,
       RAFTGPU_OUT_SPLIT_NAIVE = 4 /* only after raftgpu_split_naive */ };

/* Integer tables for tests / downstream tools (raftgpu_fetch_table). */
enum {
    RAFTGPU_TAB_QID = 0, RAFTGPU_TAB_TID, RAFTGPU_TAB_QS, RAFTGPU_TAB_QE, RAFTGPU_TAB_TS, RAFTGPU_TAB_TE, /* int32[n_records] */
    RAFTGPU_TAB_STRAND,    /* uint8[n_records] */
    RAFTGPU_TAB_BIN_OFF,   /* int64[n_reads+1]: CSR offsets of bins (nb_i = ceil(L_i/reso)) */
    RAFTGPU_TAB_COV,       /* int32[n_bins], CSR by RAFTGPU_TAB_BIN_OFF */
    RAFTGPU_TAB_REP_OFF,   /* int64[n_reads+1] */
    RAFTGPU_TAB_REP,       /* int32[2*n_repeats]: (start,end) pairs in read order */
    RAFTGPU_TAB_FRAG,      /* int32[3*n_fragments]: (read,a,b) in output order */
    RAFTGPU_TAB_COUNT
};

/* What the reference prints on stdout (chop.hpp:105,189-190; repeat.hpp:91,173-178) plus sizes. */
typedef struct {
    int64_t  n_reads;
    int64_t  n_records;        /* "INFO, length of alignments  %d()" */
    int32_t  symmetric;        /* "INFO, Symmetric overlaps %d " */
    int32_t  high_cov;         /* "high_cov %d" */
    int32_t  real_reads;       /* "Real Reads %d " */
    int32_t  total_windows;    /* int in the reference (repeat.hpp:95): wraps past 2^31 */
    int64_t  total_cov;
    int64_t  total_repeat_len; /* unclamped run lengths (repeat.hpp:127,152) */
    int64_t  total_read_len;
    int64_t  n_bins;
    int64_t  n_repeats;
    int64_t  n_fragments;
    uint64_t out_bytes[4];     /* indexed by RAFTGPU_OUT_* */
    /* device time of the stages of the last run, CUDA events on the context stream (ms) */
    float    ms_tokenize, ms_scatter, ms_scan, ms_repeat_cut, ms_layout, ms_total;
    int32_t  kernel_launches;  /* kernels launched by the library since raftgpu_reset */
    int32_t  reserved;
    /* emitters (raftgpu_fetch / raftgpu_digest): accumulated device time of the emit kernels per stream,
     * CUDA events on the launching stream, and how many emit kernels / stream bytes that covers */
    float    ms_emit[4];
    int32_t  emit_launches[4];
    uint64_t emit_bytes[4];
    float    ms_set_reads;     /* device time of raftgpu_set_reads (copies, layout scans, name table) */
    int32_t  reserved2;
} raftgpu_stats;

/* ---- lifetime ---------------------------------------------------------------------------- */
void        raftgpu_default_params(raftgpu_params *p);            /* algoParams::initParams, param.hpp:18-31 */
int         raftgpu_create(const raftgpu_params *p, int device, raftgpu_ctx **out);
int         raftgpu_destroy(raftgpu_ctx *ctx);
int         raftgpu_reset(raftgpu_ctx *ctx);                      /* drop reads, records and outputs; keep allocations */
const char *raftgpu_strerror(int status);
const char *raftgpu_last_error(const raftgpu_ctx *ctx);           /* detail of the last failure (e.g. CUDA error string) */
int64_t     raftgpu_error_index(const raftgpu_ctx *ctx);          /* record / read index the last data error points at, or -1 */

/* Options (raftgpu_set_option).
 * RAFTGPU_OPT_DEFER_SEQ_UPLOAD (default 0): with 1, a HOST `seq` passed to raftgpu_set_reads[_sharded] is not
 * copied inside that call; it is uploaded in chunks on a copy stream once the PAF is complete, overlapping the
 * kernels and the device->host transfer of the outputs (PCIe full duplex).  The host buffer must then stay valid
 * (and should be pinned) until the reads.fasta bytes have been fetched, or raftgpu_reset / raftgpu_destroy. */
enum { RAFTGPU_OPT_DEFER_SEQ_UPLOAD = 1 };
int raftgpu_set_option(raftgpu_ctx *ctx, int option, int64_t value);

/* ---- a0: reads.  Replaces loadFASTA + addStringToMap (chop.hpp:73-131) once the FASTA is tokenised.
 * seq_off/name_off have n+1 entries; names are the header up to the first whitespace (kseq.h:254);
 * ids are the array order (chop.hpp:108).  seq may be NULL (lengths only: tables and coverage /
 * long_repeats text still work, reads.fasta does not).  [host|device] */
int raftgpu_set_reads(raftgpu_ctx *ctx, int64_t n, const int64_t *seq_off, const uint8_t *seq,
                      const int64_t *name_off, const uint8_t *names);

/* Device FASTA / FASTQ ingest: the same records as loadFASTA / kseq_read (chop.hpp:88-131,
 * kseq.h:240-298) cut out of plain FASTA text (any line wrapping) or strict four-line FASTQ text by a CUDA kernel.  Call with consecutive chunks of the (inflated) file, chunks need not end
 * on newlines, last_chunk=1 on the final one (which also builds the layout and the name table, like
 * raftgpu_set_reads).  total_bytes_hint = size of the whole file if known (sizes the arena once), else 0.
 * Returns RAFTGPU_E_UNSUPPORTED — and leaves the context reset — for input the kernel does not take: FASTQ with
 * wrapped or truncated sequence / quality lines, a '+' line in FASTA, CR LF line ends, text that does not start with
 * '>' or '@'; use raftgpu_load_fasta + raftgpu_set_reads then.  [host|device] */
int raftgpu_ingest_fasta(raftgpu_ctx *ctx, const uint8_t *text, size_t nbytes, int last_chunk, uint64_t total_bytes_hint);

/* Host FASTA/FASTQ(+gz) reader with kseq_read semantics (kseq.h:240-298); arrays are malloc'd and
 * owned by the caller (raftgpu_free_host). */
int  raftgpu_load_fasta(const char *path, int64_t *n, int64_t **seq_off, uint8_t **seq, int64_t **name_off,
                        uint8_t **names);
void raftgpu_free_host(void *p);

/* ---- a1/a2: PAF ingest.  Replaces paf_open/paf_read/paf_parse (paf.hpp:24-99) and the id decode +
 * symmetric-overlap detection of create_pileup (chop.hpp:133-191).  May be called repeatedly with
 * consecutive chunks of the (inflated) file; chunks need not end on newlines; pass last_chunk=1 on
 * the final one.  [host|device] */
int raftgpu_ingest_paf(raftgpu_ctx *ctx, const uint8_t *text, size_t nbytes, int last_chunk);

/* ---- a3-a5: coverage, repeats, cut points.  Replaces profileCoverage + repeat_annotate
 * (repeat.hpp:28-204, minus file writing) and the boundary arithmetic of break_reads (chop.hpp:198-323). */
int raftgpu_run(raftgpu_ctx *ctx, raftgpu_stats *stats);
/* Current counters (stage times, emit times, kernel launches) without running anything. */
int raftgpu_get_stats(raftgpu_ctx *ctx, raftgpu_stats *stats);

/* ---- outputs: the bytes of prefix.coverage.txt / .long_repeats.txt / .long_repeats.bed /
 * .reads.fasta (repeat.hpp:105-108,180-203; chop.hpp:250-322), materialised on the device window
 * by window.  dst [host|device]. */
int raftgpu_output_size(raftgpu_ctx *ctx, int which, uint64_t *nbytes);
int raftgpu_fetch(raftgpu_ctx *ctx, int which, uint64_t off, uint8_t *dst, size_t n);
/* Asynchronous form for device destinations: the emitter is queued and the call returns; raftgpu_sync waits for
 * everything queued and reports errors.  The text streams (coverage, long_repeats, bed) and the sequence streams
 * (reads.fasta, split_naive) run on two different CUDA streams, so a caller that queues both lets the issue-bound text
 * formatter overlap the bandwidth-bound gather.  Destinations of calls queued on the same stream may alias. */
int raftgpu_fetch_async(raftgpu_ctx *ctx, int which, uint64_t off, uint8_t *dst_device, size_t n);
int raftgpu_sync(raftgpu_ctx *ctx);
/* Order-independent 64-bit digest of a whole output stream computed on the device
 * (sum over bytes of mix64(offset*257 + byte + 1)); used for parity at sizes the host cannot hold. */
int raftgpu_digest(raftgpu_ctx *ctx, int which, uint64_t *digest);
/* Same with the context's stream placed at byte `stream_base` of a larger file: a rank of a sharded run digests its
 * slice at its file offset, and the per-rank digests add up (mod 2^64) to the digest of the whole file. */
int raftgpu_digest_at(raftgpu_ctx *ctx, int which, uint64_t stream_base, uint64_t *digest);
/* Copies an integer table to dst [host|device]; *n_elems receives the element count (call with
 * dst=NULL to size). */
int raftgpu_fetch_table(raftgpu_ctx *ctx, int table, void *dst, size_t cap_bytes, size_t *n_elems);

/* ---- split_naive (split_naive.cpp:10-44): every read cut into consecutive, non-overlapping pieces of
 * `subread_length` bases, records ">" name "_" k "\n" bases "\n" (k from 1; an empty read gives no record).  Needs reads
 * with sequence bytes (raftgpu_set_reads / raftgpu_ingest_fasta); independent of the PAF.  The bytes are then available
 * as stream RAFTGPU_OUT_SPLIT_NAIVE through raftgpu_output_size / raftgpu_fetch / raftgpu_digest (same gather kernel). */
int raftgpu_split_naive(raftgpu_ctx *ctx, int32_t subread_length);

/* ---- multi-GPU inside the library: one context per GPU ("rank"), reads partitioned into contiguous id ranges
 * (bounds[nranks+1]), every rank holding all names + lengths (raftgpu_set_reads_sharded) and its own share of the PAF
 * text (any line-complete part of the file; the rank holding the file's first line must be rank 0 or the ranks
 * before it must hold no record).  The ranks may be threads of one process (the `raft` CLI with RAFT_B200_DEVICES)
 * or processes (bench.py under torchrun): they share an NCCL communicator built from one 128-byte id.
 *
 * raftgpu_run_sharded is raftgpu_ingest_paf + raftgpu_run for a rank.  Its exchange steps, all on the context's
 * stream, NCCL over NVLink: record 0 of the file (chop.hpp:171-184) = all-gather of every rank's first record;
 * symmetric flag = all-reduce(max); each contributing interval on a read another rank owns (chop.hpp:165-169)
 * = one 12-byte endpoint, counts all-gathered, endpoints exchanged with grouped ncclSend / ncclRecv; global `read=`
 * numbering (chop.hpp:195,266,319) = all-gather of fragment counts; file offsets of the rank's output slices =
 * all-gather of output sizes.  Afterwards raftgpu_output_size / raftgpu_fetch / raftgpu_digest_at serve this rank's
 * slice of every output file; info->stream_base[w] is where it starts in file w. */
#define RAFTGPU_COMM_ID_BYTES 128
typedef struct {
    int32_t  nranks, rank;
    int32_t  symmetric, reserved;
    int64_t  n_records_total, n_fragments_total, first_read_num;
    int64_t  endpoints_sent, endpoints_received; /* 12-byte endpoints this rank sent to / received from other ranks */
    uint64_t stream_base[4], stream_total[4];    /* indexed by RAFTGPU_OUT_*: file offset of this rank's slice, whole file size */
    int64_t  total_cov, total_repeat_len, total_read_len, n_bins_total; /* sums over ranks, for the stdout lines (repeat.hpp:173-178) */
    float    ms_exchange;                        /* device time of route count + pack + send/recv + scatter of the received endpoints */
    int32_t  peek_retries;
} raftgpu_shard_info;
int raftgpu_comm_unique_id(uint8_t id[RAFTGPU_COMM_ID_BYTES]);
/* collective over the ranks: every rank calls it with the same id and nranks, its own rank */
int raftgpu_comm_init(raftgpu_ctx *ctx, int nranks, int rank, const uint8_t id[RAFTGPU_COMM_ID_BYTES]);
int raftgpu_comm_destroy(raftgpu_ctx *ctx);
int raftgpu_run_sharded(raftgpu_ctx *ctx, const int64_t *bounds /* nranks+1 */, const uint8_t *text /* [host|device] */, size_t nbytes,
                        raftgpu_stats *stats, raftgpu_shard_info *info);
/* The same in three steps for a rank whose text arrives in chunks (files larger than memory): begin with the first chunk
 * (record 0 of the file is looked for in it; head_is_whole_text=1 when it is the rank's only chunk), raftgpu_ingest_paf
 * for every chunk including the first, then finish.  A data error of an ingest call is carried into finish so that every
 * rank still meets the others in the collectives (the failing rank gets its own code back, the others RAFTGPU_E_PEER). */
int raftgpu_sharded_begin(raftgpu_ctx *ctx, const int64_t *bounds, const uint8_t *head_text, size_t head_bytes, int head_is_whole_text);
int raftgpu_sharded_finish(raftgpu_ctx *ctx, raftgpu_stats *stats, raftgpu_shard_info *info);
/* Reads held by a context (after raftgpu_ingest_fasta / raftgpu_set_reads[_sharded]): how many it owns, then host copies of
 * their names (offsets relative to the first owned name) and lengths, and the device pointers of its sequence arena --
 * what a multi-GPU caller needs to turn per-rank FASTA ingests into one raftgpu_set_reads_sharded per rank. */
int raftgpu_reads_info(raftgpu_ctx *ctx, int64_t *n_local, int64_t *name_bytes, int64_t *seq_bytes);
int raftgpu_reads_copy(raftgpu_ctx *ctx, int64_t *name_off /* n_local+1 */, uint8_t *names, int64_t *lengths /* n_local */);
int raftgpu_reads_device(raftgpu_ctx *ctx, const int64_t **seq_off_device, const uint8_t **seq_device);
/* Read-id range boundaries balanced by coverage slots (ceil(L/reso) + 1 per read): bounds[nranks+1]. lengths [host]. */
int raftgpu_partition_reads(const int64_t *lengths, int64_t n, int32_t reso, int nranks, int64_t *bounds);

/* ---- multi-GPU building blocks for callers that run the exchange themselves (one context per rank).  Reads are
 * partitioned into contiguous id ranges; every rank holds all names and lengths, and sequence bytes for its own
 * range only. */
int raftgpu_set_reads_sharded(raftgpu_ctx *ctx, int64_t n, const int64_t *lengths /* int64[n] */,
                              const int64_t *name_off, const uint8_t *names, int64_t own_first,
                              int64_t own_count, const int64_t *own_seq_off /* own_count+1, local */,
                              const uint8_t *own_seq);
/* First record of the whole PAF (rank that holds byte 0): six ints qid,tid,qs,qe,ts,te; found=0 if
 * this rank's text has no record.  chop.hpp:171-184 compares every later record with it. */
int raftgpu_peek_first_record(raftgpu_ctx *ctx, const uint8_t *text, size_t nbytes, int32_t rec[6], int32_t *found);
int raftgpu_set_first_record(raftgpu_ctx *ctx, const int32_t rec[6], int32_t is_local);
int raftgpu_get_symmetric(raftgpu_ctx *ctx, int32_t *flag);  /* local OR over this rank's records */
int raftgpu_set_symmetric(raftgpu_ctx *ctx, int32_t flag);   /* global value after the all-reduce */
/* Add the intervals of this rank's own records that fall on reads it owns (no exchange needed for those). */
int raftgpu_accumulate_local(raftgpu_ctx *ctx);
/* Count / pack the 12-byte endpoint records (global read id, start, end) this rank must send to the
 * OTHER owners (reads owned by this rank are covered by raftgpu_accumulate_local; its own count is 0);
 * bounds[nranks+1] are the read-id range boundaries.  counts are int64[nranks]; sendbuf is device
 * memory of at least 12*sum(counts) bytes, laid out by destination rank. */
int raftgpu_route_count(raftgpu_ctx *ctx, int nranks, const int64_t *bounds, int64_t *counts);
int raftgpu_route_pack(raftgpu_ctx *ctx, int nranks, const int64_t *bounds, const int64_t *counts, void *sendbuf_device);
/* Add routed endpoints (device memory, 12 bytes each) owned by this rank into its coverage. */
int raftgpu_accumulate_endpoints(raftgpu_ctx *ctx, const void *endpoints_device, int64_t count);
/* raftgpu_run split in two for the exchange: finalize computes coverage/repeats/fragments for the
 * owned reads from what was accumulated; set_output_base then fixes the global read= numbering and
 * read index of this rank's first fragment / read before outputs are sized. */
int raftgpu_finalize(raftgpu_ctx *ctx, raftgpu_stats *stats);
int raftgpu_set_output_base(raftgpu_ctx *ctx, int64_t first_read_num);

/* ---- file-level drop-in for break_long_reads(reads, paf, unused, param) (chop.hpp:331-373):
 * reads both files, runs the path on `device`, writes prefix.{coverage.txt,long_repeats.txt,
 * long_repeats.bed,reads.fasta}, and prints the reference's stdout lines. */
int raftgpu_break_long_reads(const char *readfilename, const char *paffilename, const raftgpu_params *p,
                             const char *prefix, int device, raftgpu_stats *stats);
/* Same with several PAF files read back to back, as `cat a.paf b.paf` would feed them (README.md:32-38:
 * hifiasm writes <prefix>.0.ovlp.paf and <prefix>.1.ovlp.paf); gzip or plain, streamed through two pinned
 * buffers by a reader thread so file reading overlaps the device tokenizer.  n_paf >= 1. */
int raftgpu_break_long_reads_multi(const char *readfilename, int n_paf, const char *const *paffilenames,
                                   const raftgpu_params *p, const char *prefix, int device, raftgpu_stats *stats_out);
/* Same over several GPUs of this box (one host thread + context per device, raftgpu_run_sharded underneath): reads
 * sharded by id range, every PAF file cut into byte ranges at line ends, each rank writing its slice of the four output
 * files at its file offset.  Files and stdout lines are identical to the single-GPU call. */
int raftgpu_break_long_reads_mgpu(const char *readfilename, int n_paf, const char *const *paffilenames, const raftgpu_params *p,
                                  const char *prefix, int ndev, const int *devices, raftgpu_stats *stats_out);

#ifdef __cplusplus
}
#endif
#endif
