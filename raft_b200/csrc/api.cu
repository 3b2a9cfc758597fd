// api.cu — context, stage sequencing and the extern "C" boundary of libraft_b200.so
// (declared in include/raft_b200.h).  Host logic only; all compute is in the k*.cu kernels.
#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nccl.h>

#include "../../include/raft_b200.h"
#include "kernels.h"

using namespace raftk;

namespace {

struct DevBuf {
    void*  p = nullptr;
    size_t cap = 0;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    // grow to at least `bytes` (contents NOT preserved unless keep > 0 bytes)
    cudaError_t ensure(size_t bytes, size_t keep = 0, cudaStream_t st = 0)
    {
        if (bytes <= cap) return cudaSuccess;
        size_t want = bytes + bytes / 8 + 256;
        void*  np = nullptr;
        cudaError_t e = cudaMalloc(&np, want);
        if (e != cudaSuccess) { e = cudaMalloc(&np, bytes + 256); want = bytes + 256; }
        if (e != cudaSuccess) return e;
        if (keep && p) { e = cudaMemcpyAsync(np, p, keep, cudaMemcpyDeviceToDevice, st); if (e == cudaSuccess) e = cudaStreamSynchronize(st); }
        if (p) cudaFree(p);
        p = np; cap = want;
        return e;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

bool is_device_ptr(const void* p)
{
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// misc device scalars
struct Misc {
    int                rec0[8];
    int                sym_flag;
    int                ticket;
    int                work_counter;
    int                pad;
    ErrState           err;
    ErrState           err_range;
    long long          n_records_out;
    unsigned long long stats[2];
    unsigned long long digest;
    unsigned long long route_counts[64];
    unsigned long long route_list_n;
    int                fa_flags;
    int                fa_pad;
    long long          fa_totals[3];
    long long          zero;     // stays 0 (a size that does not exist on this run, for the packed all-gathers)
};

constexpr size_t WINDOW_BYTES = 256ull << 20; // staging window for host fetches

} // namespace

struct raftgpu_ctx {
    int            device = 0;
    cudaStream_t   st = nullptr, st2 = nullptr, st_zero = nullptr;
    cudaEvent_t    ev_scan_done = nullptr;
    void*          diff_zero_ptr = nullptr; // difference array known to be all zero once st_zero has drained ...
    int64_t        diff_zero_ints = 0;      // ... over this many leading ints
    raftgpu_params prm{};
    std::string    last_error;
    int64_t        err_index = -1;
    int            launches = 0;

    // reads
    bool    have_reads = false, have_seq = false;
    int64_t n = 0, own_first = 0, m = 0;
    int     real_reads = 1;
    const int64_t* d_name_off = nullptr; const uint8_t* d_names = nullptr;
    const int64_t* d_seq_off = nullptr;  const uint8_t* d_seq = nullptr;
    DevBuf  b_name_off, b_names, b_seq_off, b_seq;
    DevBuf  b_slots, b_repcap, b_cutcap, b_slot_off, b_rep_cap_off, b_cut_cap_off;
    int64_t n_slots = 0, rep_cap_total = 0, cut_cap_total = 0, total_read_len = 0;
    DevBuf  b_table; NameTable nt{};

    // records
    DevBuf  b_qid, b_tid, b_qs, b_qe, b_ts, b_te, b_strand;
    int64_t rec_cap = 0, n_rec = 0;
    DevBuf  b_text;
    std::vector<uint8_t> carry;
    int     first_is_local = 1; // 1 / 0, or -1: the tokenizer reads it from rec0[7] on the device (sharded runs)
    bool    rec0_external = false, sym_external = false, paf_done = false;
    bool    q_scattered = false; // query sides went into the difference array during tokenisation
    int     h_sym = 0;           // host copy of the (local) symmetric flag after the last ingest
    int64_t paf_bytes = 0;

    DevBuf  b_misc, b_status;
    Misc*   misc() const { return b_misc.as<Misc>(); }

    // coverage + K3 + layout
    DevBuf  b_cov, b_diff; bool diff_zeroed = false, finalized = false, sized = false;
    DevBuf  b_bounds, b_route_list; int64_t route_collected = -1; // endpoints in b_route_list (-1: list incomplete)
    DevBuf  b_rep, b_rep_cnt, b_cuts, b_frag_cnt, b_frag_base;
    DevBuf  b_frag_read, b_frag_a, b_frag_b, b_frag_size, b_frag_off;
    DevBuf  b_rep_off, b_rep_line_size, b_rep_line_off, b_cov_tile_bytes, b_cov_tile_static, b_cov_tile_off, b_cov_tile_read, b_fasta_tile_frag, b_frag_desc;
    std::vector<int64_t> h_cov_tile_off, h_rep_line_off; // every OFF_SAMPLE-th entry of the device tables (+ the last)
    DevBuf  b_off_sample, b_sim, b_bed_line_size, b_bed_line_off;
    DevBuf  b_cov_tabs; bool cov_tabs_built = false; // text tables of the coverage.txt emitter (depend on -r only)
    std::vector<int64_t> h_bed_line_off;
    int64_t G = 0, n_repeats = 0, read_num_base = 0;
    raftgpu_stats stats{};
    DevBuf  b_stage[2];
    cudaEvent_t ev[10]{};
    cudaEvent_t ev_stage[2]{};
    // two emit lanes so that the issue-bound text emitters and the bandwidth-bound gather can overlap:
    // lane 0 (stream st_aux): coverage.txt, long_repeats.txt, .bed;  lane 1 (stream st): reads.fasta, split_naive
    struct EmitLane { cudaStream_t st = nullptr; cudaEvent_t ev0 = nullptr, ev1 = nullptr; bool open = false; int kind = 0; };
    EmitLane     lane[2];
    cudaStream_t st_aux = nullptr;

    // split_naive stream (raftgpu_split_naive)
    int      sn_len = 0;
    int64_t  sn_G = 0;
    uint64_t sn_bytes = 0;
    DevBuf   b_sn_cnt, b_sn_base, b_sn_read, b_sn_a, b_sn_b, b_sn_size, b_sn_off, b_sn_desc, b_sn_tile;

    // device FASTA ingest (raftgpu_ingest_fasta)
    bool    fasta_active = false;
    int64_t fa_n = 0, fa_bases = 0, fa_name_bytes = 0, fa_rec_cap = 0;
    bool    fa_fastq = false;                         // strict four-line FASTQ mode of the device tokenizer
    int64_t fa_lines = 0, fa_text_bytes = 0, fa_gcap = 0; // lines / text bytes of the chunks so far; capacity of the per-record file offsets
    bool    fa_ends_nl = true, fa_mode_known = false; // the FASTA / FASTQ decision needs the first three lines
    std::vector<uint8_t> fa_carry;
    DevBuf  b_fa_text, b_fa_rec_pos, b_fa_name_len, b_fa_name_off_chunk, b_fa_status, b_fa_rec_gpos, b_fa_qual_gpos;

    // deferred sequence upload (RAFTGPU_OPT_DEFER_SEQ_UPLOAD)
    bool           opt_defer_seq = false;
    cudaStream_t   st_h2d = nullptr;
    const uint8_t* seq_host = nullptr;   // pending host arena
    size_t         seq_host_bytes = 0;
    bool           seq_upload_started = false;
    std::vector<cudaEvent_t> ev_chunk;   // one per uploaded chunk
    size_t         n_chunks = 0;
    std::vector<int64_t> h_frag_sample;  // (out_off, src_off) of every FRAG_SAMPLE-th record
    DevBuf         b_frag_sample;

    // sharded runs inside the library (raftgpu_comm_init / raftgpu_run_sharded): NCCL over NVLink
    ncclComm_t     comm = nullptr;
    int            nranks = 1, rank = 0;
    bool           gather_on = false;    // finalize / layout add their all-gathers (set while raftgpu_run_sharded runs)
    DevBuf         b_coll, b_send, b_recv, b_peek;
    long long*     h_coll = nullptr;     // pinned host mirror of the gathered words
    raftgpu_shard_info shard{};
    std::vector<int64_t> shard_bounds;
    bool           shard_begun = false;
    int            pending_ingest_status = 0; // data error of an ingest call between sharded_begin and sharded_finish
};
constexpr int    MAX_RANKS = 64;
// layout of b_coll / h_coll in 8-byte words
constexpr size_t COLL_SEND = 0;                                       // send staging (<= MAX_RANKS + 8 words)
constexpr size_t COLL_REC0 = COLL_SEND + MAX_RANKS + 8;               // gathered rec0: nranks x 8 ints (4 words each)
constexpr size_t COLL_CNT = COLL_REC0 + 4 * MAX_RANKS;                // gathered counts: nranks x (nranks + CNT_EXTRA)
constexpr int    CNT_EXTRA = 6;                                       // list_n, symmetric, n_rec, status, total_read_len, n_bins
constexpr size_t COLL_FIN = COLL_CNT + (size_t)MAX_RANKS * (MAX_RANKS + CNT_EXTRA); // gathered finalize words: nranks x 8
constexpr size_t COLL_OUT = COLL_FIN + 8 * MAX_RANKS;                 // gathered output sizes: nranks x 8
constexpr size_t COLL_WORDS = COLL_OUT + 8 * MAX_RANKS;
#define NCK(call)                                                                                        \
    do {                                                                                                 \
        ncclResult_t r_ = (call);                                                                        \
        if (r_ != ncclSuccess) {                                                                         \
            ctx->last_error = std::string(#call) + ": " + ncclGetErrorString(r_);                        \
            return RAFTGPU_E_CUDA;                                                                       \
        }                                                                                                \
    } while (0)

constexpr int    OFF_SAMPLE = 256;
constexpr size_t SEQ_CHUNK = 256ull << 20;
constexpr int    FRAG_SAMPLE = 256;

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            ctx->last_error = std::string(#call) + ": " + cudaGetErrorString(e_);                        \
            cudaGetLastError();                                                                          \
            return e_ == cudaErrorMemoryAllocation ? RAFTGPU_E_NOMEM : RAFTGPU_E_CUDA;                   \
        }                                                                                                \
    } while (0)
#define CKL() do { ctx->launches++; CK(cudaGetLastError()); } while (0)
#define FAIL(code, msg) do { ctx->last_error = (msg); return (code); } while (0)

static int check_params(const raftgpu_params& p)
{ // the reference divides by these (repeat.hpp:32, chop.hpp:209,248,270) or emits a repeat per bin (repeat.hpp:125 with p<1)
    if (p.reso < 1 || p.repeat_length < 1 || p.interval_length < 1 || p.read_length < p.interval_length) return RAFTGPU_E_PARAM;
    return RAFTGPU_OK;
}

// chop.hpp:99-106: ^read=[0-9]+,[a-z]+,position=[0-9]+-[0-9]+,length=[0-9]+,(.*)
static bool is_simulated_name(const std::string& s)
{
    size_t i = 0;
    auto lit = [&](const char* t) { size_t l = strlen(t); if (s.compare(i, l, t) != 0) return false; i += l; return true; };
    auto plus = [&](char lo, char hi) { size_t j = i; while (i < s.size() && s[i] >= lo && s[i] <= hi) i++; return i > j; };
    return lit("read=") && plus('0', '9') && lit(",") && plus('a', 'z') && lit(",position=") && plus('0', '9') && lit("-") &&
           plus('0', '9') && lit(",length=") && plus('0', '9') && lit(",");
}

extern "C" {

void raftgpu_default_params(raftgpu_params* p)
{
    p->reso = 50; p->est_cov = 0; p->cov_mul = 1.5; p->repeat_length = 10000; p->interval_length = 10000;
    p->read_length = 20000; p->overlap_length = 500; p->flanking_length = 1000;
}

const char* raftgpu_strerror(int s)
{
    switch (s) {
    case RAFTGPU_OK: return "ok";
    case RAFTGPU_E_PARAM: return "invalid parameters (need reso>=1, repeat_length>=1, read_length>=interval_length>=1)";
    case RAFTGPU_E_UNKNOWN_NAME: return "PAF names a read that is not in the reads file";
    case RAFTGPU_E_DUP_NAME: return "duplicate read name";
    case RAFTGPU_E_RANGE: return "overlap interval extends past the end of its read";
    case RAFTGPU_E_NEG_START: return "overlap_length larger than a cut position (fragment would start before 0)";
    case RAFTGPU_E_NOMEM: return "out of memory";
    case RAFTGPU_E_FASTQ: return "truncated FASTQ record";
    case RAFTGPU_E_CUDA: return "CUDA error (no usable sm_100 device?)";
    case RAFTGPU_E_STATE: return "call order violated";
    case RAFTGPU_E_IO: return "input file missing/empty or output not writable";
    case RAFTGPU_E_ARG: return "bad argument";
    case RAFTGPU_E_UNSUPPORTED: return "not supported";
    case RAFTGPU_E_SIM_NAME: return "simulated-read mode (first name matches read=N,align,position=a-b,length=L,chr) but a later name does not";
    case RAFTGPU_E_PEER: return "another rank of the sharded run failed";
    default: return "unknown status";
    }
}
const char* raftgpu_last_error(const raftgpu_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }
int64_t     raftgpu_error_index(const raftgpu_ctx* ctx) { return ctx ? ctx->err_index : -1; }

int raftgpu_create(const raftgpu_params* p, int device, raftgpu_ctx** out)
{
    if (!p || !out) return RAFTGPU_E_ARG;
    *out = nullptr;
    int st = check_params(*p);
    if (st) return st;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return RAFTGPU_E_CUDA; }
    if (cudaSetDevice(device) != cudaSuccess) return RAFTGPU_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) return RAFTGPU_E_CUDA; // sm_100a code only
    raftgpu_ctx* ctx = new raftgpu_ctx();
    ctx->device = device; ctx->prm = *p;
    auto bail = [&](void) { delete ctx; return RAFTGPU_E_CUDA; };
    if (cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking) != cudaSuccess) return bail();
    if (cudaStreamCreateWithFlags(&ctx->st2, cudaStreamNonBlocking) != cudaSuccess) return bail();
    if (cudaStreamCreateWithFlags(&ctx->st_zero, cudaStreamNonBlocking) != cudaSuccess) return bail();
    if (cudaEventCreateWithFlags(&ctx->ev_scan_done, cudaEventDisableTiming) != cudaSuccess) return bail();
    for (auto& e : ctx->ev) if (cudaEventCreate(&e) != cudaSuccess) return bail();
    for (auto& e : ctx->ev_stage) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return bail();
    if (cudaStreamCreateWithFlags(&ctx->st_aux, cudaStreamNonBlocking) != cudaSuccess) return bail();
    ctx->lane[0].st = ctx->st_aux; ctx->lane[1].st = ctx->st;
    for (auto& L : ctx->lane) if (cudaEventCreate(&L.ev0) != cudaSuccess || cudaEventCreate(&L.ev1) != cudaSuccess) return bail();
    if (cudaStreamCreateWithFlags(&ctx->st_h2d, cudaStreamNonBlocking) != cudaSuccess) return bail();
    if (ctx->b_misc.ensure(sizeof(Misc)) != cudaSuccess) return bail();
    *out = ctx;
    return raftgpu_reset(ctx);
}

int raftgpu_destroy(raftgpu_ctx* ctx)
{
    if (!ctx) return RAFTGPU_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st); cudaStreamSynchronize(ctx->st2);
    if (ctx->st_zero) { cudaStreamSynchronize(ctx->st_zero); cudaStreamDestroy(ctx->st_zero); }
    if (ctx->ev_scan_done) cudaEventDestroy(ctx->ev_scan_done);
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    for (auto& e : ctx->ev_stage) if (e) cudaEventDestroy(e);
    for (auto& L : ctx->lane) { if (L.ev0) cudaEventDestroy(L.ev0); if (L.ev1) cudaEventDestroy(L.ev1); }
    if (ctx->st_aux) { cudaStreamSynchronize(ctx->st_aux); cudaStreamDestroy(ctx->st_aux); }
    if (ctx->st_h2d) { cudaStreamSynchronize(ctx->st_h2d); cudaStreamDestroy(ctx->st_h2d); }
    for (auto& e : ctx->ev_chunk) if (e) cudaEventDestroy(e);
    if (ctx->comm) { ncclCommDestroy(ctx->comm); ctx->comm = nullptr; }
    if (ctx->h_coll) cudaFreeHost(ctx->h_coll);
    cudaStreamDestroy(ctx->st); cudaStreamDestroy(ctx->st2);
    delete ctx;
    return RAFTGPU_OK;
}

int raftgpu_reset(raftgpu_ctx* ctx)
{
    if (!ctx) return RAFTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (ctx->st_h2d) CK(cudaStreamSynchronize(ctx->st_h2d)); // an upload in flight still reads the caller's host arena
    ctx->seq_host = nullptr; ctx->seq_host_bytes = 0; ctx->seq_upload_started = false; ctx->n_chunks = 0;
    Misc h{};
    h.err.packed = ERR_CLEAN; h.err_range.packed = ERR_CLEAN;
    CK(cudaMemcpyAsync(ctx->b_misc.p, &h, sizeof h, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->have_reads = ctx->have_seq = false; ctx->n = ctx->m = ctx->own_first = 0;
    ctx->n_rec = 0; ctx->carry.clear(); ctx->first_is_local = 1; ctx->rec0_external = ctx->sym_external = false;
    ctx->paf_done = false; ctx->paf_bytes = 0; ctx->diff_zeroed = ctx->finalized = ctx->sized = false;
    ctx->q_scattered = false; ctx->h_sym = 0;
    ctx->shard_begun = false; ctx->pending_ingest_status = 0; ctx->gather_on = false;
    ctx->fasta_active = false; ctx->fa_n = ctx->fa_bases = ctx->fa_name_bytes = 0; ctx->fa_carry.clear();
    ctx->fa_fastq = false; ctx->fa_lines = ctx->fa_text_bytes = 0; ctx->fa_ends_nl = true; ctx->fa_mode_known = false;
    ctx->sn_len = 0; ctx->sn_G = 0; ctx->sn_bytes = 0;
    ctx->G = ctx->n_repeats = ctx->read_num_base = 0; ctx->launches = 0; ctx->err_index = -1;
    ctx->stats = raftgpu_stats{};
    for (auto& L : ctx->lane) { if (L.open) cudaStreamSynchronize(L.st); L.open = false; }
    return RAFTGPU_OK;
}

} // extern "C"

// ------------------------------------------------------------------------------------------------
static int fetch_err(raftgpu_ctx* ctx)
{ // read the device error word (stream must be synchronised by the caller or here)
    ErrState e2[2];
    CK(cudaMemcpyAsync(e2, &ctx->misc()->err, sizeof e2, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ErrState e = e2[0].packed != ERR_CLEAN ? e2[0] : e2[1]; // unknown names / duplicates first, then range errors of the fused scatter
    if (e.packed != ERR_CLEAN) {
        ctx->err_index = err_index(e);
        char buf[160];
        snprintf(buf, sizeof buf, "%s (index %lld)", raftgpu_strerror(err_code(e)), err_index(e));
        ctx->last_error = buf;
        return err_code(e);
    }
    return RAFTGPU_OK;
}

// copy `bytes` from src (host or device) into an owned device buffer, or borrow a device pointer
template <typename T>
static int adopt(raftgpu_ctx* ctx, DevBuf& buf, const T* src, size_t count, const T** out, size_t pad_bytes = 32)
{
    size_t bytes = count * sizeof(T);
    if (src && is_device_ptr(src) && ((uintptr_t)src & 15) == 0) { *out = src; return RAFTGPU_OK; }
    CK(buf.ensure(bytes + pad_bytes));
    if (bytes) CK(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyDefault, ctx->st));
    CK(cudaMemsetAsync((uint8_t*)buf.p + bytes, 0, pad_bytes, ctx->st));
    *out = buf.as<T>();
    return RAFTGPU_OK;
}

// sequence arena: device pointers are borrowed, host pointers are copied now or (option) uploaded later in chunks
static int adopt_seq(raftgpu_ctx* ctx, const uint8_t* seq, size_t bytes)
{
    if (!ctx->opt_defer_seq || is_device_ptr(seq) || bytes == 0) return adopt(ctx, ctx->b_seq, seq, bytes, &ctx->d_seq);
    CK(ctx->b_seq.ensure(bytes + 32));
    CK(cudaMemsetAsync((uint8_t*)ctx->b_seq.p + bytes, 0, 32, ctx->st));
    ctx->d_seq = ctx->b_seq.as<uint8_t>();
    ctx->seq_host = seq; ctx->seq_host_bytes = bytes; ctx->seq_upload_started = false;
    return RAFTGPU_OK;
}

// enqueue the chunked upload of a deferred host arena on the copy stream (no-op otherwise)
static int start_seq_upload(raftgpu_ctx* ctx)
{
    if (!ctx->seq_host || ctx->seq_upload_started) return RAFTGPU_OK;
    ctx->n_chunks = (ctx->seq_host_bytes + SEQ_CHUNK - 1) / SEQ_CHUNK;
    while (ctx->ev_chunk.size() < ctx->n_chunks) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->ev_chunk.push_back(e);
    }
    for (size_t c = 0; c < ctx->n_chunks; c++) {
        size_t off = c * SEQ_CHUNK, len = std::min(SEQ_CHUNK, ctx->seq_host_bytes - off);
        CK(cudaMemcpyAsync((uint8_t*)ctx->b_seq.p + off, ctx->seq_host + off, len, cudaMemcpyHostToDevice, ctx->st_h2d));
        CK(cudaEventRecord(ctx->ev_chunk[c], ctx->st_h2d));
    }
    ctx->seq_upload_started = true;
    return RAFTGPU_OK;
}

// make stream `st` wait until arena bytes [0, need_end) have arrived
static int wait_seq_bytes(raftgpu_ctx* ctx, int64_t need_end, cudaStream_t st)
{
    if (!ctx->seq_host) return RAFTGPU_OK;
    int rc = start_seq_upload(ctx);
    if (rc) return rc;
    if (need_end <= 0 || ctx->n_chunks == 0) return RAFTGPU_OK;
    size_t c = std::min<size_t>(ctx->n_chunks - 1, (size_t)((need_end - 1) / (int64_t)SEQ_CHUNK));
    CK(cudaStreamWaitEvent(st, ctx->ev_chunk[c], 0));
    return RAFTGPU_OK;
}

static int build_layout_and_names(raftgpu_ctx* ctx, const std::string& first_name)
{
    const raftgpu_params& P = ctx->prm;
    ctx->real_reads = is_simulated_name(first_name) ? 0 : 1; // chop.hpp:99-106 (first record only)
    const int64_t m = ctx->m;
    // per-read slot / capacity layout (three exclusive scans)
    CK(ctx->b_slots.ensure(sizeof(int32_t) * (m + 1))); CK(ctx->b_repcap.ensure(sizeof(int32_t) * (m + 1)));
    CK(ctx->b_cutcap.ensure(sizeof(int32_t) * (m + 1)));
    CK(ctx->b_slot_off.ensure(sizeof(int64_t) * (m + 2))); CK(ctx->b_rep_cap_off.ensure(sizeof(int64_t) * (m + 1)));
    CK(ctx->b_cut_cap_off.ensure(sizeof(int64_t) * (m + 1)));
    CK(ctx->b_status.ensure(sizeof(uint64_t) * (scan_tiles_small(m) + 8)));
    launch_read_layout(ctx->d_seq_off, m, P.reso, P.repeat_length, P.interval_length, P.read_length, ctx->b_slots.as<int32_t>(),
                       ctx->b_repcap.as<int32_t>(), ctx->b_cutcap.as<int32_t>(), ctx->st);
    CKL();
    launch_scan_i32_to_i64(ctx->b_slots.as<int32_t>(), ctx->b_slot_off.as<int64_t>(), m, ctx->b_status.as<uint64_t>(), &ctx->misc()->ticket, ctx->st);
    CKL();
    {   // one entry past the end: the coverage.txt emitter steps to "the read after the last one" without a bounds test
        static const int64_t past_end = (int64_t)1 << 62;
        CK(cudaMemcpyAsync(ctx->b_slot_off.as<int64_t>() + m + 1, &past_end, sizeof past_end, cudaMemcpyHostToDevice, ctx->st));
    }
    launch_scan_i32_to_i64(ctx->b_repcap.as<int32_t>(), ctx->b_rep_cap_off.as<int64_t>(), m, ctx->b_status.as<uint64_t>(), &ctx->misc()->ticket, ctx->st);
    CKL();
    launch_scan_i32_to_i64(ctx->b_cutcap.as<int32_t>(), ctx->b_cut_cap_off.as<int64_t>(), m, ctx->b_status.as<uint64_t>(), &ctx->misc()->ticket, ctx->st);
    CKL();
    int64_t tot[3], seq_total = 0;
    CK(cudaMemcpyAsync(&tot[0], ctx->b_slot_off.as<int64_t>() + m, 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    {   // first read of every 1024-slot tile of coverage.txt (used by the fused sizing in the scan and by the emitter)
        const int T = cov_tiles(tot[0]);
        CK(ctx->b_cov_tile_read.ensure(sizeof(int32_t) * (size_t)(T + 2)));
        CK(ctx->b_cov_tile_bytes.ensure(sizeof(int32_t) * (size_t)(T + 1)));
        CK(ctx->b_cov_tile_static.ensure(sizeof(int32_t) * (size_t)(T + 1)));
        launch_cov_tile_index(ctx->b_slot_off.as<int64_t>(), m, tot[0], ctx->b_cov_tile_read.as<int32_t>(), ctx->st);
        CKL();
        launch_cov_static_sizes(ctx->b_slot_off.as<int64_t>(), m, tot[0], ctx->own_first, P.reso, ctx->b_cov_tile_static.as<int32_t>(), ctx->st);
        CKL();
    }
    CK(cudaMemcpyAsync(&tot[1], ctx->b_rep_cap_off.as<int64_t>() + m, 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(&tot[2], ctx->b_cut_cap_off.as<int64_t>() + m, 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(&seq_total, ctx->d_seq_off + m, 8, cudaMemcpyDeviceToHost, ctx->st));
    // name table: capacity = next pow2 >= 2n, re-seeded on a 64-bit collision
    unsigned long long cap = 16;
    while (cap < 2ull * (unsigned long long)ctx->n) cap <<= 1;
    CK(ctx->b_table.ensure(cap * sizeof(NameSlot)));
    for (unsigned long long seed = 0x5EEDull;; seed = seed * 6364136223846793005ull + 1442695040888963407ull) {
        CK(launch_name_build(&ctx->nt, ctx->b_table.p, cap, seed, ctx->d_names, ctx->d_name_off, ctx->n, &ctx->misc()->err, ctx->st));
        ctx->launches += 3;
        int st = fetch_err(ctx);
        if (st == RAFTK_E_HASH_COLLISION) { // astronomically rare: rebuild with another seed
            ErrState clean{ERR_CLEAN, 0};
            CK(cudaMemcpyAsync(&ctx->misc()->err, &clean, sizeof clean, cudaMemcpyHostToDevice, ctx->st));
            continue;
        }
        if (st) return st;
        break;
    }
    ctx->n_slots = tot[0]; ctx->rep_cap_total = tot[1]; ctx->cut_cap_total = tot[2];
    ctx->total_read_len = seq_total; // owned reads only (summed across ranks by the caller when sharded)
    if (!ctx->real_reads) { // simulated-read names carry genome coordinates (chop.hpp:14-70): parse them once per owned read
        CK(ctx->b_sim.ensure(sizeof(SimInfo) * (size_t)(m + 1)));
        launch_sim_parse(ctx->d_names, ctx->d_name_off, ctx->own_first, m, ctx->b_sim.as<SimInfo>(), &ctx->misc()->err, ctx->st);
        CKL();
        int st = fetch_err(ctx);
        if (st) return st;
    }
    ctx->have_reads = true;
    return RAFTGPU_OK;
}

static int first_name_of(raftgpu_ctx* ctx, int64_t n, const int64_t* name_off, const uint8_t* names, std::string& out)
{
    out.clear();
    if (n <= 0) return RAFTGPU_OK;
    int64_t o[2];
    CK(cudaMemcpy(o, name_off, 16, cudaMemcpyDefault));
    size_t len = (size_t)(o[1] - o[0]);
    out.resize(len);
    if (len) CK(cudaMemcpy(&out[0], names + o[0], len, cudaMemcpyDefault));
    return RAFTGPU_OK;
}

extern "C" int raftgpu_set_reads(raftgpu_ctx* ctx, int64_t n, const int64_t* seq_off, const uint8_t* seq, const int64_t* name_off,
                                 const uint8_t* names)
{
    if (!ctx || n < 0 || !seq_off || !name_off || n > 0x7ffffff0ll) return RAFTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int st = raftgpu_reset(ctx);
    if (st) return st;
    cudaEventRecord(ctx->ev[0], ctx->st);
    int64_t name_bytes = 0, seq_bytes = 0;
    CK(cudaMemcpy(&name_bytes, name_off + n, 8, cudaMemcpyDefault));
    CK(cudaMemcpy(&seq_bytes, seq_off + n, 8, cudaMemcpyDefault));
    ctx->n = n; ctx->m = n; ctx->own_first = 0;
    if ((st = adopt(ctx, ctx->b_name_off, name_off, (size_t)n + 1, &ctx->d_name_off))) return st;
    if ((st = adopt(ctx, ctx->b_names, names, (size_t)name_bytes, &ctx->d_names))) return st;
    if ((st = adopt(ctx, ctx->b_seq_off, seq_off, (size_t)n + 1, &ctx->d_seq_off))) return st;
    ctx->have_seq = seq != nullptr || seq_bytes == 0;
    if (seq) { if ((st = adopt_seq(ctx, seq, (size_t)seq_bytes))) return st; }
    else ctx->d_seq = nullptr;
    std::string first;
    if ((st = first_name_of(ctx, n, name_off, names, first))) return st;
    st = build_layout_and_names(ctx, first);
    cudaEventRecord(ctx->ev[1], ctx->st);
    if (cudaEventSynchronize(ctx->ev[1]) == cudaSuccess) cudaEventElapsedTime(&ctx->stats.ms_set_reads, ctx->ev[0], ctx->ev[1]);
    return st;
}

extern "C" int raftgpu_set_reads_sharded(raftgpu_ctx* ctx, int64_t n, const int64_t* lengths, const int64_t* name_off, const uint8_t* names,
                                         int64_t own_first, int64_t own_count, const int64_t* own_seq_off, const uint8_t* own_seq)
{
    if (!ctx || n < 0 || !name_off || own_first < 0 || own_count < 0 || own_first + own_count > n || n > 0x7ffffff0ll) return RAFTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int st = raftgpu_reset(ctx);
    if (st) return st;
    cudaEventRecord(ctx->ev[0], ctx->st);
    int64_t name_bytes = 0, seq_bytes = 0;
    CK(cudaMemcpy(&name_bytes, name_off + n, 8, cudaMemcpyDefault));
    ctx->n = n; ctx->m = own_count; ctx->own_first = own_first;
    if ((st = adopt(ctx, ctx->b_name_off, name_off, (size_t)n + 1, &ctx->d_name_off))) return st;
    if ((st = adopt(ctx, ctx->b_names, names, (size_t)name_bytes, &ctx->d_names))) return st;
    if (own_seq_off) {
        CK(cudaMemcpy(&seq_bytes, own_seq_off + own_count, 8, cudaMemcpyDefault));
        if ((st = adopt(ctx, ctx->b_seq_off, own_seq_off, (size_t)own_count + 1, &ctx->d_seq_off))) return st;
    } else {
        // lengths only: build local offsets on the host from the global length array
        if (!lengths) return RAFTGPU_E_ARG;
        std::vector<int64_t> len(own_count), off(own_count + 1, 0);
        if (own_count) CK(cudaMemcpy(len.data(), lengths + own_first, sizeof(int64_t) * own_count, cudaMemcpyDefault));
        for (int64_t i = 0; i < own_count; i++) off[i + 1] = off[i] + len[i];
        seq_bytes = off[own_count];
        CK(ctx->b_seq_off.ensure(sizeof(int64_t) * (own_count + 1) + 32));
        CK(cudaMemcpy(ctx->b_seq_off.p, off.data(), sizeof(int64_t) * (own_count + 1), cudaMemcpyHostToDevice));
        ctx->d_seq_off = ctx->b_seq_off.as<int64_t>();
    }
    ctx->have_seq = own_seq != nullptr || seq_bytes == 0;
    if (own_seq) { if ((st = adopt_seq(ctx, own_seq, (size_t)seq_bytes))) return st; }
    else ctx->d_seq = nullptr;
    std::string first;
    if ((st = first_name_of(ctx, n, name_off, names, first))) return st;
    st = build_layout_and_names(ctx, first);
    cudaEventRecord(ctx->ev[1], ctx->st);
    if (cudaEventSynchronize(ctx->ev[1]) == cudaSuccess) cudaEventElapsedTime(&ctx->stats.ms_set_reads, ctx->ev[0], ctx->ev[1]);
    return st;
}

// ------------------------------------------------------------------------------------------------ device FASTA ingest
static int fasta_chunk(raftgpu_ctx* ctx, const uint8_t* dtext, int64_t len, bool first, bool last)
{
    if (len <= 0) return RAFTGPU_OK;
    Misc*     M = ctx->misc();
    const int tiles = fasta_tokenize_tiles(len);
    CK(ctx->b_fa_status.ensure(sizeof(uint64_t) * 4 * (size_t)(tiles + 8)));
    CK(ctx->b_seq.ensure((size_t)(ctx->fa_bases + len) + 64, (size_t)ctx->fa_bases, ctx->st)); // arena: at most one byte per text byte
    int64_t want_cap = len / 512 + 1024;
    for (int attempt = 0; attempt < 2; attempt++) {
        const int64_t need = ctx->fa_n + want_cap + 1;
        if (need > ctx->fa_rec_cap) {
            int64_t cap = need + need / 4;
            CK(ctx->b_seq_off.ensure(sizeof(int64_t) * (size_t)(cap + 1), sizeof(int64_t) * (size_t)ctx->fa_n, ctx->st));
            CK(ctx->b_name_off.ensure(sizeof(int64_t) * (size_t)(cap + 1), sizeof(int64_t) * (size_t)ctx->fa_n, ctx->st));
            ctx->fa_rec_cap = cap;
        }
        if (ctx->fa_fastq && need + 2 > ctx->fa_gcap) { // file offsets of every record's '@' and quality line (index = line / 4)
            // entries written so far: a record's '@' offset once its line 4k was seen, its quality offset once line 4k+3 was
            const int64_t cap = need + need / 4 + 2;
            const size_t  keep_rec = sizeof(int64_t) * (size_t)std::min<int64_t>(ctx->fa_gcap, (ctx->fa_lines + 3) >> 2);
            const size_t  keep_qual = sizeof(int64_t) * (size_t)std::min<int64_t>(ctx->fa_gcap, ctx->fa_lines >> 2);
            CK(ctx->b_fa_rec_gpos.ensure(sizeof(int64_t) * (size_t)cap, keep_rec, ctx->st));
            CK(ctx->b_fa_qual_gpos.ensure(sizeof(int64_t) * (size_t)cap, keep_qual, ctx->st));
            ctx->fa_gcap = cap;
        }
        CK(ctx->b_fa_rec_pos.ensure(sizeof(int64_t) * (size_t)(want_cap + 1)));
        CK(cudaMemsetAsync(ctx->b_fa_status.p, 0, sizeof(uint64_t) * 4 * (size_t)tiles, ctx->st));
        CK(cudaMemsetAsync(&M->ticket, 0, sizeof(int), ctx->st));
        CK(cudaMemsetAsync(&M->fa_flags, 0, sizeof(int) * 2 + sizeof(long long) * 3, ctx->st));
        FastaTokArgs a{};
        a.text = dtext; a.nbytes = len; a.n_tiles = tiles; a.first_chunk = first; a.last_chunk = last;
        a.seq_out = ctx->b_seq.as<uint8_t>() + ctx->fa_bases; a.seq_off_base = ctx->fa_bases;
        a.rec_pos = ctx->b_fa_rec_pos.as<int64_t>(); a.seq_off = ctx->b_seq_off.as<int64_t>() + ctx->fa_n; a.rec_cap = want_cap;
        a.st_carry = ctx->b_fa_status.as<uint64_t>(); a.st_keep = a.st_carry + tiles; a.st_rec = a.st_keep + tiles;
        a.ticket = &M->ticket; a.flags = &M->fa_flags; a.totals = M->fa_totals;
        a.fastq = ctx->fa_fastq; a.line_base = ctx->fa_lines; a.text_gbase = ctx->fa_text_bytes; a.st_nl = a.st_rec + tiles;
        a.rec_gpos = ctx->b_fa_rec_gpos.as<int64_t>(); a.qual_gpos = ctx->b_fa_qual_gpos.as<int64_t>(); a.rec_gcap = ctx->fa_gcap;
        CK(launch_fasta_tokenize(a, ctx->st));
        ctx->launches++;
        struct { int flags, pad; long long tot[3]; } h;
        CK(cudaMemcpyAsync(&h, &M->fa_flags, sizeof h, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        if (h.flags) FAIL(RAFTGPU_E_UNSUPPORTED, "read text needs the host reader (FASTQ that is not strictly four lines per record, a '+' line in FASTA, CR LF line ends, or no leading '>' / '@')");
        if (h.tot[0] > want_cap) { want_cap = h.tot[0]; continue; } // more records than the optimistic capacity: redo this chunk
        const int64_t nrec = h.tot[0];
        // names of this chunk's records
        CK(ctx->b_fa_name_len.ensure(sizeof(int32_t) * (size_t)(nrec + 1)));
        CK(ctx->b_fa_name_off_chunk.ensure(sizeof(int64_t) * (size_t)(nrec + 2)));
        CK(ctx->b_status.ensure(sizeof(uint64_t) * (size_t)(scan_tiles_small(nrec) + 8)));
        launch_fasta_name_len(dtext, len, ctx->b_fa_rec_pos.as<int64_t>(), nrec, ctx->b_fa_name_len.as<int32_t>(), ctx->st);
        CKL();
        launch_scan_i32_to_i64(ctx->b_fa_name_len.as<int32_t>(), ctx->b_fa_name_off_chunk.as<int64_t>(), nrec, ctx->b_status.as<uint64_t>(), &M->ticket, ctx->st);
        CKL();
        long long nb = 0;
        CK(cudaMemcpyAsync(&nb, ctx->b_fa_name_off_chunk.as<int64_t>() + nrec, 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        CK(ctx->b_names.ensure((size_t)(ctx->fa_name_bytes + nb) + 64, (size_t)ctx->fa_name_bytes, ctx->st));
        launch_fasta_name_copy(dtext, ctx->b_fa_rec_pos.as<int64_t>(), ctx->b_fa_name_off_chunk.as<int64_t>(), nrec,
                               ctx->b_names.as<uint8_t>() + ctx->fa_name_bytes, ctx->st);
        CKL();
        // global name offsets of these records = chunk offsets + bytes so far
        launch_add_offset_i64(ctx->b_fa_name_off_chunk.as<int64_t>(), nrec, ctx->fa_name_bytes, ctx->b_name_off.as<int64_t>() + ctx->fa_n, ctx->st);
        CKL();
        CK(cudaStreamSynchronize(ctx->st));
        ctx->fa_n += nrec; ctx->fa_bases += h.tot[1]; ctx->fa_name_bytes += nb;
        ctx->fa_lines += h.tot[2]; ctx->fa_text_bytes += len;
        return RAFTGPU_OK;
    }
    FAIL(RAFTGPU_E_STATE, "FASTA record capacity retry failed");
}

extern "C" int raftgpu_ingest_fasta(raftgpu_ctx* ctx, const uint8_t* text, size_t nbytes, int last_chunk, uint64_t total_hint)
{
    if (!ctx || (!text && nbytes)) return RAFTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int st;
    const bool begin = !ctx->fasta_active;
    if (begin) {
        if ((st = raftgpu_reset(ctx))) return st;
        ctx->fasta_active = true; ctx->fa_rec_cap = 0; ctx->fa_gcap = 0;
        cudaEventRecord(ctx->ev[0], ctx->st);
        if (total_hint) CK(ctx->b_seq.ensure((size_t)total_hint + 64));
    }
    const bool   dev = nbytes && is_device_ptr(text);
    const size_t total = ctx->fa_carry.size() + nbytes;
    const uint8_t* dtext = nullptr;
    if (dev && ctx->fa_carry.empty() && ((uintptr_t)text & 15) == 0) {
        dtext = text;
    } else if (total) {
        CK(ctx->b_fa_text.ensure(total + 64));
        uint8_t* d = ctx->b_fa_text.as<uint8_t>();
        if (!ctx->fa_carry.empty()) CK(cudaMemcpyAsync(d, ctx->fa_carry.data(), ctx->fa_carry.size(), cudaMemcpyHostToDevice, ctx->st));
        if (nbytes) CK(cudaMemcpyAsync(d + ctx->fa_carry.size(), text, nbytes, cudaMemcpyDefault, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        dtext = d;
    }
    const bool first = ctx->fa_text_bytes == 0; // nothing tokenised yet
    if (!ctx->fa_mode_known && total) { // strict FASTQ?  '@' first and a '+' opening the third line (anything less regular fails the checks later)
        const size_t         peek = std::min<size_t>(total, 64u << 10);
        std::vector<uint8_t> head(peek);
        CK(cudaMemcpy(head.data(), dtext, peek, cudaMemcpyDeviceToHost));
        size_t p = 0;
        int    nl = 0;
        while (p < peek && nl < 2) { if (head[p] == '\n') nl++; p++; }
        const bool decided = head[0] != '@' || (nl == 2 && p < peek) || peek < total; // not FASTQ, or three lines seen, or lines too long to bother
        if (!decided && !last_chunk) { // too little text to tell: keep it for the next call
            std::vector<uint8_t> all(total);
            CK(cudaMemcpy(all.data(), dtext, total, cudaMemcpyDeviceToHost));
            ctx->fa_carry.swap(all);
            return RAFTGPU_OK;
        }
        ctx->fa_fastq = head[0] == '@' && nl == 2 && p < peek && head[p] == '+';
        ctx->fa_mode_known = true;
    }
    size_t proc = total;
    if (!last_chunk && total) { // hold back the unterminated last line
        size_t scan = std::min<size_t>(total, 1 << 20);
        std::vector<uint8_t> tail;
        long long nlpos = -1;
        for (;;) {
            tail.resize(scan);
            CK(cudaMemcpy(tail.data(), dtext + (total - scan), scan, cudaMemcpyDeviceToHost));
            for (long long k = (long long)scan - 1; k >= 0; k--) if (tail[k] == '\n') { nlpos = (long long)(total - scan) + k; break; }
            if (nlpos >= 0 || scan == total) break;
            scan = std::min<size_t>(total, scan * 8);
        }
        proc = (size_t)(nlpos + 1);
        ctx->fa_carry.assign(tail.begin() + (proc - (total - scan)), tail.end());
    } else ctx->fa_carry.clear();
    if (proc) { // does the text seen so far end with a newline?  (only the last chunk can end without one)
        uint8_t b = '\n';
        if (last_chunk) CK(cudaMemcpy(&b, dtext + proc - 1, 1, cudaMemcpyDeviceToHost));
        ctx->fa_ends_nl = b == '\n';
    }
    st = fasta_chunk(ctx, dtext, (int64_t)proc, first, last_chunk != 0);
    if (st) { raftgpu_reset(ctx); return st; }
    if (!last_chunk) return RAFTGPU_OK;
    // ---- all records are in: close the offset arrays and continue exactly like raftgpu_set_reads
    const int64_t n = ctx->fa_n;
    if (n > 0x7ffffff0ll) { raftgpu_reset(ctx); return RAFTGPU_E_ARG; }
    if (ctx->fa_rec_cap < n + 1) { // empty input: nothing was allocated yet
        CK(ctx->b_seq_off.ensure(sizeof(int64_t) * 2)); CK(ctx->b_name_off.ensure(sizeof(int64_t) * 2)); CK(ctx->b_seq.ensure(64)); CK(ctx->b_names.ensure(64));
    }
    CK(cudaMemcpy(ctx->b_seq_off.as<int64_t>() + n, &ctx->fa_bases, 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->b_name_off.as<int64_t>() + n, &ctx->fa_name_bytes, 8, cudaMemcpyHostToDevice));
    if (ctx->fa_fastq) { // four lines per record, and every quality line as long as its bases (kseq.h:290-296)
        const int64_t lines = ctx->fa_lines + (ctx->fa_ends_nl ? 0 : 1);
        int           bad = lines != 4 * n;
        if (!bad && n > 0) {
            Misc* M = ctx->misc();
            CK(cudaMemsetAsync(&M->fa_flags, 0, sizeof(int), ctx->st));
            launch_fastq_verify(n, ctx->b_seq_off.as<int64_t>(), ctx->b_fa_rec_gpos.as<int64_t>(), ctx->b_fa_qual_gpos.as<int64_t>(),
                                ctx->fa_text_bytes + (ctx->fa_ends_nl ? 0 : 1), &M->fa_flags, ctx->st);
            CKL();
            CK(cudaMemcpyAsync(&bad, &M->fa_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
            CK(cudaStreamSynchronize(ctx->st));
        }
        if (bad) {
            raftgpu_reset(ctx);
            FAIL(RAFTGPU_E_UNSUPPORTED, "FASTQ text is not strictly four lines per record with quality lines as long as the bases: use the host reader");
        }
    }
    CK(cudaMemsetAsync(ctx->b_seq.as<uint8_t>() + ctx->fa_bases, 0, 32, ctx->st));
    ctx->fasta_active = false;
    ctx->n = n; ctx->m = n; ctx->own_first = 0;
    ctx->d_seq_off = ctx->b_seq_off.as<int64_t>(); ctx->d_seq = ctx->b_seq.as<uint8_t>();
    ctx->d_name_off = ctx->b_name_off.as<int64_t>(); ctx->d_names = ctx->b_names.as<uint8_t>();
    ctx->have_seq = true;
    std::string first_name;
    if ((st = first_name_of(ctx, n, ctx->d_name_off, ctx->d_names, first_name))) return st;
    st = build_layout_and_names(ctx, first_name);
    cudaEventRecord(ctx->ev[1], ctx->st);
    if (cudaEventSynchronize(ctx->ev[1]) == cudaSuccess) cudaEventElapsedTime(&ctx->stats.ms_set_reads, ctx->ev[0], ctx->ev[1]);
    return st;
}

// ------------------------------------------------------------------------------------------------ PAF
static int ensure_records(raftgpu_ctx* ctx, int64_t cap)
{
    if (cap <= ctx->rec_cap) return RAFTGPU_OK;
    cap += cap / 8 + 1024;
    size_t keep4 = sizeof(int32_t) * (size_t)ctx->n_rec, keep1 = (size_t)ctx->n_rec;
    CK(ctx->b_qid.ensure(sizeof(int32_t) * cap, keep4, ctx->st)); CK(ctx->b_tid.ensure(sizeof(int32_t) * cap, keep4, ctx->st));
    CK(ctx->b_qs.ensure(sizeof(int32_t) * cap, keep4, ctx->st));  CK(ctx->b_qe.ensure(sizeof(int32_t) * cap, keep4, ctx->st));
    CK(ctx->b_ts.ensure(sizeof(int32_t) * cap, keep4, ctx->st));  CK(ctx->b_te.ensure(sizeof(int32_t) * cap, keep4, ctx->st));
    CK(ctx->b_strand.ensure((size_t)cap, keep1, ctx->st));
    ctx->rec_cap = cap;
    return RAFTGPU_OK;
}

static int zero_diff(raftgpu_ctx* ctx);

// tokenise `len` bytes of device text (complete lines, except that the last line may lack its newline)
static int tokenize_device(raftgpu_ctx* ctx, const uint8_t* dtext, int64_t len)
{
    if (len <= 0) return RAFTGPU_OK;
    Misc* M = ctx->misc();
    if (!ctx->rec0_external) { // first record of the file: peek until one is found (no-op kernel once present)
        int present = 0;
        CK(cudaMemcpyAsync(&present, &M->rec0[6], sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        if (!present) { launch_paf_peek(dtext, len, ctx->nt, M->rec0, &M->err, ctx->st); CKL(); }
    }
    const int tiles = paf_tokenize_tiles(len);
    CK(ctx->b_status.ensure(sizeof(uint64_t) * (size_t)(tiles + 8)));
    int st = ensure_records(ctx, ctx->n_rec + len / 48 + 1024);
    if (st) return st;
    if ((st = zero_diff(ctx))) return st; // the tokenizer adds the query sides as it decodes them
    ctx->q_scattered = true;
    for (int attempt = 0; attempt < 2; attempt++) {
        CK(cudaMemsetAsync(ctx->b_status.p, 0, sizeof(uint64_t) * (size_t)tiles, ctx->st));
        CK(cudaMemsetAsync(&M->ticket, 0, sizeof(int), ctx->st));
        PafTokArgs a{};
        a.text = dtext; a.nbytes = len; a.rec_base = ctx->n_rec; a.rec_cap = ctx->rec_cap;
        a.qid = ctx->b_qid.as<int32_t>(); a.tid = ctx->b_tid.as<int32_t>(); a.qs = ctx->b_qs.as<int32_t>(); a.qe = ctx->b_qe.as<int32_t>();
        a.ts = ctx->b_ts.as<int32_t>(); a.te = ctx->b_te.as<int32_t>(); a.strand = ctx->b_strand.as<uint8_t>();
        a.rec0 = M->rec0; a.first_is_local = ctx->first_is_local; a.n_tiles = tiles;
        a.status = ctx->b_status.as<uint64_t>(); a.ticket = &M->ticket; a.sym_flag = &M->sym_flag; a.err = &M->err;
        a.n_records_out = (int64_t*)&M->n_records_out; a.names = ctx->nt;
        // a retry (record capacity overflow) decodes the same lines again: their query sides are already in
        a.diff = attempt == 0 ? ctx->b_diff.as<int32_t>() : nullptr; a.slot_off = ctx->b_slot_off.as<int64_t>(); a.reso = ctx->prm.reso;
        a.own_first = ctx->own_first; a.own_count = ctx->m; a.err_range = &M->err_range;
        CK(launch_paf_tokenize(a, ctx->st));
        ctx->launches++;
        long long n_out = 0;
        CK(cudaMemcpyAsync(&n_out, &M->n_records_out, sizeof n_out, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(&ctx->h_sym, &M->sym_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
        if ((st = fetch_err(ctx))) return st;
        if (n_out > 0x7fffffffll) FAIL(RAFTGPU_E_ARG, "more than 2^31-1 PAF records (the reference counts them in an int, chop.hpp:139)");
        if (n_out <= ctx->rec_cap) { ctx->n_rec = n_out; return RAFTGPU_OK; }
        if ((st = ensure_records(ctx, n_out))) return st; // optimistic capacity was too small: grow and redo this chunk
    }
    FAIL(RAFTGPU_E_STATE, "record capacity retry failed");
}

extern "C" int raftgpu_ingest_paf(raftgpu_ctx* ctx, const uint8_t* text, size_t nbytes, int last_chunk)
{
    if (!ctx || (!text && nbytes)) return RAFTGPU_E_ARG;
    if (!ctx->have_reads || ctx->paf_done) FAIL(RAFTGPU_E_STATE, "raftgpu_ingest_paf: set reads first / PAF already complete");
    CK(cudaSetDevice(ctx->device));
    cudaEventRecord(ctx->ev[0], ctx->st);
    const bool dev = nbytes && is_device_ptr(text);
    const size_t total = ctx->carry.size() + nbytes;
    ctx->paf_bytes += (int64_t)nbytes;
    const uint8_t* dtext = nullptr;
    if (dev && ctx->carry.empty() && ((uintptr_t)text & 15) == 0) {
        dtext = text; // zero copy
    } else if (total) {
        CK(ctx->b_text.ensure(total + 64));
        uint8_t* d = ctx->b_text.as<uint8_t>();
        if (!ctx->carry.empty()) CK(cudaMemcpyAsync(d, ctx->carry.data(), ctx->carry.size(), cudaMemcpyHostToDevice, ctx->st));
        if (nbytes) CK(cudaMemcpyAsync(d + ctx->carry.size(), text, nbytes, cudaMemcpyDefault, ctx->st));
        CK(cudaStreamSynchronize(ctx->st)); // carry is reused below
        dtext = d;
    }
    size_t proc = total;
    if (!last_chunk && total) {
        // keep the unterminated tail for the next chunk (chunks need not end on newlines)
        size_t scan = std::min<size_t>(total, 1 << 20);
        std::vector<uint8_t> tail;
        long long nlpos = -1;
        for (;;) {
            tail.resize(scan);
            CK(cudaMemcpy(tail.data(), dtext + (total - scan), scan, cudaMemcpyDeviceToHost));
            for (long long k = (long long)scan - 1; k >= 0; k--) if (tail[k] == '\n') { nlpos = (long long)(total - scan) + k; break; }
            if (nlpos >= 0 || scan == total) break;
            scan = std::min<size_t>(total, scan * 8);
        }
        proc = (size_t)(nlpos + 1);
        size_t off_in_tail = proc - (total - scan);
        ctx->carry.assign(tail.begin() + off_in_tail, tail.end());
    } else {
        ctx->carry.clear();
    }
    int st = tokenize_device(ctx, dtext, (int64_t)proc);
    if (st) {
        // a sharded run must still reach its collectives: remember the data error for raftgpu_sharded_finish
        if (ctx->shard_begun && (st == RAFTGPU_E_UNKNOWN_NAME || st == RAFTGPU_E_RANGE)) { ctx->pending_ingest_status = st; ctx->paf_done = true; }
        return st;
    }
    if (last_chunk) {
        ctx->paf_done = true;
        // PAF is on the device: the arena upload can overlap everything that follows.  A sharded run starts it after its
        // exchange instead: the copy engine serves host-to-device copies in order, and the small uploads of the exchange
        // (bounds, bucket cursors) would otherwise queue behind tens of GB of arena chunks.
        if (!ctx->shard_begun && (st = start_seq_upload(ctx))) return st;
    }
    cudaEventRecord(ctx->ev[1], ctx->st);
    CK(cudaStreamSynchronize(ctx->st));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->stats.ms_tokenize += ms;
    return RAFTGPU_OK;
}

extern "C" int raftgpu_peek_first_record(raftgpu_ctx* ctx, const uint8_t* text, size_t nbytes, int32_t rec[6], int32_t* found)
{
    if (!ctx || !rec || !found) return RAFTGPU_E_ARG;
    if (!ctx->have_reads) FAIL(RAFTGPU_E_STATE, "set reads first");
    CK(cudaSetDevice(ctx->device));
    *found = 0;
    if (!nbytes) return RAFTGPU_OK;
    size_t head = std::min<size_t>(nbytes, 4u << 20); // the first record sits at the very start of any sane PAF
    for (;;) {
        CK(ctx->b_text.ensure(head + 64));
        CK(cudaMemcpyAsync(ctx->b_text.p, text, head, cudaMemcpyDefault, ctx->st));
        launch_paf_peek(ctx->b_text.as<uint8_t>(), (int64_t)head, ctx->nt, ctx->misc()->rec0, &ctx->misc()->err, ctx->st);
        CKL();
        int r[8];
        CK(cudaMemcpyAsync(r, ctx->misc()->rec0, sizeof r, cudaMemcpyDeviceToHost, ctx->st));
        int st = fetch_err(ctx);
        if (st) return st;
        if (r[6] || head == nbytes) { for (int k = 0; k < 6; k++) rec[k] = r[k]; *found = r[6]; break; }
        head = std::min<size_t>(nbytes, head * 8);
    }
    // leave rec0 unset on the device: the caller decides through raftgpu_set_first_record
    int zero = 0;
    CK(cudaMemcpy(&ctx->misc()->rec0[6], &zero, sizeof zero, cudaMemcpyHostToDevice));
    return RAFTGPU_OK;
}

extern "C" int raftgpu_set_first_record(raftgpu_ctx* ctx, const int32_t rec[6], int32_t is_local)
{
    if (!ctx) return RAFTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int r[8] = {0};
    if (rec) { for (int k = 0; k < 6; k++) r[k] = rec[k]; r[6] = 1; }
    CK(cudaMemcpy(ctx->misc()->rec0, r, sizeof r, cudaMemcpyHostToDevice));
    ctx->rec0_external = true; ctx->first_is_local = is_local != 0 ? 1 : 0;
    return RAFTGPU_OK;
}
extern "C" int raftgpu_get_symmetric(raftgpu_ctx* ctx, int32_t* flag)
{
    if (!ctx || !flag) return RAFTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaMemcpy(flag, &ctx->misc()->sym_flag, sizeof(int), cudaMemcpyDeviceToHost));
    return RAFTGPU_OK;
}
extern "C" int raftgpu_set_symmetric(raftgpu_ctx* ctx, int32_t flag)
{
    if (!ctx) return RAFTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int f = flag != 0;
    CK(cudaMemcpy(&ctx->misc()->sym_flag, &f, sizeof(int), cudaMemcpyHostToDevice));
    ctx->sym_external = true; ctx->h_sym = f;
    return RAFTGPU_OK;
}

// ------------------------------------------------------------------------------------------------ coverage
static ScatterArgs scatter_args(raftgpu_ctx* ctx)
{
    ScatterArgs a{};
    a.qid = ctx->b_qid.as<int32_t>(); a.tid = ctx->b_tid.as<int32_t>(); a.qs = ctx->b_qs.as<int32_t>(); a.qe = ctx->b_qe.as<int32_t>();
    a.ts = ctx->b_ts.as<int32_t>(); a.te = ctx->b_te.as<int32_t>(); a.n_rec = ctx->n_rec;
    a.slot_off = ctx->b_slot_off.as<int64_t>(); a.diff = ctx->b_diff.as<int32_t>(); a.reso = ctx->prm.reso;
    a.own_first = ctx->own_first; a.own_count = ctx->m; a.sym_flag = &ctx->misc()->sym_flag; a.err = &ctx->misc()->err;
    return a;
}

static int zero_diff(raftgpu_ctx* ctx)
{
    if (ctx->diff_zeroed) return RAFTGPU_OK;
    CK(cudaStreamSynchronize(ctx->st_zero)); // a re-zeroing of the previous run may still be writing the old buffer
    CK(ctx->b_diff.ensure(sizeof(int32_t) * (size_t)(ctx->n_slots + 8)));
    if (ctx->diff_zero_ptr == ctx->b_diff.p && ctx->diff_zero_ints >= ctx->n_slots) {
        // the previous run zeroed the array again right after its scan had consumed it (finalize), behind K3 and the emitters
    } else {
        CK(cudaMemsetAsync(ctx->b_diff.p, 0, sizeof(int32_t) * (size_t)ctx->n_slots, ctx->st));
    }
    ctx->diff_zero_ints = 0;
    ctx->diff_zeroed = true;
    return RAFTGPU_OK;
}

static int accumulate_local(raftgpu_ctx* ctx)
{
    int st = zero_diff(ctx);
    if (st) return st;
    ScatterArgs a = scatter_args(ctx);
    a.skip_query = ctx->q_scattered ? 1 : 0;
    // symmetric overlaps + query sides already in: nothing left to add (target sides do not contribute, repeat.hpp:54)
    const bool sym_known = ctx->sym_external || ctx->paf_done;
    if (a.skip_query && sym_known && ctx->h_sym) return RAFTGPU_OK;
    launch_scatter_records(a, ctx->st);
    CKL();
    return RAFTGPU_OK;
}

extern "C" int raftgpu_accumulate_local(raftgpu_ctx* ctx)
{
    if (!ctx) return RAFTGPU_E_ARG;
    if (!ctx->have_reads || !ctx->paf_done || ctx->finalized) FAIL(RAFTGPU_E_STATE, "raftgpu_accumulate_local: wrong state");
    CK(cudaSetDevice(ctx->device));
    cudaEventRecord(ctx->ev[2], ctx->st);
    int st = accumulate_local(ctx);
    if (st) return st;
    cudaEventRecord(ctx->ev[3], ctx->st);
    CK(cudaStreamSynchronize(ctx->st));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]); ctx->stats.ms_scatter += ms;
    return RAFTGPU_OK;
}

extern "C" int raftgpu_accumulate_endpoints(raftgpu_ctx* ctx, const void* ep, int64_t count)
{
    if (!ctx || count < 0 || (count && !ep)) return RAFTGPU_E_ARG;
    if (!ctx->have_reads || ctx->finalized) FAIL(RAFTGPU_E_STATE, "raftgpu_accumulate_endpoints: wrong state");
    CK(cudaSetDevice(ctx->device));
    int st = zero_diff(ctx);
    if (st) return st;
    launch_scatter_endpoints((const int32_t*)ep, count, ctx->b_slot_off.as<int64_t>(), ctx->b_diff.as<int32_t>(), ctx->prm.reso, ctx->own_first,
                             ctx->m, &ctx->misc()->err, ctx->st);
    CKL();
    CK(cudaStreamSynchronize(ctx->st));
    return RAFTGPU_OK;
}

constexpr unsigned long long ROUTE_LIST_CAP = 4ull << 20; // collected endpoints (16 B each); more: two-pass packing

extern "C" int raftgpu_route_count(raftgpu_ctx* ctx, int nranks, const int64_t* bounds, int64_t* counts)
{
    if (!ctx || nranks < 1 || nranks > 64 || !bounds || !counts) return RAFTGPU_E_ARG;
    if (!ctx->paf_done) FAIL(RAFTGPU_E_STATE, "raftgpu_route_count: finish PAF ingest first");
    CK(cudaSetDevice(ctx->device));
    Misc* M = ctx->misc();
    CK(ctx->b_bounds.ensure(sizeof(int64_t) * 65));
    CK(ctx->b_route_list.ensure(sizeof(int4) * ROUTE_LIST_CAP));
    CK(cudaMemcpyAsync(ctx->b_bounds.p, bounds, sizeof(int64_t) * (nranks + 1), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemsetAsync(M->route_counts, 0, sizeof(unsigned long long) * 65, ctx->st)); // counts and the list cursor behind them
    unsigned long long cap = ROUTE_LIST_CAP;
    if (const char* e = getenv("RAFT_B200_ROUTE_CAP")) { long long v = atoll(e); if (v >= 0 && (unsigned long long)v < cap) cap = (unsigned long long)v; } // test knob: force the two-pass packing
    launch_route_collect(scatter_args(ctx), nranks, ctx->b_bounds.as<int64_t>(), M->route_counts, ctx->b_route_list.as<int4>(), &M->route_list_n,
                         cap, ctx->st);
    CKL();
    unsigned long long h[65];
    CK(cudaMemcpyAsync(h, M->route_counts, sizeof h, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    for (int r = 0; r < nranks; r++) counts[r] = (int64_t)h[r];
    ctx->route_collected = h[64] <= cap ? (int64_t)h[64] : -1;
    return RAFTGPU_OK;
}

extern "C" int raftgpu_route_pack(raftgpu_ctx* ctx, int nranks, const int64_t* bounds, const int64_t* counts, void* sendbuf)
{
    if (!ctx || nranks < 1 || nranks > 64 || !bounds || !counts) return RAFTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    Misc* M = ctx->misc();
    unsigned long long cur[64] = {0};
    for (int r = 1; r < nranks; r++) cur[r] = cur[r - 1] + (unsigned long long)counts[r - 1];
    CK(cudaMemcpyAsync(M->route_counts, cur, sizeof cur, cudaMemcpyHostToDevice, ctx->st));
    if (ctx->route_collected >= 0) { // the count pass kept every endpoint: bucket the short list
        launch_route_pack_list(ctx->b_route_list.as<int4>(), ctx->route_collected, M->route_counts, (int32_t*)sendbuf, ctx->st);
    } else {
        CK(ctx->b_bounds.ensure(sizeof(int64_t) * 65));
        CK(cudaMemcpyAsync(ctx->b_bounds.p, bounds, sizeof(int64_t) * (nranks + 1), cudaMemcpyHostToDevice, ctx->st));
        launch_route_pack(scatter_args(ctx), nranks, ctx->b_bounds.as<int64_t>(), M->route_counts, (int32_t*)sendbuf, ctx->st);
    }
    CKL();
    ctx->route_collected = -1;
    CK(cudaStreamSynchronize(ctx->st));
    return RAFTGPU_OK;
}

// ------------------------------------------------------------------------------------------------ finalize
static int layout_outputs(raftgpu_ctx* ctx);

extern "C" int raftgpu_finalize(raftgpu_ctx* ctx, raftgpu_stats* out)
{
    if (!ctx) return RAFTGPU_E_ARG;
    if (!ctx->have_reads) FAIL(RAFTGPU_E_STATE, "raftgpu_finalize: no reads");
    if (ctx->finalized) FAIL(RAFTGPU_E_STATE, "raftgpu_finalize: already finalized");
    CK(cudaSetDevice(ctx->device));
    const raftgpu_params& P = ctx->prm;
    Misc* M = ctx->misc();
    int st = zero_diff(ctx);
    if (st) return st;
    const int64_t m = ctx->m;
    // K2b: coverage = inclusive scan of the difference array
    cudaEventRecord(ctx->ev[3], ctx->st);
    CK(ctx->b_status.ensure(sizeof(uint64_t) * (size_t)(scan_tiles_cov(ctx->n_slots) + scan_tiles_small(std::max<int64_t>(m, ctx->cut_cap_total + m)) + 16)));
    CovSizeArgs cs{ctx->b_cov_tile_bytes.as<int32_t>(), ctx->b_cov_tile_static.as<int32_t>(), ctx->b_slot_off.as<int64_t>(), ctx->b_cov_tile_read.as<int32_t>()};
    CK(ctx->b_cov.ensure(sizeof(int32_t) * (size_t)(ctx->n_slots + 8)));
    launch_scan_cov(ctx->b_diff.as<int32_t>(), ctx->b_cov.as<int32_t>(), ctx->n_slots, ctx->b_status.as<uint64_t>(), &M->ticket, cs, ctx->st);
    CKL();
    cudaEventRecord(ctx->ev[4], ctx->st);
    // K3: repeats + cut points
    const int H = (int)(P.est_cov * P.cov_mul); // repeat.hpp:89-90: int * double, truncated
    CK(ctx->b_rep.ensure(sizeof(int2) * (size_t)(ctx->rep_cap_total + 1))); CK(ctx->b_rep_cnt.ensure(sizeof(int32_t) * (size_t)(m + 1)));
    CK(ctx->b_cuts.ensure(sizeof(int32_t) * (size_t)(ctx->cut_cap_total + 1))); CK(ctx->b_frag_cnt.ensure(sizeof(int32_t) * (size_t)(m + 1)));
    CK(cudaMemsetAsync(&M->work_counter, 0, sizeof(int), ctx->st));
    CK(cudaMemsetAsync(M->stats, 0, sizeof M->stats, ctx->st));
    RepeatCutArgs ra{};
    ra.cov = ctx->b_cov.as<int32_t>(); ra.slot_off = ctx->b_slot_off.as<int64_t>(); ra.seq_off = ctx->d_seq_off; ra.m = m;
    ra.reso = P.reso; ra.H = H; ra.p = P.repeat_length; ra.P = P.interval_length; ra.f = P.flanking_length; ra.l = P.read_length;
    ra.rep_cap_off = ctx->b_rep_cap_off.as<int64_t>(); ra.cut_cap_off = ctx->b_cut_cap_off.as<int64_t>();
    ra.rep = ctx->b_rep.as<int2>(); ra.rep_cnt = ctx->b_rep_cnt.as<int32_t>(); ra.cuts = ctx->b_cuts.as<int32_t>();
    ra.frag_cnt = ctx->b_frag_cnt.as<int32_t>(); ra.stats = M->stats; ra.work_counter = &M->work_counter;
    launch_repeat_cut(ra, ctx->st);
    CKL();
    cudaEventRecord(ctx->ev[5], ctx->st);
    // the differences were consumed by the scan: zero them for the next run on a side stream, once K3 (latency-bound,
    // it would feel the competition) is done -- the memset then runs beside the layout kernels and the issue-bound text
    // emitter (measured: 76.5 -> 75.7 ms per pass; a thin persistent zeroing kernel instead of the memset gained less)
    CK(cudaEventRecord(ctx->ev_scan_done, ctx->st));
    CK(cudaStreamWaitEvent(ctx->st_zero, ctx->ev_scan_done, 0));
    CK(cudaMemsetAsync(ctx->b_diff.p, 0, sizeof(int32_t) * (size_t)ctx->n_slots, ctx->st_zero));
    ctx->diff_zero_ptr = ctx->b_diff.p; ctx->diff_zero_ints = ctx->n_slots;
    // fragment numbering inside this context
    CK(ctx->b_frag_base.ensure(sizeof(int64_t) * (size_t)(m + 1)));
    launch_scan_i32_to_i64(ctx->b_frag_cnt.as<int32_t>(), ctx->b_frag_base.as<int64_t>(), m, ctx->b_status.as<uint64_t>(), &M->ticket, ctx->st);
    CKL();
    CK(ctx->b_rep_off.ensure(sizeof(int64_t) * (size_t)(m + 1)));
    launch_scan_i32_to_i64(ctx->b_rep_cnt.as<int32_t>(), ctx->b_rep_off.as<int64_t>(), m, ctx->b_status.as<uint64_t>(), &M->ticket, ctx->st);
    CKL();
    long long G = 0, R = 0, nrec = ctx->n_rec;
    int       S = 0;
    unsigned long long stats2[2];
    if (ctx->gather_on) { // sharded run: every rank learns every rank's fragment count, statistics and error words in this same synchronisation
        const long long* src[3] = {(const long long*)(ctx->b_frag_base.as<int64_t>() + m), (const long long*)&M->stats[0], (const long long*)&M->stats[1]};
        long long*       snd = ctx->b_coll.as<long long>() + COLL_SEND;
        launch_pack_scalars(src, 3, &M->err, &M->err_range, snd, ctx->st);
        CKL();
        NCK(ncclAllGather(snd, ctx->b_coll.as<long long>() + COLL_FIN, 8, ncclInt64, ctx->comm, ctx->st));
        CK(cudaMemcpyAsync(ctx->h_coll + COLL_FIN, ctx->b_coll.as<long long>() + COLL_FIN, sizeof(long long) * 8 * (size_t)ctx->nranks, cudaMemcpyDeviceToHost, ctx->st));
    }
    CK(cudaMemcpyAsync(&G, ctx->b_frag_base.as<int64_t>() + m, 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(&R, ctx->b_rep_off.as<int64_t>() + m, 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(&S, &M->sym_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(stats2, M->stats, sizeof stats2, cudaMemcpyDeviceToHost, ctx->st));
    st = fetch_err(ctx);
    if (ctx->gather_on) {
        const long long* g = ctx->h_coll + COLL_FIN;
        raftgpu_shard_info& sh = ctx->shard;
        sh.n_fragments_total = sh.total_cov = sh.total_repeat_len = 0;
        long long before = 0;
        for (int r = 0; r < ctx->nranks; r++) {
            if (!st && ((unsigned long long)g[8 * r + 3] != ERR_CLEAN || (unsigned long long)g[8 * r + 4] != ERR_CLEAN)) st = RAFTGPU_E_PEER;
            if (r < ctx->rank) before += g[8 * r];
            sh.n_fragments_total += g[8 * r]; sh.total_cov += g[8 * r + 1]; sh.total_repeat_len += g[8 * r + 2];
        }
        if (st == RAFTGPU_E_PEER) ctx->last_error = "another rank reported a data error";
        ctx->read_num_base = before;
        sh.first_read_num = before + 1;
    }
    if (st) return st;
    ctx->G = G; ctx->n_repeats = R; ctx->finalized = true;
    raftgpu_stats& s = ctx->stats;
    s.n_reads = ctx->n; s.n_records = nrec; s.symmetric = S; s.high_cov = H; s.real_reads = ctx->real_reads;
    s.n_bins = ctx->n_slots - m;
    s.total_windows = (int32_t)(uint32_t)(uint64_t)s.n_bins; // int counter in the reference (repeat.hpp:95,117)
    s.total_cov = (int64_t)stats2[0]; s.total_repeat_len = (int64_t)stats2[1]; s.total_read_len = ctx->total_read_len;
    s.n_repeats = R; s.n_fragments = G;
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]); s.ms_scan = ms;
    cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]); s.ms_repeat_cut = ms;
    s.kernel_launches = ctx->launches;
    if (out) *out = s;
    return RAFTGPU_OK;
}

extern "C" int raftgpu_set_output_base(raftgpu_ctx* ctx, int64_t first_read_num)
{
    if (!ctx || first_read_num < 1) return RAFTGPU_E_ARG;
    if (!ctx->finalized) FAIL(RAFTGPU_E_STATE, "raftgpu_set_output_base: finalize first");
    ctx->read_num_base = first_read_num - 1;
    ctx->sized = false;
    return RAFTGPU_OK;
}

// sizes and offsets of the three output streams
static int layout_outputs(raftgpu_ctx* ctx)
{
    if (ctx->sized) return RAFTGPU_OK;
    if (!ctx->finalized) FAIL(RAFTGPU_E_STATE, "outputs requested before raftgpu_run / raftgpu_finalize");
    CK(cudaSetDevice(ctx->device));
    Misc* M = ctx->misc();
    const int64_t m = ctx->m, G = ctx->G;
    cudaEventRecord(ctx->ev[6], ctx->st);
    // fragments
    CK(ctx->b_frag_read.ensure(sizeof(int32_t) * (size_t)(G + 1))); CK(ctx->b_frag_a.ensure(sizeof(int32_t) * (size_t)(G + 1)));
    CK(ctx->b_frag_b.ensure(sizeof(int32_t) * (size_t)(G + 1)));    CK(ctx->b_frag_size.ensure(sizeof(int32_t) * (size_t)(G + 1)));
    CK(ctx->b_frag_off.ensure(sizeof(int64_t) * (size_t)(G + 1)));
    CK(ctx->b_status.ensure(sizeof(uint64_t) * (size_t)(scan_tiles_small(std::max<int64_t>(G, m)) + cov_tiles(ctx->n_slots) / 256 + 64)));
    FragExpandArgs fa{};
    fa.m = m; fa.seq_off = ctx->d_seq_off; fa.name_off = ctx->d_name_off; fa.own_first = ctx->own_first;
    fa.cut_cap_off = ctx->b_cut_cap_off.as<int64_t>(); fa.cuts = ctx->b_cuts.as<int32_t>(); fa.frag_cnt = ctx->b_frag_cnt.as<int32_t>();
    fa.frag_base = ctx->b_frag_base.as<int64_t>(); fa.v = ctx->prm.overlap_length; fa.read_num_base = ctx->read_num_base;
    fa.frag_read = ctx->b_frag_read.as<int32_t>(); fa.frag_a = ctx->b_frag_a.as<int32_t>(); fa.frag_b = ctx->b_frag_b.as<int32_t>();
    fa.frag_size = ctx->b_frag_size.as<int32_t>(); fa.err = &M->err; fa.sim = ctx->real_reads ? nullptr : ctx->b_sim.as<SimInfo>();
    launch_frag_expand(fa, ctx->st);
    CKL();
    launch_scan_i32_to_i64(ctx->b_frag_size.as<int32_t>(), ctx->b_frag_off.as<int64_t>(), G, ctx->b_status.as<uint64_t>(), &M->ticket, ctx->st);
    CKL();
    // long_repeats.txt lines
    CK(ctx->b_rep_line_size.ensure(sizeof(int32_t) * (size_t)(m + 1))); CK(ctx->b_rep_line_off.ensure(sizeof(int64_t) * (size_t)(m + 1)));
    launch_rep_sizes(ctx->b_rep_cnt.as<int32_t>(), ctx->b_rep_cap_off.as<int64_t>(), ctx->b_rep.as<int2>(), m, ctx->own_first,
                     ctx->b_rep_line_size.as<int32_t>(), ctx->st);
    CKL();
    launch_scan_i32_to_i64(ctx->b_rep_line_size.as<int32_t>(), ctx->b_rep_line_off.as<int64_t>(), m, ctx->b_status.as<uint64_t>(), &M->ticket, ctx->st);
    CKL();
    // long_repeats.bed lines (simulated reads only; the file is created empty otherwise, repeat.hpp:87,187)
    if (!ctx->real_reads) {
        CK(ctx->b_bed_line_size.ensure(sizeof(int32_t) * (size_t)(m + 1))); CK(ctx->b_bed_line_off.ensure(sizeof(int64_t) * (size_t)(m + 1)));
        launch_bed_sizes(ctx->b_rep_cnt.as<int32_t>(), ctx->b_rep_cap_off.as<int64_t>(), ctx->b_rep.as<int2>(), ctx->b_sim.as<SimInfo>(),
                         ctx->d_name_off, ctx->own_first, m, ctx->b_bed_line_size.as<int32_t>(), ctx->st);
        CKL();
        launch_scan_i32_to_i64(ctx->b_bed_line_size.as<int32_t>(), ctx->b_bed_line_off.as<int64_t>(), m, ctx->b_status.as<uint64_t>(), &M->ticket, ctx->st);
        CKL();
    }
    // coverage.txt tiles
    const int T = cov_tiles(ctx->n_slots);
    CK(ctx->b_cov_tile_off.ensure(sizeof(int64_t) * (size_t)(T + 1)));
    // tile byte counts were produced by the coverage scan (fused sizing)
    launch_scan_i32_to_i64(ctx->b_cov_tile_bytes.as<int32_t>(), ctx->b_cov_tile_off.as<int64_t>(), T, ctx->b_status.as<uint64_t>(), &M->ticket, ctx->st);
    CKL();
    // the host only needs a coarse view of the two offset tables (window -> tile / read range): every 256th entry
    const size_t nc = (size_t)(T / OFF_SAMPLE + 2), nr = (size_t)(m / OFF_SAMPLE + 2);
    ctx->h_cov_tile_off.resize(nc); ctx->h_rep_line_off.resize(nr);
    CK(ctx->b_off_sample.ensure(sizeof(int64_t) * (nc + nr)));
    launch_sample_i64(ctx->b_cov_tile_off.as<int64_t>(), T + 1, OFF_SAMPLE, ctx->b_off_sample.as<int64_t>(), ctx->st);
    CKL();
    launch_sample_i64(ctx->b_rep_line_off.as<int64_t>(), m + 1, OFF_SAMPLE, ctx->b_off_sample.as<int64_t>() + nc, ctx->st);
    CKL();
    long long fasta_bytes = 0, cov_bytes = 0, rep_bytes = 0, bed_bytes = 0;
    if (!ctx->real_reads) {
        ctx->h_bed_line_off.resize(nr);
        DevBuf& bs = ctx->b_bed_line_size; // reuse as the sample target (sizes are no longer needed): nr int64 fit? ensure
        CK(bs.ensure(sizeof(int64_t) * nr));
        launch_sample_i64(ctx->b_bed_line_off.as<int64_t>(), m + 1, OFF_SAMPLE, bs.as<int64_t>(), ctx->st);
        CKL();
        CK(cudaMemcpyAsync(ctx->h_bed_line_off.data(), bs.p, sizeof(int64_t) * nr, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(&bed_bytes, ctx->b_bed_line_off.as<int64_t>() + m, 8, cudaMemcpyDeviceToHost, ctx->st));
    }
    if (ctx->gather_on) { // file offsets of this rank's slices: all-gather of the four sizes (+ error words), read in the synchronisation below
        const long long* src[4] = {(const long long*)(ctx->b_cov_tile_off.as<int64_t>() + T), (const long long*)(ctx->b_rep_line_off.as<int64_t>() + m),
                                   ctx->real_reads ? (const long long*)&M->zero : (const long long*)(ctx->b_bed_line_off.as<int64_t>() + m),
                                   (const long long*)(ctx->b_frag_off.as<int64_t>() + G)};
        long long*       snd = ctx->b_coll.as<long long>() + COLL_SEND;
        launch_pack_scalars(src, 4, &M->err, &M->err_range, snd, ctx->st);
        CKL();
        NCK(ncclAllGather(snd, ctx->b_coll.as<long long>() + COLL_OUT, 8, ncclInt64, ctx->comm, ctx->st));
        CK(cudaMemcpyAsync(ctx->h_coll + COLL_OUT, ctx->b_coll.as<long long>() + COLL_OUT, sizeof(long long) * 8 * (size_t)ctx->nranks, cudaMemcpyDeviceToHost, ctx->st));
    }
    CK(cudaMemcpyAsync(ctx->h_cov_tile_off.data(), ctx->b_off_sample.p, sizeof(int64_t) * nc, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(ctx->h_rep_line_off.data(), ctx->b_off_sample.as<int64_t>() + nc, sizeof(int64_t) * nr, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(&cov_bytes, ctx->b_cov_tile_off.as<int64_t>() + T, 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(&rep_bytes, ctx->b_rep_line_off.as<int64_t>() + m, 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(&fasta_bytes, ctx->b_frag_off.as<int64_t>() + G, 8, cudaMemcpyDeviceToHost, ctx->st));
    cudaEventRecord(ctx->ev[7], ctx->st);
    int st = fetch_err(ctx);
    if (ctx->gather_on) {
        const long long* g = ctx->h_coll + COLL_OUT;
        raftgpu_shard_info& sh = ctx->shard;
        for (int w = 0; w < 4; w++) { sh.stream_base[w] = 0; sh.stream_total[w] = 0; }
        for (int r = 0; r < ctx->nranks; r++) {
            if (!st && ((unsigned long long)g[8 * r + 4] != ERR_CLEAN || (unsigned long long)g[8 * r + 5] != ERR_CLEAN)) st = RAFTGPU_E_PEER;
            for (int w = 0; w < 4; w++) { if (r < ctx->rank) sh.stream_base[w] += (uint64_t)g[8 * r + w]; sh.stream_total[w] += (uint64_t)g[8 * r + w]; }
        }
        if (st == RAFTGPU_E_PEER) ctx->last_error = "another rank reported a data error";
    }
    if (st) return st;
    raftgpu_stats& s = ctx->stats;
    s.out_bytes[RAFTGPU_OUT_COVERAGE] = (uint64_t)cov_bytes;
    s.out_bytes[RAFTGPU_OUT_LONG_REPEATS] = (uint64_t)rep_bytes;
    s.out_bytes[RAFTGPU_OUT_BED] = (uint64_t)bed_bytes; // real reads: the file is created empty (repeat.hpp:87,187)
    s.out_bytes[RAFTGPU_OUT_READS_FASTA] = (uint64_t)fasta_bytes;
    // record containing the first byte of every 16 KiB tile of reads.fasta
    CK(ctx->b_fasta_tile_frag.ensure(sizeof(int32_t) * (size_t)(fasta_bytes / FASTA_TILE + 2)));
    launch_fasta_tile_index(ctx->b_frag_off.as<int64_t>(), G, ctx->b_fasta_tile_frag.as<int32_t>(), ctx->st);
    CKL();
    CK(ctx->b_frag_desc.ensure(sizeof(FragDesc) * (size_t)(G + 1)));
    launch_frag_desc(ctx->b_frag_read.as<int32_t>(), ctx->b_frag_a.as<int32_t>(), ctx->b_frag_b.as<int32_t>(), ctx->b_frag_size.as<int32_t>(),
                     ctx->b_frag_off.as<int64_t>(), ctx->d_seq_off, G, ctx->b_frag_desc.as<FragDesc>(), ctx->st);
    CKL();
    if (ctx->seq_host) { // coarse output-offset -> arena-offset map for the chunked upload
        size_t cnt = (size_t)(G / FRAG_SAMPLE + 1);
        CK(ctx->b_frag_sample.ensure(sizeof(int64_t) * 2 * cnt));
        launch_frag_sample(ctx->b_frag_desc.as<FragDesc>(), G, FRAG_SAMPLE, ctx->b_frag_sample.as<int64_t>(), ctx->st);
        CKL();
        ctx->h_frag_sample.resize(2 * cnt);
        CK(cudaMemcpyAsync(ctx->h_frag_sample.data(), ctx->b_frag_sample.p, sizeof(int64_t) * 2 * cnt, cudaMemcpyDeviceToHost, ctx->st));
    }
    cudaEventRecord(ctx->ev[7], ctx->st);
    CK(cudaStreamSynchronize(ctx->st));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]); s.ms_layout = ms;
    s.kernel_launches = ctx->launches;
    ctx->sized = true;
    return RAFTGPU_OK;
}

extern "C" int raftgpu_run(raftgpu_ctx* ctx, raftgpu_stats* out)
{
    if (!ctx) return RAFTGPU_E_ARG;
    if (!ctx->have_reads) FAIL(RAFTGPU_E_STATE, "raftgpu_run: no reads");
    if (!ctx->paf_done) { // allow a PAF whose last chunk was not flagged: flush the carry
        int st = raftgpu_ingest_paf(ctx, nullptr, 0, 1);
        if (st) return st;
    }
    CK(cudaSetDevice(ctx->device));
    cudaEventRecord(ctx->ev[2], ctx->st);
    int st = accumulate_local(ctx);
    if (st) return st;
    cudaEventRecord(ctx->ev[3], ctx->st);
    if ((st = raftgpu_finalize(ctx, nullptr))) return st;
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]); ctx->stats.ms_scatter = ms;
    if ((st = layout_outputs(ctx))) return st;
    raftgpu_stats& s = ctx->stats;
    s.ms_total = s.ms_tokenize + s.ms_scatter + s.ms_scan + s.ms_repeat_cut + s.ms_layout;
    if (out) *out = s;
    return RAFTGPU_OK;
}


// ------------------------------------------------------------------------------------------------ sharded run inside the library (NCCL)
extern "C" int raftgpu_comm_unique_id(uint8_t id[RAFTGPU_COMM_ID_BYTES])
{
    static_assert(sizeof(ncclUniqueId) == RAFTGPU_COMM_ID_BYTES, "ncclUniqueId size");
    if (!id) return RAFTGPU_E_ARG;
    ncclUniqueId u;
    if (ncclGetUniqueId(&u) != ncclSuccess) return RAFTGPU_E_CUDA;
    memcpy(id, &u, sizeof u);
    return RAFTGPU_OK;
}

extern "C" int raftgpu_comm_init(raftgpu_ctx* ctx, int nranks, int rank, const uint8_t id[RAFTGPU_COMM_ID_BYTES])
{
    if (!ctx || !id || nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks) return RAFTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (ctx->comm) { ncclCommDestroy(ctx->comm); ctx->comm = nullptr; }
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    NCK(ncclCommInitRank(&ctx->comm, nranks, u, rank));
    ctx->nranks = nranks; ctx->rank = rank;
    CK(ctx->b_coll.ensure(sizeof(long long) * COLL_WORDS));
    if (!ctx->h_coll) CK(cudaHostAlloc((void**)&ctx->h_coll, sizeof(long long) * COLL_WORDS, cudaHostAllocDefault));
    return RAFTGPU_OK;
}

extern "C" int raftgpu_comm_destroy(raftgpu_ctx* ctx)
{
    if (!ctx) return RAFTGPU_E_ARG;
    if (ctx->comm) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->st); ncclCommDestroy(ctx->comm); ctx->comm = nullptr; }
    ctx->nranks = 1; ctx->rank = 0;
    return RAFTGPU_OK;
}

extern "C" int raftgpu_partition_reads(const int64_t* lengths, int64_t n, int32_t reso, int nranks, int64_t* bounds)
{
    if (!lengths || !bounds || n < 0 || reso < 1 || nranks < 1) return RAFTGPU_E_ARG;
    // boundary r = first read whose slots start at or after total * r / nranks (slots = bins + 1 per read)
    long long total = 0;
    for (int64_t i = 0; i < n; i++) total += (lengths[i] + reso - 1) / reso + 1;
    bounds[0] = 0;
    long long acc = 0;
    int64_t   i = 0;
    for (int r = 1; r < nranks; r++) {
        const long long want = total * r / nranks;
        while (i < n && acc < want) { acc += (lengths[i] + reso - 1) / reso + 1; i++; }
        bounds[r] = i;
    }
    bounds[nranks] = n;
    return RAFTGPU_OK;
}

// forget the PAF of this context (records, flags, differences) but keep the reads: used when the first-record peek has to be redone
static int reset_paf_state(raftgpu_ctx* ctx)
{
    Misc* M = ctx->misc();
    ctx->n_rec = 0; ctx->carry.clear(); ctx->paf_done = false; ctx->paf_bytes = 0; ctx->q_scattered = false; ctx->h_sym = 0;
    ctx->diff_zeroed = false; ctx->diff_zero_ints = 0; ctx->stats.ms_tokenize = 0; ctx->shard_begun = false;
    CK(cudaMemsetAsync(&M->sym_flag, 0, sizeof(int), ctx->st));
    CK(cudaMemsetAsync(&M->n_records_out, 0, sizeof(long long), ctx->st));
    return RAFTGPU_OK;
}

// record 0 of the whole file (chop.hpp:171-184 compares every later record with it): every rank peeks the head of its
// text, the first records are all-gathered and the first rank that has one wins -- all on the stream, the tokenizer reads
// the result from device memory.  `whole`: the head is the rank's whole text.
static int sharded_begin_impl(raftgpu_ctx* ctx, const int64_t* bounds, const uint8_t* head_text, size_t head, bool whole)
{
    if (!ctx->comm) FAIL(RAFTGPU_E_STATE, "sharded run: raftgpu_comm_init first");
    if (!ctx->have_reads || ctx->paf_done || ctx->n_rec || ctx->finalized || !ctx->carry.empty()) FAIL(RAFTGPU_E_STATE, "sharded run: set the reads of this run first");
    const int P = ctx->nranks, R = ctx->rank;
    if (bounds[R] != ctx->own_first || bounds[R + 1] != ctx->own_first + ctx->m || bounds[0] != 0 || bounds[P] != ctx->n)
        FAIL(RAFTGPU_E_ARG, "sharded run: bounds do not match the read range given to raftgpu_set_reads_sharded");
    CK(cudaSetDevice(ctx->device));
    ctx->shard_bounds.assign(bounds, bounds + P + 1);
    Misc*      M = ctx->misc();
    long long* coll = ctx->b_coll.as<long long>();
    int*       srec = reinterpret_cast<int*>(coll + COLL_SEND);
    CK(cudaMemsetAsync(srec, 0, sizeof(int) * 8, ctx->st));
    const uint8_t* dhead = head_text;
    if (head && !(is_device_ptr(head_text) && ((uintptr_t)head_text & 15) == 0)) {
        CK(ctx->b_peek.ensure(head + 64));
        CK(cudaMemcpyAsync(ctx->b_peek.p, head_text, head, cudaMemcpyDefault, ctx->st));
        dhead = ctx->b_peek.as<uint8_t>();
    }
    if (head) { launch_paf_peek(dhead, (int64_t)head, ctx->nt, srec, &M->err, ctx->st, whole ? 1 : 0); CKL(); }
    else { const int w = 1; CK(cudaMemcpyAsync(srec + 7, &w, sizeof(int), cudaMemcpyHostToDevice, ctx->st)); }
    int* grec = reinterpret_cast<int*>(coll + COLL_REC0);
    NCK(ncclAllGather(srec, grec, 8, ncclInt32, ctx->comm, ctx->st));
    launch_pick_rec0(grec, P, R, M->rec0, ctx->st);
    CKL();
    CK(cudaMemcpyAsync(ctx->h_coll + COLL_REC0, grec, sizeof(int) * 8 * (size_t)P, cudaMemcpyDeviceToHost, ctx->st));
    ctx->rec0_external = true; ctx->first_is_local = -1; ctx->shard_begun = true;
    return RAFTGPU_OK;
}
// Were the peeked heads enough?  Every rank before the first one that found a record must have peeked its whole text
// (all ranks read the same gathered flags, so all agree).  Valid once the stream has been synchronised after the begin.
static bool sharded_peek_ok(const raftgpu_ctx* ctx)
{
    const int* hrec = reinterpret_cast<const int*>(ctx->h_coll + COLL_REC0);
    for (int r = 0; r < ctx->nranks; r++) {
        if (hrec[8 * r + 6]) return true;
        if (!hrec[8 * r + 7]) return false;
    }
    return true;
}
static int sharded_finish_impl(raftgpu_ctx* ctx, int ing, raftgpu_stats* out, raftgpu_shard_info* info);

extern "C" int raftgpu_sharded_begin(raftgpu_ctx* ctx, const int64_t* bounds, const uint8_t* head_text, size_t head_bytes, int head_is_whole_text)
{
    if (!ctx || !bounds || (!head_text && head_bytes)) return RAFTGPU_E_ARG;
    ctx->shard = raftgpu_shard_info{};
    return sharded_begin_impl(ctx, bounds, head_text, head_bytes, head_is_whole_text != 0);
}

extern "C" int raftgpu_sharded_finish(raftgpu_ctx* ctx, raftgpu_stats* out, raftgpu_shard_info* info)
{
    if (!ctx) return RAFTGPU_E_ARG;
    if (!ctx->shard_begun) FAIL(RAFTGPU_E_STATE, "raftgpu_sharded_finish: raftgpu_sharded_begin first");
    int ing = ctx->pending_ingest_status;
    if (!ing && !ctx->paf_done) ing = raftgpu_ingest_paf(ctx, nullptr, 0, 1);
    if (ing != RAFTGPU_OK && ing != RAFTGPU_E_UNKNOWN_NAME && ing != RAFTGPU_E_RANGE) return ing;
    CK(cudaStreamSynchronize(ctx->st));
    if (!sharded_peek_ok(ctx)) FAIL(RAFTGPU_E_UNSUPPORTED, "the first chunk of a rank's PAF text holds no record although later chunks may: use raftgpu_run_sharded");
    return sharded_finish_impl(ctx, ing, out, info);
}

extern "C" int raftgpu_run_sharded(raftgpu_ctx* ctx, const int64_t* bounds, const uint8_t* text, size_t nbytes, raftgpu_stats* out,
                                   raftgpu_shard_info* info)
{
    if (!ctx || !bounds || (!text && nbytes)) return RAFTGPU_E_ARG;
    ctx->shard = raftgpu_shard_info{};
    // Should a rank before the winner have a record beyond its peeked head (a PAF whose first megabytes hold no record at
    // all), every rank sees that in the gathered flags and the step is redone with whole-text peeks.
    size_t peek_cap = 4u << 20;
    if (const char* e = getenv("RAFT_B200_PEEK_BYTES")) { long long v = atoll(e); if (v > 0) peek_cap = (size_t)v; } // test knob
    int ing = RAFTGPU_OK;
    for (int attempt = 0;; attempt++) {
        const size_t head = attempt == 0 ? std::min(nbytes, peek_cap) : nbytes;
        int          st = sharded_begin_impl(ctx, bounds, text, head, head == nbytes);
        if (st) return st;
        ing = raftgpu_ingest_paf(ctx, text, nbytes, 1); // synchronises the stream: the gathered peeks are on the host now
        if (ing != RAFTGPU_OK && ing != RAFTGPU_E_UNKNOWN_NAME && ing != RAFTGPU_E_RANGE) return ing; // not a data error: nothing to agree on
        if (sharded_peek_ok(ctx) || attempt == 1) break;
        ctx->shard.peek_retries++;
        if ((st = reset_paf_state(ctx))) return st;
    }
    return sharded_finish_impl(ctx, ing, out, info);
}

static int sharded_finish_impl(raftgpu_ctx* ctx, int ing, raftgpu_stats* out, raftgpu_shard_info* info)
{
    const int  P = ctx->nranks, R = ctx->rank;
    Misc*      M = ctx->misc();
    long long* coll = ctx->b_coll.as<long long>();
    const int64_t* bounds = ctx->shard_bounds.data();
    raftgpu_shard_info& sh = ctx->shard;
    sh.nranks = P; sh.rank = R;
    struct GatherGuard { raftgpu_ctx* c; ~GatherGuard() { c->gather_on = false; c->shard_begun = false; } } guard{ctx};
    ctx->gather_on = true;

    // ---- symmetric flag: OR over the ranks, in place on the device
    NCK(ncclAllReduce(&M->sym_flag, &M->sym_flag, 1, ncclInt32, ncclMax, ctx->comm, ctx->st));
    ctx->sym_external = true;
    cudaEventRecord(ctx->ev[8], ctx->st);
    // ---- endpoints for reads of other ranks: count (and collect) per destination, all-gather the count rows
    CK(ctx->b_bounds.ensure(sizeof(int64_t) * (MAX_RANKS + 1)));
    CK(ctx->b_route_list.ensure(sizeof(int4) * ROUTE_LIST_CAP));
    CK(cudaMemcpyAsync(ctx->b_bounds.p, bounds, sizeof(int64_t) * (P + 1), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemsetAsync(M->route_counts, 0, sizeof(unsigned long long) * 65, ctx->st));
    unsigned long long cap = ROUTE_LIST_CAP;
    if (const char* e = getenv("RAFT_B200_ROUTE_CAP")) { long long v = atoll(e); if (v >= 0 && (unsigned long long)v < cap) cap = (unsigned long long)v; }
    if (ing == RAFTGPU_OK) {
        launch_route_collect(scatter_args(ctx), P, ctx->b_bounds.as<int64_t>(), M->route_counts, ctx->b_route_list.as<int4>(), &M->route_list_n, cap, ctx->st);
        CKL();
    }
    const int       W = P + CNT_EXTRA;
    const long long extra[4] = {(long long)ctx->n_rec, (long long)ing, (long long)ctx->total_read_len, (long long)(ctx->n_slots - ctx->m)};
    launch_pack_counts(M->route_counts, &M->route_list_n, &M->sym_flag, extra, P, coll + COLL_SEND, ctx->st);
    CKL();
    NCK(ncclAllGather(coll + COLL_SEND, coll + COLL_CNT, (size_t)W, ncclInt64, ctx->comm, ctx->st));
    CK(cudaMemcpyAsync(ctx->h_coll + COLL_CNT, coll + COLL_CNT, sizeof(long long) * (size_t)W * P, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    const long long* g = ctx->h_coll + COLL_CNT;
    int              st = ing;
    sh.n_records_total = sh.total_read_len = sh.n_bins_total = 0;
    for (int r = 0; r < P; r++) {
        if (!st && g[(size_t)r * W + P + 3]) st = RAFTGPU_E_PEER;
        sh.n_records_total += g[(size_t)r * W + P + 2]; sh.total_read_len += g[(size_t)r * W + P + 4]; sh.n_bins_total += g[(size_t)r * W + P + 5];
    }
    if (st == RAFTGPU_E_PEER) ctx->last_error = "another rank reported a data error while reading its part of the PAF";
    if (st) return st;
    if (sh.n_records_total > 0x7fffffffll) FAIL(RAFTGPU_E_ARG, "more than 2^31-1 PAF records (the reference counts them in an int, chop.hpp:139)");
    const int S = (int)g[(size_t)R * W + P + 1];
    ctx->h_sym = S; sh.symmetric = S;
    long long soff[MAX_RANKS + 1], roff[MAX_RANKS + 1];
    soff[0] = roff[0] = 0;
    for (int r = 0; r < P; r++) { soff[r + 1] = soff[r] + g[(size_t)R * W + r]; roff[r + 1] = roff[r] + g[(size_t)r * W + R]; }
    sh.endpoints_sent = soff[P]; sh.endpoints_received = roff[P];
    const long long listed = g[(size_t)R * W + P];
    // ---- intervals on reads this rank owns go straight into its differences; the others are bucketed by owner and exchanged
    if ((st = accumulate_local(ctx))) return st;
    CK(ctx->b_send.ensure(sizeof(int32_t) * 3 * (size_t)(soff[P] + 1))); CK(ctx->b_recv.ensure(sizeof(int32_t) * 3 * (size_t)(roff[P] + 1)));
    if (soff[P] > 0) {
        unsigned long long cur[64] = {0};
        for (int r = 0; r < P; r++) cur[r] = (unsigned long long)soff[r];
        CK(cudaMemcpyAsync(M->route_counts, cur, sizeof cur, cudaMemcpyHostToDevice, ctx->st));
        if ((unsigned long long)listed <= cap) launch_route_pack_list(ctx->b_route_list.as<int4>(), listed, M->route_counts, ctx->b_send.as<int32_t>(), ctx->st);
        else launch_route_pack(scatter_args(ctx), P, ctx->b_bounds.as<int64_t>(), M->route_counts, ctx->b_send.as<int32_t>(), ctx->st);
        CKL();
    }
    NCK(ncclGroupStart());
    for (int r = 0; r < P; r++) {
        if (r == R) continue;
        const long long ns = soff[r + 1] - soff[r], nr = roff[r + 1] - roff[r];
        if (ns) NCK(ncclSend(ctx->b_send.as<int32_t>() + 3 * soff[r], (size_t)(3 * ns), ncclInt32, r, ctx->comm, ctx->st));
        if (nr) NCK(ncclRecv(ctx->b_recv.as<int32_t>() + 3 * roff[r], (size_t)(3 * nr), ncclInt32, r, ctx->comm, ctx->st));
    }
    NCK(ncclGroupEnd());
    if (roff[P] > 0) {
        if ((st = zero_diff(ctx))) return st;
        launch_scatter_endpoints(ctx->b_recv.as<int32_t>(), roff[P], ctx->b_slot_off.as<int64_t>(), ctx->b_diff.as<int32_t>(), ctx->prm.reso, ctx->own_first,
                                 ctx->m, &M->err, ctx->st);
        CKL();
    }
    cudaEventRecord(ctx->ev[9], ctx->st);
    if ((st = start_seq_upload(ctx))) return st; // deferred arena upload (no-op otherwise): behind the exchange, beside everything that follows
    // ---- local coverage / repeats / cut points; global numbering and file offsets ride on the synchronisations of these two
    if ((st = raftgpu_finalize(ctx, nullptr))) return st;
    if ((st = layout_outputs(ctx))) return st;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]) == cudaSuccess) sh.ms_exchange = ms;
    raftgpu_stats& s = ctx->stats;
    s.ms_scatter = ms;
    s.ms_total = s.ms_tokenize + s.ms_scatter + s.ms_scan + s.ms_repeat_cut + s.ms_layout;
    if (out) *out = s;
    if (info) *info = sh;
    return RAFTGPU_OK;
}

// ------------------------------------------------------------------------------------------------ read tables of a context
extern "C" int raftgpu_reads_info(raftgpu_ctx* ctx, int64_t* n_local, int64_t* name_bytes, int64_t* seq_bytes)
{
    if (!ctx) return RAFTGPU_E_ARG;
    if (!ctx->have_reads) FAIL(RAFTGPU_E_STATE, "raftgpu_reads_info: no reads");
    CK(cudaSetDevice(ctx->device));
    int64_t no[2] = {0, 0};
    CK(cudaMemcpy(&no[0], ctx->d_name_off + ctx->own_first, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&no[1], ctx->d_name_off + ctx->own_first + ctx->m, 8, cudaMemcpyDeviceToHost));
    if (n_local) *n_local = ctx->m;
    if (name_bytes) *name_bytes = no[1] - no[0];
    if (seq_bytes) *seq_bytes = ctx->total_read_len;
    return RAFTGPU_OK;
}

extern "C" int raftgpu_reads_copy(raftgpu_ctx* ctx, int64_t* name_off, uint8_t* names, int64_t* lengths)
{
    if (!ctx) return RAFTGPU_E_ARG;
    if (!ctx->have_reads) FAIL(RAFTGPU_E_STATE, "raftgpu_reads_copy: no reads");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->st));
    const int64_t m = ctx->m;
    std::vector<int64_t> no((size_t)m + 1);
    CK(cudaMemcpy(no.data(), ctx->d_name_off + ctx->own_first, sizeof(int64_t) * (size_t)(m + 1), cudaMemcpyDeviceToHost));
    if (names && no[m] > no[0]) CK(cudaMemcpy(names, ctx->d_names + no[0], (size_t)(no[m] - no[0]), cudaMemcpyDeviceToHost));
    if (name_off) for (int64_t i = 0; i <= m; i++) name_off[i] = no[i] - no[0];
    if (lengths) {
        std::vector<int64_t> so((size_t)m + 1);
        CK(cudaMemcpy(so.data(), ctx->d_seq_off, sizeof(int64_t) * (size_t)(m + 1), cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < m; i++) lengths[i] = so[i + 1] - so[i];
    }
    return RAFTGPU_OK;
}

extern "C" int raftgpu_reads_device(raftgpu_ctx* ctx, const int64_t** seq_off, const uint8_t** seq)
{
    if (!ctx || !seq_off || !seq) return RAFTGPU_E_ARG;
    if (!ctx->have_reads || !ctx->have_seq) FAIL(RAFTGPU_E_STATE, "raftgpu_reads_device: no reads with sequence bytes");
    *seq_off = ctx->d_seq_off; *seq = ctx->d_seq;
    return RAFTGPU_OK;
}

// ------------------------------------------------------------------------------------------------ outputs
// materialise stream bytes [w0, w1) of `which` at device address d (d[k] = stream byte w0+k)
static int emit_window_impl(raftgpu_ctx* ctx, int which, int64_t w0, int64_t w1, uint8_t* d, cudaStream_t st);
static inline int lane_of(int which) { return (which == RAFTGPU_OUT_READS_FASTA || which == RAFTGPU_OUT_SPLIT_NAIVE) ? 1 : 0; }
// close a lane's timing span: device time between its first launch since the last flush and its last launch
static void lane_flush(raftgpu_ctx* ctx, int li)
{
    auto& L = ctx->lane[li];
    if (!L.open) return;
    cudaEventSynchronize(L.ev1);
    float ms = 0;
    if (cudaEventElapsedTime(&ms, L.ev0, L.ev1) == cudaSuccess) ctx->stats.ms_emit[L.kind] += ms;
    L.open = false;
}
static void flush_emit_timer(raftgpu_ctx* ctx) { lane_flush(ctx, 0); lane_flush(ctx, 1); }
// launch the emitter of stream `which` for bytes [w0, w1) on its lane's CUDA stream (no host synchronisation)
static int emit_window(raftgpu_ctx* ctx, int which, int64_t w0, int64_t w1, uint8_t* d)
{
    if (w1 <= w0) return RAFTGPU_OK;
    const int li = lane_of(which), kind = which > 3 ? 3 : which; // the split_naive stream is accounted with the gather kernel
    auto&     L = ctx->lane[li];
    if (L.open && L.kind != kind) lane_flush(ctx, li);
    if (!L.open) { cudaEventRecord(L.ev0, L.st); L.open = true; L.kind = kind; }
    int rc = emit_window_impl(ctx, which, w0, w1, d, L.st);
    cudaEventRecord(L.ev1, L.st);
    ctx->stats.emit_launches[kind]++; ctx->stats.emit_bytes[kind] += (uint64_t)(w1 - w0);
    return rc;
}
static int emit_window_impl(raftgpu_ctx* ctx, int which, int64_t w0, int64_t w1, uint8_t* d, cudaStream_t st)
{
    const int64_t m = ctx->m;
    if (which == RAFTGPU_OUT_COVERAGE) {
        // coarse tile range from the sampled offsets; tiles outside the window return early
        const auto& off = ctx->h_cov_tile_off;
        const int64_t T = cov_tiles(ctx->n_slots);
        int64_t t0 = (std::upper_bound(off.begin(), off.end(), w0) - off.begin() - 1) * OFF_SAMPLE;
        int64_t t1 = (std::lower_bound(off.begin(), off.end(), w1) - off.begin()) * OFF_SAMPLE;
        t0 = std::max<int64_t>(0, std::min(t0, T)); t1 = std::min(t1, T);
        CovEmitArgs ca{};
        ca.cov = ctx->b_cov.as<int32_t>(); ca.slot_off = ctx->b_slot_off.as<int64_t>(); ca.m = m; ca.n_slots = ctx->n_slots;
        ca.own_first = ctx->own_first; ca.reso = ctx->prm.reso; ca.tile_off = ctx->b_cov_tile_off.as<int64_t>();
        ca.dst = d; ca.w0 = w0; ca.w1 = w1; ca.tile_first = t0; ca.tile_read = ctx->b_cov_tile_read.as<int32_t>();
        if (!ctx->cov_tabs_built) {
            CK(ctx->b_cov_tabs.ensure(sizeof(unsigned long long) * COV_POS_TAB_ENTRIES + sizeof(unsigned) * COV_TAB_ENTRIES));
            launch_cov_tables(ctx->b_cov_tabs.as<unsigned long long>(), COV_POS_TAB_ENTRIES, ctx->prm.reso,
                              reinterpret_cast<unsigned*>(ctx->b_cov_tabs.as<unsigned long long>() + COV_POS_TAB_ENTRIES), st);
            CKL();
            ctx->cov_tabs_built = true;
        }
        ca.pos_tab = ctx->b_cov_tabs.as<unsigned long long>(); ca.tab_n = COV_POS_TAB_ENTRIES;
        ca.cov_tab = reinterpret_cast<const unsigned*>(ctx->b_cov_tabs.as<unsigned long long>() + COV_POS_TAB_ENTRIES);
        launch_cov_emit(ca, t1 - t0, st);
        CKL();
    } else if (which == RAFTGPU_OUT_LONG_REPEATS) {
        const auto& off = ctx->h_rep_line_off;
        int64_t r0 = (std::upper_bound(off.begin(), off.end(), w0) - off.begin() - 1) * OFF_SAMPLE;
        int64_t r1 = (std::lower_bound(off.begin(), off.end(), w1) - off.begin()) * OFF_SAMPLE;
        r0 = std::max<int64_t>(0, std::min(r0, m));
        RepEmitArgs ra{};
        ra.rep_cnt = ctx->b_rep_cnt.as<int32_t>(); ra.rep_cap_off = ctx->b_rep_cap_off.as<int64_t>(); ra.rep = ctx->b_rep.as<int2>();
        ra.line_off = ctx->b_rep_line_off.as<int64_t>(); ra.m = m; ra.own_first = ctx->own_first; ra.dst = d; ra.w0 = w0; ra.w1 = w1;
        ra.read_first = r0; ra.read_last = std::min<int64_t>(r1, m);
        launch_rep_emit(ra, st);
        CKL();
    } else if (which == RAFTGPU_OUT_SPLIT_NAIVE) {
        int rc = wait_seq_bytes(ctx, (int64_t)ctx->seq_host_bytes, st); // needs the whole arena (no output->arena map kept for this stream)
        if (rc) return rc;
        FastaEmitArgs fa{};
        fa.desc = ctx->b_sn_desc.as<FragDesc>(); fa.G = ctx->sn_G; fa.seq = ctx->d_seq; fa.seq_off = ctx->d_seq_off;
        fa.names = ctx->d_names; fa.name_off = ctx->d_name_off; fa.own_first = ctx->own_first; fa.read_num_base = 0;
        fa.dst = d; fa.w0 = w0; fa.w1 = w1; fa.tile_frag = ctx->b_sn_tile.as<int32_t>();
        fa.seq_safe_end = (ctx->d_seq == ctx->b_seq.as<uint8_t>()) ? ((ctx->total_read_len + 31) & ~(int64_t)15) : (ctx->total_read_len & ~(int64_t)15);
        fa.sim = nullptr; fa.split_len = ctx->sn_len;
        launch_fasta_emit(fa, st);
        CKL();
    } else if (which == RAFTGPU_OUT_BED) {
        if (ctx->real_reads) return RAFTGPU_OK;
        const auto& off = ctx->h_bed_line_off;
        int64_t r0 = (std::upper_bound(off.begin(), off.end(), w0) - off.begin() - 1) * OFF_SAMPLE;
        int64_t r1 = (std::lower_bound(off.begin(), off.end(), w1) - off.begin()) * OFF_SAMPLE;
        r0 = std::max<int64_t>(0, std::min(r0, m));
        RepEmitArgs ra{};
        ra.rep_cnt = ctx->b_rep_cnt.as<int32_t>(); ra.rep_cap_off = ctx->b_rep_cap_off.as<int64_t>(); ra.rep = ctx->b_rep.as<int2>();
        ra.line_off = ctx->b_bed_line_off.as<int64_t>(); ra.m = m; ra.own_first = ctx->own_first; ra.dst = d; ra.w0 = w0; ra.w1 = w1;
        ra.read_first = r0; ra.read_last = std::min<int64_t>(r1, m);
        ra.sim = ctx->b_sim.as<SimInfo>(); ra.names = ctx->d_names; ra.name_off = ctx->d_name_off;
        launch_bed_emit(ra, st);
        CKL();
    } else if (which == RAFTGPU_OUT_READS_FASTA) {
        if (!ctx->have_seq) FAIL(RAFTGPU_E_STATE, "reads.fasta requested but no sequence bytes were given");
        if (ctx->seq_host) {
            // arena bytes needed by stream window [w0, w1): records are in arena order, consecutive records of a
            // read overlap by at most overlap_length bytes
            int64_t need = (int64_t)ctx->seq_host_bytes;
            const auto& sm = ctx->h_frag_sample;
            size_t lo = 0, hi = sm.size() / 2; // first sample with out_off >= w1
            while (lo < hi) { size_t mid = (lo + hi) / 2; if (sm[2 * mid] >= w1) hi = mid; else lo = mid + 1; }
            if (lo < sm.size() / 2) need = std::min<int64_t>(need, sm[2 * lo + 1] + std::max(0, ctx->prm.overlap_length) + 64);
            int rc = wait_seq_bytes(ctx, need, st);
            if (rc) return rc;
        }
        FastaEmitArgs fa{};
        fa.desc = ctx->b_frag_desc.as<FragDesc>(); fa.G = ctx->G; fa.seq = ctx->d_seq; fa.seq_off = ctx->d_seq_off;
        fa.names = ctx->d_names; fa.name_off = ctx->d_name_off; fa.own_first = ctx->own_first; fa.read_num_base = ctx->read_num_base;
        fa.dst = d; fa.w0 = w0; fa.w1 = w1; fa.tile_frag = ctx->b_fasta_tile_frag.as<int32_t>();
        fa.sim = ctx->real_reads ? nullptr : ctx->b_sim.as<SimInfo>();
        // owned arenas are padded by 32 bytes; borrowed ones are only trusted up to their last whole 16-byte block
        fa.seq_safe_end = (ctx->d_seq == ctx->b_seq.as<uint8_t>()) ? ((ctx->total_read_len + 31) & ~(int64_t)15) : (ctx->total_read_len & ~(int64_t)15);
        launch_fasta_emit(fa, st);
        CKL();
    }
    return RAFTGPU_OK;
}

extern "C" int raftgpu_split_naive(raftgpu_ctx* ctx, int32_t sublen)
{
    if (!ctx || sublen < 1) return RAFTGPU_E_ARG;
    if (!ctx->have_reads || !ctx->have_seq) FAIL(RAFTGPU_E_STATE, "raftgpu_split_naive: set reads (with sequence bytes) first");
    CK(cudaSetDevice(ctx->device));
    const int64_t m = ctx->m;
    Misc*         M = ctx->misc();
    CK(ctx->b_sn_cnt.ensure(sizeof(int32_t) * (size_t)(m + 1))); CK(ctx->b_sn_base.ensure(sizeof(int64_t) * (size_t)(m + 1)));
    CK(ctx->b_status.ensure(sizeof(uint64_t) * (size_t)(scan_tiles_small(m) + 8)));
    launch_split_counts(ctx->d_seq_off, m, sublen, ctx->b_sn_cnt.as<int32_t>(), ctx->st);
    CKL();
    launch_scan_i32_to_i64(ctx->b_sn_cnt.as<int32_t>(), ctx->b_sn_base.as<int64_t>(), m, ctx->b_status.as<uint64_t>(), &M->ticket, ctx->st);
    CKL();
    long long G = 0;
    CK(cudaMemcpyAsync(&G, ctx->b_sn_base.as<int64_t>() + m, 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    CK(ctx->b_sn_read.ensure(4 * (size_t)(G + 1))); CK(ctx->b_sn_a.ensure(4 * (size_t)(G + 1))); CK(ctx->b_sn_b.ensure(4 * (size_t)(G + 1)));
    CK(ctx->b_sn_size.ensure(4 * (size_t)(G + 1))); CK(ctx->b_sn_off.ensure(8 * (size_t)(G + 1))); CK(ctx->b_sn_desc.ensure(sizeof(FragDesc) * (size_t)(G + 1)));
    CK(ctx->b_status.ensure(sizeof(uint64_t) * (size_t)(scan_tiles_small(G) + 8)));
    launch_split_expand(ctx->d_seq_off, ctx->d_name_off, ctx->own_first, m, sublen, ctx->b_sn_base.as<int64_t>(), ctx->b_sn_read.as<int32_t>(),
                        ctx->b_sn_a.as<int32_t>(), ctx->b_sn_b.as<int32_t>(), ctx->b_sn_size.as<int32_t>(), ctx->st);
    CKL();
    launch_scan_i32_to_i64(ctx->b_sn_size.as<int32_t>(), ctx->b_sn_off.as<int64_t>(), G, ctx->b_status.as<uint64_t>(), &M->ticket, ctx->st);
    CKL();
    long long bytes = 0;
    CK(cudaMemcpyAsync(&bytes, ctx->b_sn_off.as<int64_t>() + G, 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    CK(ctx->b_sn_tile.ensure(sizeof(int32_t) * (size_t)(bytes / FASTA_TILE + 2)));
    launch_fasta_tile_index(ctx->b_sn_off.as<int64_t>(), G, ctx->b_sn_tile.as<int32_t>(), ctx->st);
    CKL();
    launch_frag_desc(ctx->b_sn_read.as<int32_t>(), ctx->b_sn_a.as<int32_t>(), ctx->b_sn_b.as<int32_t>(), ctx->b_sn_size.as<int32_t>(),
                     ctx->b_sn_off.as<int64_t>(), ctx->d_seq_off, G, ctx->b_sn_desc.as<FragDesc>(), ctx->st);
    CKL();
    CK(cudaStreamSynchronize(ctx->st));
    ctx->sn_len = sublen; ctx->sn_G = G; ctx->sn_bytes = (uint64_t)bytes;
    return RAFTGPU_OK;
}

extern "C" int raftgpu_set_option(raftgpu_ctx* ctx, int option, int64_t value)
{
    if (!ctx) return RAFTGPU_E_ARG;
    if (option == RAFTGPU_OPT_DEFER_SEQ_UPLOAD) { ctx->opt_defer_seq = value != 0; return RAFTGPU_OK; }
    return RAFTGPU_E_ARG;
}

extern "C" int raftgpu_get_stats(raftgpu_ctx* ctx, raftgpu_stats* out)
{
    if (!ctx || !out) return RAFTGPU_E_ARG;
    flush_emit_timer(ctx);
    ctx->stats.kernel_launches = ctx->launches;
    *out = ctx->stats;
    return RAFTGPU_OK;
}

extern "C" int raftgpu_output_size(raftgpu_ctx* ctx, int which, uint64_t* nbytes)
{
    if (!ctx || !nbytes || which < 0 || which > 4) return RAFTGPU_E_ARG;
    if (which == RAFTGPU_OUT_SPLIT_NAIVE) {
        if (!ctx->sn_len) FAIL(RAFTGPU_E_STATE, "call raftgpu_split_naive first");
        *nbytes = ctx->sn_bytes;
        return RAFTGPU_OK;
    }
    int st = layout_outputs(ctx);
    if (st) return st;
    *nbytes = ctx->stats.out_bytes[which];
    return RAFTGPU_OK;
}

extern "C" int raftgpu_fetch(raftgpu_ctx* ctx, int which, uint64_t off, uint8_t* dst, size_t n)
{
    if (!ctx || which < 0 || which > 4 || (!dst && n)) return RAFTGPU_E_ARG;
    int st = RAFTGPU_OK;
    uint64_t stream_bytes = 0;
    if ((st = raftgpu_output_size(ctx, which, &stream_bytes))) return st;
    if (off > stream_bytes || n > stream_bytes - off) return RAFTGPU_E_ARG;
    if (!n) return RAFTGPU_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t ls = ctx->lane[lane_of(which)].st;
    if (is_device_ptr(dst)) {
        if ((st = emit_window(ctx, which, (int64_t)off, (int64_t)(off + n), dst))) return st;
        CK(cudaStreamSynchronize(ls));
        return fetch_err(ctx);
    }
    // host destination: double-buffered windows, emit on st, copy out on st2
    const size_t W = std::min<size_t>(WINDOW_BYTES, n);
    for (int k = 0; k < 2; k++) CK(ctx->b_stage[k].ensure(W + 64));
    size_t done = 0;
    int    k = 0;
    bool   used[2] = {false, false};
    while (done < n) {
        size_t len = std::min(W, n - done);
        if (used[k]) CK(cudaEventSynchronize(ctx->ev_stage[k])); // the copy out of this buffer has finished
        if ((st = emit_window(ctx, which, (int64_t)(off + done), (int64_t)(off + done + len), ctx->b_stage[k].as<uint8_t>()))) return st;
        CK(cudaStreamSynchronize(ls));
        CK(cudaMemcpyAsync(dst + done, ctx->b_stage[k].p, len, cudaMemcpyDeviceToHost, ctx->st2));
        CK(cudaEventRecord(ctx->ev_stage[k], ctx->st2));
        used[k] = true;
        done += len; k ^= 1;
    }
    CK(cudaStreamSynchronize(ctx->st2));
    return fetch_err(ctx);
}

extern "C" int raftgpu_fetch_async(raftgpu_ctx* ctx, int which, uint64_t off, uint8_t* dst_device, size_t n)
{
    if (!ctx || which < 0 || which > 4 || (!dst_device && n)) return RAFTGPU_E_ARG;
    uint64_t stream_bytes = 0;
    int      st = raftgpu_output_size(ctx, which, &stream_bytes);
    if (st) return st;
    if (off > stream_bytes || n > stream_bytes - off) return RAFTGPU_E_ARG;
    if (!n) return RAFTGPU_OK;
    if (!is_device_ptr(dst_device)) FAIL(RAFTGPU_E_ARG, "raftgpu_fetch_async needs a device destination");
    CK(cudaSetDevice(ctx->device));
    return emit_window(ctx, which, (int64_t)off, (int64_t)(off + n), dst_device);
}

extern "C" int raftgpu_sync(raftgpu_ctx* ctx)
{
    if (!ctx) return RAFTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->st_aux));
    CK(cudaStreamSynchronize(ctx->st));
    flush_emit_timer(ctx);
    return fetch_err(ctx);
}

extern "C" int raftgpu_digest(raftgpu_ctx* ctx, int which, uint64_t* digest) { return raftgpu_digest_at(ctx, which, 0, digest); }

extern "C" int raftgpu_digest_at(raftgpu_ctx* ctx, int which, uint64_t stream_base, uint64_t* digest)
{
    if (!ctx || !digest || which < 0 || which > 4) return RAFTGPU_E_ARG;
    uint64_t total = 0;
    int st = raftgpu_output_size(ctx, which, &total);
    if (st) return st;
    CK(cudaSetDevice(ctx->device));
    const size_t   W = (size_t)std::min<uint64_t>(WINDOW_BYTES, std::max<uint64_t>(total, 1));
    CK(ctx->b_stage[0].ensure(W + 64));
    cudaStream_t ls = ctx->lane[lane_of(which)].st;
    CK(cudaStreamSynchronize(ctx->st)); CK(cudaStreamSynchronize(ctx->st_aux));
    CK(cudaMemsetAsync(&ctx->misc()->digest, 0, sizeof(unsigned long long), ls));
    for (uint64_t off = 0; off < total; off += W) {
        size_t len = (size_t)std::min<uint64_t>(W, total - off);
        if ((st = emit_window(ctx, which, (int64_t)off, (int64_t)(off + len), ctx->b_stage[0].as<uint8_t>()))) return st;
        launch_digest(ctx->b_stage[0].as<uint8_t>(), (int64_t)len, (int64_t)(stream_base + off), &ctx->misc()->digest, ls);
        CKL();
    }
    unsigned long long d = 0;
    CK(cudaMemcpyAsync(&d, &ctx->misc()->digest, sizeof d, cudaMemcpyDeviceToHost, ls));
    CK(cudaStreamSynchronize(ls));
    if ((st = fetch_err(ctx))) return st;
    *digest = d;
    return RAFTGPU_OK;
}

extern "C" int raftgpu_fetch_table(raftgpu_ctx* ctx, int table, void* dst, size_t cap_bytes, size_t* n_elems)
{
    if (!ctx || table < 0 || table >= RAFTGPU_TAB_COUNT || !n_elems) return RAFTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->st));
    const int64_t m = ctx->m;
    size_t elem = 4, count = 0;
    std::vector<uint8_t> host; // tables assembled on the host
    const void* dsrc = nullptr;
    if (table <= RAFTGPU_TAB_STRAND) {
        const DevBuf* cols[] = {&ctx->b_qid, &ctx->b_tid, &ctx->b_qs, &ctx->b_qe, &ctx->b_ts, &ctx->b_te, &ctx->b_strand};
        elem = table == RAFTGPU_TAB_STRAND ? 1 : 4; count = (size_t)ctx->n_rec; dsrc = cols[table]->p;
    } else if (table == RAFTGPU_TAB_BIN_OFF || table == RAFTGPU_TAB_COV) {
        if (!ctx->have_reads) FAIL(RAFTGPU_E_STATE, "no reads");
        std::vector<int64_t> so(m + 1);
        CK(cudaMemcpy(so.data(), ctx->b_slot_off.p, sizeof(int64_t) * (m + 1), cudaMemcpyDeviceToHost));
        if (table == RAFTGPU_TAB_BIN_OFF) {
            elem = 8; count = (size_t)m + 1; host.resize(count * 8);
            int64_t* o = (int64_t*)host.data();
            for (int64_t i = 0; i <= m; i++) o[i] = so[i] - i;
        } else {
            if (!ctx->finalized) FAIL(RAFTGPU_E_STATE, "coverage requested before raftgpu_run");
            elem = 4; count = (size_t)(ctx->n_slots - m);
            if (dst) {
                std::vector<int32_t> slots((size_t)ctx->n_slots);
                if (ctx->n_slots) CK(cudaMemcpy(slots.data(), ctx->b_cov.p, sizeof(int32_t) * (size_t)ctx->n_slots, cudaMemcpyDeviceToHost));
                host.resize(count * 4);
                int32_t* o = (int32_t*)host.data();
                for (int64_t i = 0; i < m; i++) { int64_t nb = so[i + 1] - so[i] - 1; if (nb) memcpy(o + (so[i] - i), slots.data() + so[i], sizeof(int32_t) * (size_t)nb); }
            }
        }
    } else if (table == RAFTGPU_TAB_REP_OFF) {
        if (!ctx->finalized) FAIL(RAFTGPU_E_STATE, "repeats requested before raftgpu_run");
        elem = 8; count = (size_t)m + 1; dsrc = ctx->b_rep_off.p;
    } else if (table == RAFTGPU_TAB_REP) {
        if (!ctx->finalized) FAIL(RAFTGPU_E_STATE, "repeats requested before raftgpu_run");
        elem = 4; count = (size_t)ctx->n_repeats * 2;
        if (dst && count) {
            DevBuf tmp;
            CK(tmp.ensure(count * 4));
            launch_rep_compact(ctx->b_rep_cnt.as<int32_t>(), ctx->b_rep_cap_off.as<int64_t>(), ctx->b_rep_off.as<int64_t>(), ctx->b_rep.as<int2>(), m,
                               tmp.as<int32_t>(), ctx->st);
            CKL();
            host.resize(count * 4);
            CK(cudaMemcpyAsync(host.data(), tmp.p, count * 4, cudaMemcpyDeviceToHost, ctx->st));
            CK(cudaStreamSynchronize(ctx->st));
        }
    } else if (table == RAFTGPU_TAB_FRAG) {
        int st = layout_outputs(ctx);
        if (st) return st;
        elem = 4; count = (size_t)ctx->G * 3;
        if (dst && count) {
            std::vector<int32_t> r(ctx->G), a(ctx->G), b(ctx->G);
            CK(cudaMemcpy(r.data(), ctx->b_frag_read.p, 4 * (size_t)ctx->G, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(a.data(), ctx->b_frag_a.p, 4 * (size_t)ctx->G, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(b.data(), ctx->b_frag_b.p, 4 * (size_t)ctx->G, cudaMemcpyDeviceToHost));
            host.resize(count * 4);
            int32_t* o = (int32_t*)host.data();
            for (int64_t g = 0; g < ctx->G; g++) { o[3 * g] = (int32_t)(r[g] + ctx->own_first); o[3 * g + 1] = a[g]; o[3 * g + 2] = b[g]; }
        }
    }
    *n_elems = count;
    if (!dst) return RAFTGPU_OK;
    if (cap_bytes < count * elem) return RAFTGPU_E_ARG;
    if (count) {
        if (!host.empty()) CK(cudaMemcpy(dst, host.data(), count * elem, cudaMemcpyDefault));
        else CK(cudaMemcpy(dst, dsrc, count * elem, cudaMemcpyDefault));
    }
    return RAFTGPU_OK;
}
