// kernels.h — host-visible launchers of the raft_b200 CUDA kernels (internal; the public boundary
// is include/raft_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nametable.cuh"

namespace raftk {

// device-side error codes (same numbering as raftgpu_status)
enum { RAFTK_E_UNKNOWN_NAME = -2, RAFTK_E_DUP_NAME = -3, RAFTK_E_RANGE = -4, RAFTK_E_NEG_START = -5, RAFTK_E_SIM_NAME = -13, RAFTK_E_HASH_COLLISION = -100 };

// First error of a run: (index << 8) | (-code) in ONE word, so that the smallest offending record / read index and
// its own code are published together by a single atomicMin (ERR_CLEAN when there is none).
struct ErrState {
    unsigned long long packed;
    unsigned long long pad;
};
constexpr unsigned long long ERR_CLEAN = ~0ull;
inline int       err_code(const ErrState& e) { return -(int)(e.packed & 0xffull); }
inline long long err_index(const ErrState& e) { return (long long)(e.packed >> 8); }
#ifdef __CUDACC__
__device__ __forceinline__ void err_min(ErrState* err, int code, long long index)
{
    atomicMin(&err->packed, ((unsigned long long)index << 8) | (unsigned long long)(unsigned)(-code));
}
#endif

// ---------------------------------------------------------------- name table (nametable.cu)
cudaError_t launch_name_build(NameTable* out_table_desc_host, void* slots, unsigned long long capacity, unsigned long long seed,
                              const uint8_t* names, const int64_t* name_off, int64_t n, ErrState* err, cudaStream_t st);

// ---------------------------------------------------------------- K0 (k0_fasta.cu)
struct FastaTokArgs {
    const uint8_t* text;      // chunk of FASTA text starting at a line start, 16-byte aligned
    int64_t        nbytes;
    int            n_tiles, first_chunk, last_chunk;
    uint8_t*       seq_out;   // arena position of this chunk's first kept byte
    int64_t        seq_off_base; // arena offset of seq_out
    int64_t*       rec_pos;   // chunk-relative position of each record's marker   (this chunk's records)
    int64_t*       seq_off;   // arena offset of each record's first base
    int64_t        rec_cap;
    uint64_t *     st_carry, *st_keep, *st_rec; // n_tiles words each, zeroed
    int*           ticket;    // zeroed
    int*           flags;     // 1 CR before newline, 2 '+' line (FASTQ), 4 no leading marker, 8 marker is the last byte,
                              // 16 not a strict four-line FASTQ (fastq mode)
    long long*     totals;    // [0] records, [1] kept bytes, [2] newlines (fastq mode) of this chunk
    // strict four-line FASTQ mode: line k of the file has role k & 3 (0 '@' header, 1 bases, 2 '+' line, 3 qualities)
    int            fastq;
    int64_t        line_base;  // lines before this chunk
    int64_t        text_gbase; // file offset of the chunk's first byte
    uint64_t*      st_nl;      // n_tiles words, zeroed
    int64_t *      rec_gpos, *qual_gpos; // per record of the FILE (index = line / 4): file offsets of its '@' and of its quality line
    int64_t        rec_gcap;
};
// strict FASTQ: every record's quality line is as long as its bases and ends right before the next record (or the file)
void launch_fastq_verify(int64_t n, const int64_t* seq_off, const int64_t* rec_gpos, const int64_t* qual_gpos, int64_t file_end, int* flags,
                         cudaStream_t st);
int         fasta_tokenize_tiles(int64_t nbytes);
cudaError_t launch_fasta_tokenize(const FastaTokArgs& a, cudaStream_t st);
void launch_fasta_name_len(const uint8_t* text, int64_t nbytes, const int64_t* rec_pos, int64_t n, int32_t* name_len, cudaStream_t st);
void launch_add_offset_i64(const int64_t* src, int64_t n, int64_t add, int64_t* dst, cudaStream_t st);
void launch_fasta_name_copy(const uint8_t* text, const int64_t* rec_pos, const int64_t* name_off, int64_t n, uint8_t* names, cudaStream_t st);

// ---------------------------------------------------------------- K1 (k1_paf.cu)
struct PafTokArgs {
    const uint8_t* text;
    int64_t        nbytes;
    int64_t        rec_base, rec_cap;
    int32_t *      qid, *tid, *qs, *qe, *ts, *te;
    uint8_t*       strand;
    const int*     rec0;          // device int[8]: qid,tid,qs,qe,ts,te,present,first_is_local
    int            first_is_local; // record index 0 of this context is the file's record 0 (1 / 0); -1: read rec0[7] on the device
    int            n_tiles;
    uint64_t*      status;        // n_tiles words, zeroed
    int*           ticket;        // zeroed
    int*           sym_flag;
    ErrState*      err;
    int64_t*       n_records_out;
    NameTable      names;
    // fused K2a: query-side intervals (they contribute whatever the symmetric flag turns out to be, repeat.hpp:50-53)
    int32_t*       diff;      // null: do not scatter
    const int64_t* slot_off;
    int            reso;
    int64_t        own_first, own_count;
    ErrState*      err_range; // separate from `err`: an unknown name outranks a range error, as in the oracle
};

int         paf_tokenize_tiles(int64_t nbytes);
cudaError_t launch_paf_tokenize(const PafTokArgs& a, cudaStream_t st);
// first record of the text into rec0[0..6] (rec0[6] = found); rec0[7] = tail_flag (the caller's "this was my whole text")
void        launch_paf_peek(const uint8_t* text, int64_t nbytes, const NameTable& nt, int* rec0, ErrState* err, cudaStream_t st, int tail_flag = 0);

// ---------------------------------------------------------------- layout + K2 (k2_coverage.cu)
// per owned read: slots = nb+1, repeat capacity, cut capacity  (int32 each)
void launch_read_layout(const int64_t* seq_off_local, int64_t m, int reso, int p, int P, int l, int32_t* slots, int32_t* rep_cap,
                        int32_t* cut_cap, cudaStream_t st);
// exclusive scan int32[n] -> int64[n+1]; tmp_status must hold scan_tiles(n) uint64 (+1 int ticket after it), zeroed by the launcher
int  scan_tiles_small(int64_t n);
void launch_scan_i32_to_i64(const int32_t* in, int64_t* out, int64_t n, uint64_t* status, int* ticket, cudaStream_t st);
// inclusive scan of the int32 difference array into the coverage array, fused with the sizing of coverage.txt (bytes per 1024-slot tile)
constexpr int COV_TILE_SLOTS = 1024;
struct CovSizeArgs {
    int32_t*       tile_bytes;  // null: scan only
    const int32_t* tile_static; // coverage-independent bytes of every tile (launch_cov_static_sizes); tile_bytes starts from it
    const int64_t* slot_off;    // m+1
    const int32_t* tile_read;   // read holding the first slot of every 1024-slot tile
};
// Coverage-independent part of coverage.txt per 1024-slot tile: "read i " prefixes, the digits of every bin position
// (closed form per read and tile) and the sentinel newline, minus the 3 bytes the scan counts for a sentinel (whose
// coverage is always 0).  The scan then only adds digits(cov) + 2 per slot.
void launch_cov_static_sizes(const int64_t* slot_off, int64_t m, int64_t n_slots, int64_t own_first, int reso, int32_t* tile_static, cudaStream_t st);
int  scan_tiles_cov(int64_t n);
void launch_scan_cov(const int32_t* diff, int32_t* cov, int64_t n, uint64_t* status, int* ticket, const CovSizeArgs& cs, cudaStream_t st);

struct ScatterArgs {
    const int32_t *qid, *tid, *qs, *qe, *ts, *te;
    int64_t        n_rec;
    const int64_t* slot_off; // owned reads, m+1
    int32_t*       diff;
    int            reso;
    int64_t        own_first, own_count;
    const int*     sym_flag;
    ErrState*      err;
    int            skip_query; // query sides were already added by the tokenizer
};
void launch_scatter_records(const ScatterArgs& a, cudaStream_t st);
// routed endpoints: int32 triples (global read id, start, end)
void launch_scatter_endpoints(const int32_t* ep, int64_t n, const int64_t* slot_off, int32_t* diff, int reso, int64_t own_first,
                              int64_t own_count, ErrState* err, cudaStream_t st);
// multi-GPU routing
// count endpoints per destination rank and collect them (id, start, end, destination) into a bounded list; *list_n > cap
// means the list is incomplete and launch_route_pack (a second pass over the records) must be used
void launch_route_collect(const ScatterArgs& a, int nranks, const int64_t* bounds_dev, unsigned long long* counts_dev, int4* list,
                          unsigned long long* list_n, unsigned long long cap, cudaStream_t st);
void launch_route_pack_list(const int4* list, int64_t n, unsigned long long* cursors_dev, int32_t* sendbuf, cudaStream_t st);
void launch_route_pack(const ScatterArgs& a, int nranks, const int64_t* bounds_dev, unsigned long long* cursors_dev, int32_t* sendbuf,
                       cudaStream_t st);
// sharded runs inside the library (NCCL): tiny glue kernels around the collectives, so that no host round trip sits between them
// gathered = nranks x 8 ints (qid,tid,qs,qe,ts,te,found,whole_text_peeked): rec0 <- the first rank that found a record, rec0[7] = it is `rank`
void launch_pick_rec0(const int* gathered, int nranks, int rank, int* rec0, cudaStream_t st);
// out[0..nranks) = counts, out[nranks] = list_n, out[nranks+1] = *sym_flag, out[nranks+2..nranks+6) = extra (host scalars)
void launch_pack_counts(const unsigned long long* counts, const unsigned long long* list_n, const int* sym_flag, const long long extra[4],
                        int nranks, long long* out, cudaStream_t st);
// out[0..n) = *src[k] (scalars scattered over device memory), then the two error words
void launch_pack_scalars(const long long* const* src_host, int n, const ErrState* err, const ErrState* err_range, long long* out, cudaStream_t st);

// ---------------------------------------------------------------- K3 (k3_repeat_cut.cu)
struct RepeatCutArgs {
    const int32_t* cov;      // scanned slots
    const int64_t* slot_off; // m+1
    const int64_t* seq_off;  // m+1 (lengths)
    int64_t        m;
    int            reso, H, p, P, f, l;
    const int64_t* rep_cap_off; // m+1
    const int64_t* cut_cap_off; // m+1
    int2*          rep;         // capacity layout
    int32_t*       rep_cnt;     // m
    int32_t*       cuts;        // capacity layout
    int32_t*       frag_cnt;    // m
    unsigned long long* stats;  // [0]=sum cov, [1]=sum raw repeat len
    int*           work_counter; // zeroed
};
void launch_repeat_cut(const RepeatCutArgs& a, cudaStream_t st);

// Simulated-read names "read=N,forward|reverse,position=a-b,length=L,chr" (chop.hpp:14-70): what the header /
// BED variants need, parsed once per owned read.
struct __align__(16) SimInfo {
    int start_pos, end_pos;      // chop.hpp:25-47
    int align_off, align_len;    // chop.hpp:49-59, name relative
    int tail_off;                // offset of the last ',' (read_name.substr(find_last_of(',')), chop.hpp:257)
    int flags;                   // 1 forward, 2 reverse
    int pad[2];
};
// the three numbers of a simulated-read header: position=x-y,length=ln
__device__ __forceinline__ void sim_header_numbers(const SimInfo& si, bool whole, int fa, int fb, int L, int* x, int* y, int* ln)
{
    if (whole) { *x = si.start_pos; *y = si.end_pos; *ln = L; }
    else if (si.flags & 1) { *x = si.start_pos + fa; *y = si.start_pos + fb; *ln = fb - fa; }
    else { *x = si.end_pos - fb; *y = si.end_pos - fa; *ln = fb - fa; }
}

void launch_sim_parse(const uint8_t* names, const int64_t* name_off, int64_t own_first, int64_t m, SimInfo* out, ErrState* err, cudaStream_t st);

struct FragExpandArgs {
    int64_t        m;
    const int64_t* seq_off;     // local, m+1
    const int64_t* name_off;    // global name offsets, indexed by global id
    int64_t        own_first;
    const int64_t* cut_cap_off;
    const int32_t* cuts;
    const int32_t* frag_cnt;
    const int64_t* frag_base;   // m+1 (exclusive scan of frag_cnt)
    int            v;
    int64_t        read_num_base; // read= number of this context's first fragment minus 1
    int32_t *      frag_read, *frag_a, *frag_b; // G
    int32_t*       frag_size;   // G: bytes of the FASTA record
    ErrState*      err;
    const SimInfo* sim;         // null for real reads
};
void launch_frag_expand(const FragExpandArgs& a, cudaStream_t st);
// split_naive (split_naive.cpp:27-33): piece counts per read, then (read, a, b, record bytes) per piece
void launch_split_counts(const int64_t* seq_off, int64_t m, int sublen, int32_t* cnt, cudaStream_t st);
void launch_split_expand(const int64_t* seq_off, const int64_t* name_off, int64_t own_first, int64_t m, int sublen, const int64_t* base,
                         int32_t* frag_read, int32_t* frag_a, int32_t* frag_b, int32_t* frag_size, cudaStream_t st);
// compact repeats: rep_off = exclusive scan of rep_cnt; rep_out[2*k] pairs in read order; also text size per read line
void launch_rep_sizes(const int32_t* rep_cnt, const int64_t* rep_cap_off, const int2* rep, int64_t m, int64_t own_first, int32_t* line_size,
                      cudaStream_t st);
void launch_rep_compact(const int32_t* rep_cnt, const int64_t* rep_cap_off, const int64_t* rep_off, const int2* rep, int64_t m, int32_t* out,
                        cudaStream_t st);

// ---------------------------------------------------------------- K5 (k5_emit.cu)
struct CovEmitArgs {
    const int32_t* cov;      // scanned slots
    const int64_t* slot_off; // m+2: [m] = n_slots, [m+1] = 2^62 (lets the emitter step past the last read)
    int64_t        m, n_slots;
    int64_t        own_first; // global id of local read 0
    int            reso;
    const int64_t* tile_off;  // n_tiles+1 (bytes)
    uint8_t*       dst;       // window buffer: byte (w0 + k) of the stream goes to dst[k]
    int64_t        w0, w1;
    int64_t        tile_first; // first tile of this launch
    const int32_t* tile_read;  // n_tiles+1: read containing the tile's first slot (last entry = m-1 sentinel)
    int            text_cap;   // 32-slot chunks with more text than this take the direct (byte-wise) path
    const unsigned long long* pos_tab; // text of "<k*reso>," per bin index k < tab_n (launch_cov_tables)
    int            tab_n;
    const unsigned* cov_tab;   // text of "<c> " for c < 1000 (COV_TAB_ENTRIES entries)
};
constexpr int COV_POS_TAB_ENTRIES = 65536, COV_TAB_ENTRIES = 1024;
void launch_cov_tables(unsigned long long* pos_tab, int n, int reso, unsigned* cov_tab, cudaStream_t st);
// tile_read[T] = read whose slots contain slot T*COV_TILE_SLOTS
void launch_cov_tile_index(const int64_t* slot_off, int64_t m, int64_t n_slots, int32_t* tile_read, cudaStream_t st);
int  cov_tiles(int64_t n_slots);
void launch_cov_emit(const CovEmitArgs& a, int64_t n_tiles_launch, cudaStream_t st);

// long_repeats.bed (simulated reads only, repeat.hpp:187-199): chr \t x \t y \n per repeat
void launch_bed_sizes(const int32_t* rep_cnt, const int64_t* rep_cap_off, const int2* rep, const SimInfo* sim, const int64_t* name_off,
                      int64_t own_first, int64_t m, int32_t* line_size, cudaStream_t st);
struct RepEmitArgs {
    const int32_t* rep_cnt;
    const int64_t* rep_cap_off;
    const int2*    rep;
    const int64_t* line_off; // m+1
    int64_t        m, own_first;
    uint8_t*       dst;
    int64_t        w0, w1;
    int64_t        read_first, read_last; // local read range to emit
    const SimInfo* sim;      // BED variant when non-null (with names / name_off)
    const uint8_t* names;
    const int64_t* name_off;
};
void launch_bed_emit(const RepEmitArgs& a, cudaStream_t st);
void launch_rep_emit(const RepEmitArgs& a, cudaStream_t st);

constexpr int FASTA_TILE = 16384;
// everything the gather needs about one FASTA record, one 32-byte load
struct __align__(32) FragDesc {
    long long out_off; // stream offset of the record's '>'
    long long src_off; // arena offset of its first base (seq_off[read] + a)
    int       len;     // bases (b - a)
    int       hdr_len; // header bytes including its newline
    int       read;    // local read index
    int       a;       // first base on the original read
};
void launch_frag_desc(const int32_t* frag_read, const int32_t* frag_a, const int32_t* frag_b, const int32_t* frag_size, const int64_t* frag_off,
                      const int64_t* seq_off, int64_t G, FragDesc* desc, cudaStream_t st);
// every `step`-th record's (out_off, src_off): a coarse host-side map from output offset to arena offset
void launch_frag_sample(const FragDesc* desc, int64_t G, int step, int64_t* out2, cudaStream_t st);
struct FastaEmitArgs {
    const FragDesc* desc;    // G+1 (last: out_off = stream length)
    int64_t        G;
    const uint8_t* seq;      // local arena
    const int64_t* seq_off;  // local m+1
    const uint8_t* names;
    const int64_t* name_off; // global
    int64_t        own_first, read_num_base;
    uint8_t*       dst;
    int64_t        w0, w1;   // stream window; CTA b covers stream bytes [(w0/TILE + b)*TILE, +TILE) clipped to it
    const int32_t* tile_frag; // record containing stream byte T*FASTA_TILE, for every tile of the stream
    int64_t        seq_safe_end; // bytes of the arena that may be read in 16-byte blocks (multiple of 16)
    const SimInfo* sim;          // null for real reads
    int            split_len;    // > 0: split_naive records  ">" name "_" k "\n" bases "\n"  with k = a / split_len + 1
};
void launch_fasta_tile_index(const int64_t* frag_off, int64_t G, int32_t* tile_frag, cudaStream_t st);
void launch_fasta_emit(const FastaEmitArgs& a, cudaStream_t st);

// dst[k] = src[min(k*step, n-1)] for k in [0, (n-1)/step + 1]  (coarse host-side copies of offset tables)
void launch_sample_i64(const int64_t* src, int64_t n, int step, int64_t* dst, cudaStream_t st);
void launch_digest(const uint8_t* buf, int64_t n, int64_t abs_off, unsigned long long* acc, cudaStream_t st);

} // namespace raftk
