// k1_paf.cu — K1: PAF tokenizer + id decode + symmetric-overlap predicate.
//
// Replaces paf_read/paf_parse (paf.hpp:50-99, line reader kseq.h:107-193) and the per-record part
// of create_pileup (chop.hpp:147-187): the whole text sits in HBM, each CTA stages one 16 KiB tile
// (+ overhang) into shared memory with a 1-D TMA bulk copy, finds newline and tab positions with
// SWAR byte compares + popc + warp-shuffle scans, decides record validity (>= 10 fields,
// paf.hpp:84) from cumulative tab counts, obtains its first record index with a decoupled
// look-back over tiles, then decodes one line per thread into SoA int32 columns.
//
// Semantics kept from the reference:
//  * a record is a line with >= 9 tabs; shorter / blank lines are skipped (paf.hpp:84-85,96-98);
//  * only fields 1,3,4,5,6,8,9 are used (chop.hpp:157-163); numeric fields follow strtol(base 10)
//    -> uint32_t -> int (paf.hpp:62-81, overlap.hpp:14-16): leading isspace, optional sign, digits,
//    stop at the first other byte, saturate at LONG_MAX/LONG_MIN before truncation;
//  * strand = first byte of field 5 is '-' (paf.hpp:68-69);
//  * a trailing '\r' (kseq.h:189-190) can only sit in the last field of a line, which is never
//    one of the used fields of a valid record, so it needs no handling;
//  * symmetric flag: some record k >= 1 mirrors record 0 (chop.hpp:171-184).
#include "kernels.h"
#include "nametable.cuh"

namespace raftk {

constexpr int K1_THREADS = 256;
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int K1_TILE = 16384;             // bytes of text whose newlines a CTA owns
constexpr int K1_OVER = 1024;              // staged overhang: lines starting near the tile end
constexpr int K1_STAGE = K1_TILE + K1_OVER;
constexpr int K1_MAXLINES = K1_TILE / 2 + 1; // non-empty lines that can start in a tile (+1: file start)
constexpr int K1_GROUPS_PER_WARP = K1_TILE / 16 / K1_WARPS; // 128 16-byte groups per warp
constexpr int K1_ROUNDS = K1_GROUPS_PER_WARP / 32;          // 4

struct __align__(16) K1Smem {
    uint8_t            text[K1_STAGE];
    uint16_t           ls[K1_MAXLINES + 7]; // line start, tile relative (1..TILE; 0 only for file start)
    uint16_t           lt[K1_MAXLINES + 7]; // cumulative tabs in [tile start, line start); later: record rank
    uint64_t           bar;
    uint64_t           bcast;
    int                scan_ws[34];
    int                wcnt_nl[K1_WARPS + 1];
    int                wcnt_tab[K1_WARPS + 1];
    int                tile;
    int                last_line_tabs;
    int                last_nl; // tile-relative position of the last newline byte in the tile, -1 if none
};

// bit k set iff byte k of the 16-byte group equals c
__device__ __forceinline__ unsigned eq_mask16(const uint4& v, unsigned c4)
{
    auto m4 = [&](unsigned w) -> unsigned {
        unsigned x = __vcmpeq4(w, c4);                 // 0xFF per equal byte
        x &= 0x01010101u;                              // bit 0 of each byte
        return ((x * 0x01020408u) >> 24) & 0xFu;       // gather to 4 bits (byte i -> bit i)
    };
    return m4(v.x) | (m4(v.y) << 4) | (m4(v.z) << 8) | (m4(v.w) << 12);
}

// Byte reader over the text: staged bytes come from shared memory (32-bit word cache), the rest
// from global memory.  -1 at end of text.
struct TextReader {
    const uint8_t* text;
    int64_t        nbytes;
    int64_t        pos;      // absolute
    int64_t        stage_lo; // absolute offset of staged byte 0
    int            stage_len;
    const unsigned* swords;
    unsigned       word;
    __device__ __forceinline__ void seek(int64_t p)
    {
        pos = p;
        int64_t rel = p - stage_lo;
        if (swords && rel >= 0 && rel < stage_len) word = swords[rel >> 2];
    }
    __device__ __forceinline__ int next()
    {
        if (pos >= nbytes) return -1;
        int64_t  rel = pos - stage_lo;
        unsigned c;
        if (swords && rel >= 0 && rel < stage_len) {
            if ((rel & 3) == 0) word = swords[rel >> 2];
            c = (word >> ((unsigned)(rel & 3) * 8u)) & 0xffu;
        } else {
            c = text[pos];
        }
        pos++;
        return (int)c;
    }
};

struct NumState {
    unsigned long long acc;
    int                st; // 0 leading space, 1 sign seen, 2 digits, 3 done
    bool               neg, sat;
    __device__ __forceinline__ void reset() { acc = 0; st = 0; neg = false; sat = false; }
    __device__ __forceinline__ void add(unsigned c)
    {
        if (st == 3) return;
        unsigned d = c - '0';
        if (d <= 9u) {
            if (acc > (0xFFFFFFFFFFFFFFFFull - d) / 10ull) sat = true; else acc = acc * 10ull + d;
            st = 2;
        } else if (st == 0 && (c == ' ' || (c >= 9u && c <= 13u))) {
        } else if (st == 0 && (c == '+' || c == '-')) {
            neg = (c == '-'); st = 1;
        } else {
            st = 3;
        }
    }
    __device__ __forceinline__ int value() const
    { // strtol saturation, then (uint32_t) then (int)
        long long v;
        if (neg) v = (sat || acc > 0x8000000000000000ull) ? (long long)0x8000000000000000ull : (long long)(0ull - acc);
        else     v = (sat || acc > 0x7FFFFFFFFFFFFFFFull) ? 0x7FFFFFFFFFFFFFFFll : (long long)acc;
        return (int)(unsigned)(unsigned long long)v;
    }
};

struct ParsedRec { int qid, tid, qs, qe, ts, te; unsigned strand; };

// Parses fields 0..8 of the line the reader is positioned at.  Returns false when the line ends
// before its 9th tab (not a record).
__device__ __forceinline__ bool parse_record(TextReader& rd, const NameTable& nt, ParsedRec& r)
{
    NameHasher hq, ht;
    hq.init(nt.seed); ht.init(nt.seed);
    NumState num; num.reset();
    int  field = 0;
    bool first = true;
    r.qs = r.qe = r.ts = r.te = 0; r.strand = 0;
    for (;;) {
        int c = rd.next();
        if (c < 0 || c == '\n') return false;
        if (c == '\t') {
            if (field == 2) r.qs = num.value(); else if (field == 3) r.qe = num.value();
            else if (field == 7) r.ts = num.value(); else if (field == 8) r.te = num.value();
            num.reset(); first = true;
            if (++field == 9) break;
            continue;
        }
        if (field == 0) hq.add((unsigned)c);
        else if (field == 5) ht.add((unsigned)c);
        else if (field == 4) { if (first) r.strand = (c == '-'); }
        else if (field == 2 || field == 3 || field == 7 || field == 8) num.add((unsigned)c);
        first = false;
    }
    r.qid = nametable_find(nt, hq.finish());
    r.tid = nametable_find(nt, ht.finish());
    return true;
}

__device__ __forceinline__ void report_error(ErrState* err, int code, long long index)
{
    // keep the smallest index; the code of that record wins (codes are set once per index race-free enough
    // for diagnostics: any racing writer carries a real error)
    long long old = atomicMin(&err->index, index);
    if (index <= old) err->code = code;
}

__global__ void __launch_bounds__(K1_THREADS) k_paf_tokenize(PafTokArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    K1Smem& s = *reinterpret_cast<K1Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();

    if (tid == 0) {
        s.tile = atomicAdd(a.ticket, 1);
        s.last_nl = -1;
        mbar_init(&s.bar, 1);
    }
    __syncthreads();
    const int     tile = s.tile;
    const int64_t t0 = (int64_t)tile * K1_TILE;
    const int64_t avail = a.nbytes - t0;                           // > 0
    const int     want = (int)(avail < K1_STAGE ? avail : K1_STAGE);
    const int     bulk = want & ~15;                               // 16-byte multiple moved by TMA

    // ---- stage: TMA bulk for the aligned part, threads for the <16-byte tail and the '\n' fill
    if (tid == 0 && bulk > 0) {
        mbar_expect_tx(&s.bar, (uint32_t)bulk);
        tma_load_1d(s.text, a.text + t0, (uint32_t)bulk, &s.bar);
    }
    for (int j = bulk + tid; j < K1_STAGE; j += K1_THREADS) s.text[j] = (j < want) ? a.text[t0 + j] : (uint8_t)'\n';
    if (bulk > 0) mbar_wait(&s.bar, 0);
    __syncthreads();

    // ---- pass 1: per-warp newline (line-start) and tab totals over its 2 KiB
    const uint4* t16 = reinterpret_cast<const uint4*>(s.text);
    unsigned     m_nl[K1_ROUNDS], m_tab[K1_ROUNDS];
    int          c_nl = 0, c_tab = 0, my_last_nl = -1;
#pragma unroll
    for (int r = 0; r < K1_ROUNDS; r++) {
        int      g = warp * K1_GROUPS_PER_WARP + r * 32 + lane;
        uint4    v = t16[g];
        unsigned nl = eq_mask16(v, 0x0A0A0A0Au);
        if (nl) my_last_nl = g * 16 + (31 - __clz(nl));
        unsigned nxt = (nl >> 1) | ((s.text[g * 16 + 16] == '\n') ? 0x8000u : 0u);
        m_nl[r] = nl & ~nxt;                 // newline followed by a non-newline byte: a non-empty line starts
        m_tab[r] = eq_mask16(v, 0x09090909u);
        c_nl += __popc(m_nl[r]); c_tab += __popc(m_tab[r]);
    }
    {
        int wn = warp_sum(c_nl), wt = warp_sum(c_tab);
        int wl = my_last_nl;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) wl = max(wl, __shfl_xor_sync(FULL, wl, d));
        if (lane == 0) { s.wcnt_nl[warp] = wn; s.wcnt_tab[warp] = wt; if (wl >= 0) atomicMax(&s.last_nl, wl); }
    }
    __syncthreads();
    if (tid == 0) {
        int an = (tile == 0) ? 1 : 0, at = 0; // the file's first byte starts a line (no preceding newline)
        for (int w = 0; w < K1_WARPS; w++) { int n = s.wcnt_nl[w], t = s.wcnt_tab[w]; s.wcnt_nl[w] = an; s.wcnt_tab[w] = at; an += n; at += t; }
        s.wcnt_nl[K1_WARPS] = an; s.wcnt_tab[K1_WARPS] = at;
        if (tile == 0) { s.ls[0] = 0; s.lt[0] = 0; }
    }
    __syncthreads();
    // ---- pass 2: write (line start, cumulative tabs) in text order
    {
        int bn = s.wcnt_nl[warp], bt = s.wcnt_tab[warp];
#pragma unroll
        for (int r = 0; r < K1_ROUNDS; r++) {
            int g = warp * K1_GROUPS_PER_WARP + r * 32 + lane;
            int cn = __popc(m_nl[r]), ct = __popc(m_tab[r]);
            int in_ = warp_inclusive_sum(cn), it = warp_inclusive_sum(ct);
            int on = bn + in_ - cn, ot = bt + it - ct;
            unsigned m = m_nl[r];
            while (m) {
                int k = __ffs(m) - 1; m &= m - 1;
                s.ls[on] = (uint16_t)(g * 16 + k + 1);
                s.lt[on] = (uint16_t)(ot + __popc(m_tab[r] & ((2u << k) - 1u)));
                on++;
            }
            bn += __shfl_sync(FULL, in_, 31); bt += __shfl_sync(FULL, it, 31);
        }
    }
    __syncthreads();
    const int nl = s.wcnt_nl[K1_WARPS];
    const int tabs_tile = s.wcnt_tab[K1_WARPS];

    TextReader rd;
    rd.text = a.text; rd.nbytes = a.nbytes; rd.stage_lo = t0; rd.stage_len = K1_STAGE;
    rd.swords = reinterpret_cast<const unsigned*>(s.text); rd.word = 0;

    // the last line may run past the tile: count its tabs up to the 9th by reading on
    if (tid == 0 && nl > 0) {
        int tabs = tabs_tile - s.lt[nl - 1];
        // closed inside the tile (a newline at or after its start)? then all its tabs are already counted:
        // only blank lines can follow the last recorded line start
        bool closed = s.last_nl >= (int)s.ls[nl - 1];
        if (!closed && tabs < 9) {
            rd.seek(t0 + K1_TILE);
            for (;;) { int c = rd.next(); if (c < 0 || c == '\n') break; if (c == '\t' && ++tabs >= 9) break; }
        }
        s.last_line_tabs = tabs;
    }
    __syncthreads();

    // ---- phase A: validity per line -> ranks (stored over lt) and the tile's record count
    int base = 0;
    for (int i0 = 0; i0 < nl; i0 += K1_THREADS) {
        int i = i0 + tid;
        int valid = 0;
        if (i < nl) {
            int tabs = (i == nl - 1) ? s.last_line_tabs : (int)s.lt[i + 1] - (int)s.lt[i];
            valid = tabs >= 9;
        }
        int tot;
        int ex = block_exclusive_sum<int, K1_THREADS>(valid, s.scan_ws, &tot);
        if (i < nl) s.lt[i] = valid ? (uint16_t)(base + ex) : (uint16_t)0xFFFF;
        base += tot;
        __syncthreads();
    }
    const int      n_valid = base;
    const uint64_t prefix = lookback_block(a.status, tile, (uint64_t)n_valid, &s.bcast);
    if (tid == 0 && tile == a.n_tiles - 1) *a.n_records_out = a.rec_base + (int64_t)prefix + n_valid;

    // ---- phase B: decode valid lines, one per thread
    int r0[7];
#pragma unroll
    for (int k = 0; k < 7; k++) r0[k] = a.rec0[k];
    for (int i = tid; i < nl; i += K1_THREADS) {
        unsigned rank = s.lt[i];
        if (rank == 0xFFFFu) continue;
        int64_t rec = a.rec_base + (int64_t)prefix + rank;
        rd.seek(t0 + s.ls[i]);
        ParsedRec pr;
        parse_record(rd, a.names, pr);
        if (pr.qid < 0 || pr.tid < 0) { report_error(a.err, RAFTK_E_UNKNOWN_NAME, rec); continue; }
        if (rec < a.rec_cap) {
            a.qid[rec] = pr.qid; a.tid[rec] = pr.tid; a.qs[rec] = pr.qs; a.qe[rec] = pr.qe;
            a.ts[rec] = pr.ts; a.te[rec] = pr.te; a.strand[rec] = (uint8_t)pr.strand;
        }
        // chop.hpp:171-184: record k >= 1 mirrors record 0
        if (r0[6] && (rec != 0 || !a.first_is_local) && r0[0] == pr.tid && r0[1] == pr.qid && r0[2] == pr.ts &&
            r0[3] == pr.te && r0[4] == pr.qs && r0[5] == pr.qe)
            *a.sym_flag = 1;
    }
}

// First record of the text (single thread; the first line is a record in any sane PAF).
__global__ void k_paf_peek(const uint8_t* text, int64_t nbytes, NameTable nt, int* rec0, ErrState* err)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    TextReader rd;
    rd.text = text; rd.nbytes = nbytes; rd.stage_lo = 0; rd.stage_len = 0; rd.swords = nullptr; rd.word = 0;
    int64_t ls = 0;
    rec0[6] = 0;
    while (ls < nbytes) {
        rd.seek(ls);
        ParsedRec pr;
        if (parse_record(rd, nt, pr)) {
            if (pr.qid < 0 || pr.tid < 0) { report_error(err, RAFTK_E_UNKNOWN_NAME, 0); return; }
            rec0[0] = pr.qid; rec0[1] = pr.tid; rec0[2] = pr.qs; rec0[3] = pr.qe; rec0[4] = pr.ts; rec0[5] = pr.te;
            rec0[6] = 1;
            return;
        }
        // parse_record stopped at the newline / EOF that ended the short line
        ls = rd.pos;
    }
}

void launch_paf_peek(const uint8_t* text, int64_t nbytes, const NameTable& nt, int* rec0, ErrState* err, cudaStream_t st)
{
    k_paf_peek<<<1, 32, 0, st>>>(text, nbytes, nt, rec0, err);
}

int paf_tokenize_tiles(int64_t nbytes) { return (int)((nbytes + K1_TILE - 1) / K1_TILE); }

cudaError_t launch_paf_tokenize(const PafTokArgs& a, cudaStream_t st)
{
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_paf_tokenize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1Smem));
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (a.n_tiles <= 0) return cudaSuccess;
    k_paf_tokenize<<<a.n_tiles, K1_THREADS, sizeof(K1Smem), st>>>(a);
    return cudaGetLastError();
}

} // namespace raftk
