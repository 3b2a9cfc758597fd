// k1_paf.cu — K1: PAF tokenizer + id decode + symmetric-overlap predicate.
//
// Replaces paf_read/paf_parse (paf.hpp:50-99, line reader kseq.h:107-193) and the per-record part
// of create_pileup (chop.hpp:147-187).  The whole text sits in HBM; each CTA stages one 16 KiB tile
// (+ 1 KiB overhang) into shared memory with a 1-D TMA bulk copy and turns it into three bit masks
// (newline, tab, start-of-non-empty-line) with SWAR byte compares.  Everything per line is then
// derived from the masks with popc / ffs: validity (>= 9 tabs before the line's newline,
// paf.hpp:84) by popcounts, the tile's record count by a block scan, its first record index by a
// decoupled look-back over tiles, the j-th valid line by select on the mask, and the nine field
// boundaries by ffs hops over the tab mask.  Fields are decoded straight from shared memory:
// names are hashed four bytes per step (funnel-shifted unaligned words), numbers by a digit loop.
// Lines that run past the staged bytes (e.g. kilobyte cg:Z: CIGAR tags) take a byte-wise path
// that reads global memory.
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
constexpr int K1_TILE = 16384;             // bytes of text whose newlines a CTA owns
constexpr int K1_OVER = 1024;              // staged overhang: lines starting near the tile end
constexpr int K1_STAGE = K1_TILE + K1_OVER;
constexpr int K1_WORDS = K1_STAGE / 32;    // 544 mask words over the staged bytes
constexpr int K1_TWORDS = K1_TILE / 32;    // 512: line starts live in bits [0, TILE] -> words [0, 512]

struct __align__(16) K1Smem {
    uint8_t  text[K1_STAGE + 16];
    unsigned nl[K1_WORDS];    // byte is '\n'
    unsigned tab[K1_WORDS];   // byte is '\t'
    unsigned ls[K1_WORDS];    // a non-empty line owned by this tile starts at this byte
    unsigned valid[K1_WORDS]; // ... and it is a record (>= 9 tabs)
    uint16_t vpre[K1_WORDS + 2]; // records before word w
    uint64_t bar;
    uint64_t bcast;
    int      scan_ws[34];
    int      tile;
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

// first set bit of mask m at position >= from (bit index over K1_WORDS words), or K1_STAGE
__device__ __forceinline__ int next_bit(const unsigned* m, int from)
{
    if (from >= K1_STAGE) return K1_STAGE;
    int      w = from >> 5;
    unsigned x = m[w] & (0xFFFFFFFFu << (from & 31));
    while (!x) { if (++w >= K1_WORDS) return K1_STAGE; x = m[w]; }
    return (w << 5) + __ffs(x) - 1;
}
// number of set bits of m in [from, to)
__device__ __forceinline__ int count_bits(const unsigned* m, int from, int to)
{
    if (to <= from) return 0;
    int      w0 = from >> 5, w1 = (to - 1) >> 5;
    unsigned first = 0xFFFFFFFFu << (from & 31), last = 0xFFFFFFFFu >> (31 - ((to - 1) & 31));
    if (w0 == w1) return __popc(m[w0] & first & last);
    int c = __popc(m[w0] & first) + __popc(m[w1] & last);
    for (int w = w0 + 1; w < w1; w++) c += __popc(m[w]);
    return c;
}

// strtol(base 10) -> uint32_t -> int over staged bytes [a, b)
__device__ __noinline__ int parse_num_generic(const uint8_t* s, int a, int b)
{
    NumState num;
    num.reset();
    for (int i = a; i < b; i++) { num.add(s[i]); if (num.st == 3) break; }
    return num.value();
}
// fast path: 1..9 plain digits filling the whole field
__device__ __forceinline__ int parse_num_smem(const uint8_t* s, int a, int b)
{
    unsigned acc = 0;
    int      i = a;
    for (; i < b; i++) { unsigned d = (unsigned)s[i] - '0'; if (d > 9u) break; acc = acc * 10u + d; }
    if (i == b && b > a && b - a <= 9) return (int)acc;
    return parse_num_generic(s, a, b);
}

// hash of staged bytes [a, b): same value as NameHasher fed byte by byte
__device__ __forceinline__ unsigned long long hash_name_smem(const unsigned* sw, int a, int b, unsigned long long seed)
{
    unsigned long long h = seed ^ 0x9E3779B97F4A7C15ull;
    const int          n = b - a;
    int                wi = a >> 2;
    const unsigned     sh = (unsigned)(a & 3) * 8u;
    unsigned           lo = sw[wi];
    int                k = 0;
    for (; k + 4 <= n; k += 4) {
        unsigned hi = sw[++wi];
        unsigned w = __funnelshift_r(lo, hi, sh);
        lo = hi;
        h = (h ^ w) * 0xD6E8FEB86659FD93ull; h ^= h >> 29;
    }
    const int rem = n - k;
    if (rem) {
        unsigned hi = sw[wi + 1];
        unsigned w = __funnelshift_r(lo, hi, sh) & ((1u << (8 * rem)) - 1u);
        h = (h ^ w) * 0xD6E8FEB86659FD93ull; h ^= h >> 29;
    }
    unsigned long long r = mix64(h ^ ((unsigned long long)(unsigned)n << 32));
    return r ? r : 1ull;
}

// position of the j-th set bit of a mask given its per-word exclusive prefix counts (pre has K1_TWORDS+2 entries)
__device__ __forceinline__ int select_bit(const unsigned* m, const uint16_t* pre, int j)
{
    int lo = 0, hi = K1_TWORDS + 1;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if ((int)pre[mid] <= j) lo = mid; else hi = mid; }
    unsigned v = m[lo];
    for (int k = j - (int)pre[lo]; k > 0; k--) v &= v - 1;
    return (lo << 5) + __ffs(v) - 1;
}

// exclusive prefix of popc over mask words [0, K1_TWORDS] -> pre[0..K1_TWORDS+1]; returns the total.  Block-wide.
__device__ __forceinline__ int mask_prefix(const unsigned* m, uint16_t* pre, int* scan_ws)
{
    const int tid = threadIdx.x;
    const int wlo = tid * 2, whi = (tid == K1_THREADS - 1) ? K1_TWORDS + 1 : wlo + 2;
    int       mine = 0;
    for (int w = wlo; w < whi; w++) mine += __popc(m[w]);
    int tot;
    int run = block_exclusive_sum<int, K1_THREADS>(mine, scan_ws, &tot);
    for (int w = wlo; w < whi; w++) { pre[w] = (uint16_t)run; run += __popc(m[w]); }
    if (tid == K1_THREADS - 1) pre[K1_TWORDS + 1] = (uint16_t)run;
    return tot;
}

__global__ void __launch_bounds__(K1_THREADS) k_paf_tokenize(PafTokArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    K1Smem& s = *reinterpret_cast<K1Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = lane_id();

    if (tid == 0) {
        s.tile = atomicAdd(a.ticket, 1);
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
    for (int j = bulk + tid; j < K1_STAGE + 16; j += K1_THREADS) s.text[j] = (j < want) ? a.text[t0 + j] : (uint8_t)'\n';
    if (bulk > 0) mbar_wait(&s.bar, 0);
    __syncthreads();

    // ---- newline / tab masks: one 16-byte group per lane, two lanes make one 32-bit word
    const uint4* t16 = reinterpret_cast<const uint4*>(s.text);
    for (int g = tid; g < K1_STAGE / 16; g += K1_THREADS) { // 1088 groups; the last round is two full warps
        uint4    v = t16[g];
        unsigned mn = eq_mask16(v, 0x0A0A0A0Au), mt = eq_mask16(v, 0x09090909u);
        unsigned pn = __shfl_down_sync(FULL, mn, 1), pt = __shfl_down_sync(FULL, mt, 1);
        if (!(lane & 1)) { s.nl[g >> 1] = mn | (pn << 16); s.tab[g >> 1] = mt | (pt << 16); }
    }
    __syncthreads();
    // ---- line starts: byte p starts a non-empty line iff byte p-1 is '\n' (owned: p-1 inside the tile) and byte p is not
    for (int w = tid; w < K1_WORDS; w += K1_THREADS) {
        unsigned m = 0;
        if (w <= K1_TWORDS) {
            unsigned prev = w ? (s.nl[w - 1] >> 31) : ((tile == 0) ? 1u : 0u); // the file's first byte starts a line
            m = ((s.nl[w] << 1) | prev) & ~s.nl[w];
            if (w == K1_TWORDS) m &= 1u;
        }
        s.ls[w] = m;
        s.valid[w] = 0;
    }
    __syncthreads();
    const int n_lines = mask_prefix(s.ls, s.vpre, s.scan_ws); // vpre: line-start prefix for now
    __syncthreads();

    TextReader rd;
    rd.text = a.text; rd.nbytes = a.nbytes; rd.stage_lo = 0; rd.stage_len = 0; rd.swords = nullptr; rd.word = 0;

    // ---- phase A: one thread per line: a record iff >= 9 tabs before its newline (paf.hpp:84)
    for (int j = tid; j < n_lines; j += K1_THREADS) {
        int p = select_bit(s.ls, s.vpre, j);
        int e = next_bit(s.nl, p);                 // the line's newline, or K1_STAGE when it is not staged
        int tabs = count_bits(s.tab, p, e);
        if (tabs < 9 && e == K1_STAGE && t0 + K1_STAGE < a.nbytes) { // runs past the staged bytes: read on
            rd.seek(t0 + K1_STAGE);
            for (;;) { int c = rd.next(); if (c < 0 || c == '\n') break; if (c == '\t' && ++tabs >= 9) break; }
        }
        if (tabs >= 9) atomicOr(&s.valid[p >> 5], 1u << (p & 31));
    }
    __syncthreads();
    const int n_valid = mask_prefix(s.valid, s.vpre, s.scan_ws); // vpre: record prefix from here on
    if (tid == 0) {
        lookback_publish(a.status, tile, (uint64_t)n_valid);     // successors can look back while this tile decodes
    }
    __syncthreads();

    // ---- phase B: decode.  Two threads per record when the tile's records fit (even lane: query side, odd lane: target side)
    int r0[7];
#pragma unroll
    for (int k = 0; k < 7; k++) r0[k] = a.rec0[k];
    const unsigned* sw = reinterpret_cast<const unsigned*>(s.text);
    const bool      pair = n_valid * 2 <= K1_THREADS;
    const int       per_round = pair ? K1_THREADS / 2 : K1_THREADS;
    uint64_t        prefix = 0;
    bool            have_prefix = false;
    for (int base = 0; base < n_valid || !have_prefix; base += per_round) {
        const int  j = base + (pair ? (tid >> 1) : tid);
        const bool active = j < n_valid;
        const bool doA = active && (!pair || !(tid & 1)), doB = active && (!pair || (tid & 1));
        int  p = 0, t4 = 0, qid = 0, tidv = 0, qs = 0, qe = 0, ts = 0, te = 0;
        unsigned strand = 0;
        bool staged = true;
        if (active) p = select_bit(s.valid, s.vpre, j);
        if (doA) {
            int t0_ = next_bit(s.tab, p), t1 = next_bit(s.tab, t0_ + 1), t2 = next_bit(s.tab, t1 + 1), t3 = next_bit(s.tab, t2 + 1);
            t4 = next_bit(s.tab, t3 + 1);
            if (t4 < K1_STAGE) {
                qid = nametable_find(a.names, hash_name_smem(sw, p, t0_, a.names.seed));
                qs = parse_num_smem(s.text, t1 + 1, t2); qe = parse_num_smem(s.text, t2 + 1, t3);
                strand = (t4 > t3 + 1) && s.text[t3 + 1] == '-';
            } else staged = false;
        }
        if (pair) t4 = __shfl_sync(FULL, t4, lane & ~1);
        if (doB) {
            if (t4 < K1_STAGE) {
                int t5 = next_bit(s.tab, t4 + 1), t6 = next_bit(s.tab, t5 + 1), t7 = next_bit(s.tab, t6 + 1), t8 = next_bit(s.tab, t7 + 1);
                if (t8 < K1_STAGE) {
                    tidv = nametable_find(a.names, hash_name_smem(sw, t4 + 1, t5, a.names.seed));
                    ts = parse_num_smem(s.text, t6 + 1, t7); te = parse_num_smem(s.text, t7 + 1, t8);
                } else staged = false;
            } else staged = false;
        }
        if (pair) staged = __shfl_sync(FULL, (int)staged, lane & ~1) && __shfl_sync(FULL, (int)staged, lane | 1);
        if (active && !staged) { // the record runs past the staged bytes: byte-wise reader over global memory (both lanes of a pair)
            ParsedRec pr;
            rd.seek(t0 + p);
            parse_record(rd, a.names, pr);
            qid = pr.qid; tidv = pr.tid; qs = pr.qs; qe = pr.qe; ts = pr.ts; te = pr.te; strand = pr.strand;
        }
        if (!have_prefix) { prefix = lookback_wait(a.status, tile, (uint64_t)n_valid, &s.bcast); have_prefix = true; }
        const int64_t rec = a.rec_base + (int64_t)prefix + j;
        // symmetric predicate (chop.hpp:171-184): record k >= 1 mirrors record 0; each lane checks its half
        bool mA = r0[1] == qid && r0[4] == qs && r0[5] == qe, mB = r0[0] == tidv && r0[2] == ts && r0[3] == te;
        if (pair) { bool oA = __shfl_sync(FULL, (int)mA, lane & ~1), oB = __shfl_sync(FULL, (int)mB, lane | 1); mA = oA; mB = oB; }
        if (doA) {
            if (qid < 0) report_error(a.err, RAFTK_E_UNKNOWN_NAME, rec);
            else if (rec < a.rec_cap) { a.qid[rec] = qid; a.qs[rec] = qs; a.qe[rec] = qe; a.strand[rec] = (uint8_t)strand; }
            if (r0[6] && (rec != 0 || !a.first_is_local) && mA && mB) *a.sym_flag = 1;
        }
        if (doB) {
            if (tidv < 0) report_error(a.err, RAFTK_E_UNKNOWN_NAME, rec);
            else if (rec < a.rec_cap) { a.tid[rec] = tidv; a.ts[rec] = ts; a.te[rec] = te; }
        }
    }
    if (tid == 0 && tile == a.n_tiles - 1) *a.n_records_out = a.rec_base + (int64_t)prefix + n_valid;
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
