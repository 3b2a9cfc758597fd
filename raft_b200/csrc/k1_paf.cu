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
// names are hashed four bytes per step (funnel-shifted unaligned words), numbers eight digits at a time (SWAR).
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
#include "coverage.cuh"
#include "nametable.cuh"

namespace raftk {

constexpr int K1_THREADS = 160;             // 5 warps: one decode round covers the ~135 records of a typical tile
constexpr int K1_WPT = 4;                   // mask words per thread when enumerating line starts (160 * 4 >= 513)
constexpr int K1_LCAP = 1024;               // line list capacity of the dense validity pass (more lines: in-thread pass)
constexpr int K1_TILE = 16384;             // bytes of text whose newlines a CTA owns
constexpr int K1_OVER = 1024;              // staged overhang: lines starting near the tile end
constexpr int K1_STAGE = K1_TILE + K1_OVER;
constexpr int K1_WORDS = K1_STAGE / 32;    // 544 mask words over the staged bytes
constexpr int K1_TWORDS = K1_TILE / 32;    // 512: line starts live in bits [0, TILE] -> words [0, 512]

constexpr int K1_MAXREC = K1_TILE / 10 + 2; // a record has >= 9 tabs + newline: at most 1639 start in a tile

struct __align__(16) K1Smem {
    uint8_t  pad_front[16];       // lets the 8-byte number window start before the tile
    uint8_t  text[K1_STAGE + 16];
    unsigned nl[K1_WORDS + 1];    // byte is '\n'   (+1: all-ones sentinel word)
    unsigned tab[K1_WORDS + 1];   // byte is '\t'   (+1: all-ones sentinel word)
    uint16_t vpos[K1_MAXREC];     // start of the j-th record of the tile, tile relative
    uint16_t lpos[K1_LCAP];       // start of the j-th non-empty line (bit 15: it is a record)
    int      n_invalid;
    uint64_t bar;
    uint64_t bcast;
    int      scan_ws[34];
    int      tile;
};

// bit k set iff byte k of the 16-byte group equals c (c4 = c in every byte).  Exact per-byte zero test on
// v ^ c4 (0x80 per equal byte), then two unsigned dp4a per 8 bytes gather the flags: sum(128 * 2^i).
__device__ __forceinline__ unsigned zero_flags(unsigned x) { return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u; }
__device__ __forceinline__ unsigned eq_mask16(const uint4& v, unsigned c4)
{
    unsigned lo = __dp4a(zero_flags(v.x ^ c4), 0x08040201u, 0u);
    lo = __dp4a(zero_flags(v.y ^ c4), 0x80402010u, lo);
    unsigned hi = __dp4a(zero_flags(v.z ^ c4), 0x08040201u, 0u);
    hi = __dp4a(zero_flags(v.w ^ c4), 0x80402010u, hi);
    return (lo >> 7) | ((hi >> 7) << 8);
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

__device__ __forceinline__ void report_error(ErrState* err, int code, long long index) { err_min(err, code, index); }

// first set bit of mask m at position >= from; the sentinel word makes it return >= K1_STAGE when there is none
__device__ __forceinline__ int next_bit(const unsigned* m, int from)
{
    int      w = from >> 5;
    unsigned x = m[w] & (0xFFFFFFFFu << (from & 31));
    while (!x) x = m[++w];
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
// four ASCII digits (first digit in the lowest byte) -> value
__device__ __forceinline__ unsigned parse4(unsigned w)
{
    unsigned d = w & 0x0F0F0F0Fu;
    d = d * 10u + (d >> 8);                 // byte0 = 10*d0 + d1, byte2 = 10*d2 + d3
    return (d & 0xFFu) * 100u + ((d >> 16) & 0xFFu);
}
// Field [a, b) of 1..8 plain digits: take the 8 bytes ending at b, turn the bytes before the field into '0',
// check that all are digits and convert both halves at once.  Anything else (sign, spaces, junk, 9+ digits)
// goes to the strtol-exact byte loop.  `tw` = 32-bit view of the staged text INCLUDING the 16-byte front pad.
__device__ __forceinline__ int parse_num_smem(const uint8_t* text, const unsigned* tw, int a, int b)
{
    const int n = b - a;
    if (n < 1 || n > 8) return parse_num_generic(text, a, b);
    const int      k = b - 8 + 16; // >= 8 thanks to the front pad
    const int      idx = k >> 2;
    const unsigned sh = (unsigned)(k & 3) * 8u;
    const unsigned x0 = tw[idx], x1 = tw[idx + 1], x2 = tw[idx + 2];
    unsigned       w0 = __funnelshift_r(x0, x1, sh), w1 = __funnelshift_r(x1, x2, sh);
    const int      lead = 8 - n;
    if (lead >= 4) {
        const unsigned m = 0xFFFFFFFFu << (8 * (lead - 4));
        w0 = 0x30303030u; w1 = (w1 & m) | (0x30303030u & ~m);
    } else {
        const unsigned m = 0xFFFFFFFFu << (8 * lead);
        w0 = (w0 & m) | (0x30303030u & ~m);
    }
    const unsigned bad = ((w0 & 0xF0F0F0F0u) ^ 0x30303030u) | (((w0 + 0x06060606u) & 0xF0F0F0F0u) ^ 0x30303030u) |
                         ((w1 & 0xF0F0F0F0u) ^ 0x30303030u) | (((w1 + 0x06060606u) & 0xF0F0F0F0u) ^ 0x30303030u);
    if (bad) return parse_num_generic(text, a, b);
    return (int)(parse4(w0) * 10000u + parse4(w1));
}

// hash of staged bytes [a, b): same value as NameHasher fed byte by byte
__device__ __forceinline__ unsigned long long hash_name_smem(const unsigned* sw, int a, int b, unsigned long long seed)
{
    NameHasher hs;
    hs.init(seed);
    const int      n = b - a;
    int            wi = a >> 2;
    const unsigned sh = (unsigned)(a & 3) * 8u;
    unsigned       lo = sw[wi];
    int            k = 0;
    for (; k + 4 <= n; k += 4) {
        unsigned hi = sw[++wi];
        hs.word(__funnelshift_r(lo, hi, sh));
        lo = hi;
    }
    const int rem = n - k;
    if (rem) {
        unsigned hi = sw[wi + 1];
        hs.word(__funnelshift_r(lo, hi, sh) & ((1u << (8 * rem)) - 1u));
    }
    return name_hash_finish(hs.a, hs.b, (unsigned)n);
}

__global__ void __launch_bounds__(K1_THREADS, 8) k_paf_tokenize(PafTokArgs a)
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
    for (int w = tid; w < K1_WORDS; w += K1_THREADS) {
        const uint4 v0 = t16[2 * w], v1 = t16[2 * w + 1];
        s.nl[w] = eq_mask16(v0, 0x0A0A0A0Au) | (eq_mask16(v1, 0x0A0A0A0Au) << 16);
        s.tab[w] = eq_mask16(v0, 0x09090909u) | (eq_mask16(v1, 0x09090909u) << 16);
    }
    if (tid == 0) { s.nl[K1_WORDS] = 0xFFFFFFFFu; s.tab[K1_WORDS] = 0xFFFFFFFFu; s.n_invalid = 0; }
    __syncthreads();

    TextReader rd;
    rd.text = a.text; rd.nbytes = a.nbytes; rd.stage_lo = 0; rd.stage_len = 0; rd.swords = nullptr; rd.word = 0;

    // a line starting at tile-relative byte p is a record iff it has >= 9 tabs before its newline (paf.hpp:84)
    auto is_record = [&](int p) -> bool {
        int e = next_bit(s.nl, p);                 // the line's newline (>= K1_STAGE when it is not staged)
        if (e > K1_STAGE) e = K1_STAGE;
        int tabs = count_bits(s.tab, p, e);
        if (tabs < 9 && e == K1_STAGE && t0 + K1_STAGE < a.nbytes) { // runs past the staged bytes: read on
            rd.seek(t0 + K1_STAGE);
            for (;;) { int c = rd.next(); if (c < 0 || c == '\n') break; if (c == '\t' && ++tabs >= 9) break; }
        }
        return tabs >= 9;
    };

    // ---- phase A1: line starts.  Thread t owns mask words [4t, 4t+4).  Byte p starts a non-empty line owned by this
    // tile iff byte p-1 is '\n' inside the tile (or p is the file's first byte) and byte p is not '\n'.
    unsigned lsm[K1_WPT];
    int      my_lines = 0;
#pragma unroll
    for (int i = 0; i < K1_WPT; i++) {
        const int w = tid * K1_WPT + i;
        unsigned  m = 0;
        if (w <= K1_TWORDS) {
            unsigned prev = w ? (s.nl[w - 1] >> 31) : ((tile == 0) ? 1u : 0u);
            m = ((s.nl[w] << 1) | prev) & ~s.nl[w];
            if (w == K1_TWORDS) m &= 1u;
        }
        lsm[i] = m;
        my_lines += __popc(m);
    }
    int n_lines, n_valid;
    int lex = block_exclusive_sum<int, K1_THREADS>(my_lines, s.scan_ws, &n_lines);
    __syncthreads();
    if (n_lines <= K1_LCAP) {
        // ---- A2 (dense): list the line starts, then one thread per line checks validity
#pragma unroll
        for (int i = 0; i < K1_WPT; i++) {
            unsigned m = lsm[i];
            while (m) { int bit = __ffs(m) - 1; m &= m - 1; s.lpos[lex++] = (uint16_t)(((tid * K1_WPT + i) << 5) + bit); }
        }
        __syncthreads();
        int bad = 0;
        for (int j = tid; j < n_lines; j += K1_THREADS) {
            const int p = s.lpos[j];
            if (is_record(p)) s.lpos[j] = (uint16_t)(p | 0x8000); else bad++;
        }
        if (bad) atomicAdd(&s.n_invalid, bad);
        __syncthreads();
        if (s.n_invalid == 0) {
            // every line is a record (the usual case): the record list is the line list
            for (int j = tid; j < n_lines; j += K1_THREADS) s.vpos[j] = (uint16_t)(s.lpos[j] & 0x7FFF);
            n_valid = n_lines;
        } else {
            // order-preserving compaction: thread t owns list entries [t*C, (t+1)*C)
            const int C = (n_lines + K1_THREADS - 1) / K1_THREADS; // <= 7
            int       mine = 0;
            for (int k = 0; k < C; k++) { int j = tid * C + k; if (j < n_lines && (s.lpos[j] & 0x8000)) mine++; }
            int vex = block_exclusive_sum<int, K1_THREADS>(mine, s.scan_ws, &n_valid);
            for (int k = 0; k < C; k++) { int j = tid * C + k; if (j < n_lines && (s.lpos[j] & 0x8000)) s.vpos[vex++] = (uint16_t)(s.lpos[j] & 0x7FFF); }
        }
    } else {
        // ---- A2 (sparse, pathological tiles with > 1024 lines): every thread checks the lines of its own words
        unsigned vm[K1_WPT];
        int      mine = 0;
#pragma unroll
        for (int i = 0; i < K1_WPT; i++) {
            unsigned m = lsm[i], v = 0;
            while (m) { int bit = __ffs(m) - 1; m &= m - 1; if (is_record(((tid * K1_WPT + i) << 5) + bit)) v |= 1u << bit; }
            vm[i] = v; mine += __popc(v);
        }
        int vex = block_exclusive_sum<int, K1_THREADS>(mine, s.scan_ws, &n_valid);
#pragma unroll
        for (int i = 0; i < K1_WPT; i++) {
            unsigned v = vm[i];
            while (v) { int bit = __ffs(v) - 1; v &= v - 1; s.vpos[vex++] = (uint16_t)(((tid * K1_WPT + i) << 5) + bit); }
        }
    }
    if (tid == 0) lookback_publish(a.status, tile, (uint64_t)n_valid); // successors can look back while this tile decodes
    __syncthreads();

    // ---- phase B: decode the j-th record of the tile, one per thread
    int r0[8];
#pragma unroll
    for (int k = 0; k < 8; k++) r0[k] = a.rec0[k];
    const bool first_is_local = a.first_is_local >= 0 ? a.first_is_local != 0 : r0[7] != 0; // sharded runs decide it on the device
    const unsigned* sw = reinterpret_cast<const unsigned*>(s.text);
    const unsigned* tw = reinterpret_cast<const unsigned*>(s.pad_front);
    uint64_t        prefix = 0;
    bool            have_prefix = false;
    for (int base = 0; base < n_valid || !have_prefix; base += K1_THREADS) {
        const int  j = base + tid;
        const bool active = j < n_valid;
        ParsedRec  pr;
        pr.qid = pr.tid = pr.qs = pr.qe = pr.ts = pr.te = 0; pr.strand = 0;
        int p = 0;
        if (active) {
            p = s.vpos[j];
            // the nine tabs that end fields 0..8: walk the tab mask word by word
            int      tp[9];
            int      w = p >> 5;
            unsigned cur = s.tab[w] & (0xFFFFFFFFu << (p & 31));
#pragma unroll
            for (int k = 0; k < 9; k++) {
                while (!cur) cur = s.tab[++w];
                tp[k] = (w << 5) + __ffs(cur) - 1;
                cur &= cur - 1;
            }
            if (tp[8] < K1_STAGE) {
                pr.qid = nametable_find(a.names, hash_name_smem(sw, p, tp[0], a.names.seed));
                pr.tid = nametable_find(a.names, hash_name_smem(sw, tp[4] + 1, tp[5], a.names.seed));
                pr.qs = parse_num_smem(s.text, tw, tp[1] + 1, tp[2]); pr.qe = parse_num_smem(s.text, tw, tp[2] + 1, tp[3]);
                pr.strand = (tp[4] > tp[3] + 1) && s.text[tp[3] + 1] == '-';
                pr.ts = parse_num_smem(s.text, tw, tp[6] + 1, tp[7]); pr.te = parse_num_smem(s.text, tw, tp[7] + 1, tp[8]);
            } else { // the record runs past the staged bytes: byte-wise reader over global memory
                rd.seek(t0 + p);
                parse_record(rd, a.names, pr);
            }
        }
        if (!have_prefix) { prefix = lookback_wait(a.status, tile, (uint64_t)n_valid, &s.bcast); have_prefix = true; }
        if (!active) continue;
        const int64_t rec = a.rec_base + (int64_t)prefix + j;
        if (pr.qid < 0 || pr.tid < 0) { report_error(a.err, RAFTK_E_UNKNOWN_NAME, rec); continue; }
        if (rec < a.rec_cap) {
            a.qid[rec] = pr.qid; a.tid[rec] = pr.tid; a.qs[rec] = pr.qs; a.qe[rec] = pr.qe;
            a.ts[rec] = pr.ts; a.te[rec] = pr.te; a.strand[rec] = (uint8_t)pr.strand;
        }
        if (a.diff) { // fused K2a: the query side of every record covers [qs, qe) of its read (repeat.hpp:50-53)
            const int64_t lq = (int64_t)pr.qid - a.own_first;
            if (lq >= 0 && lq < a.own_count && !add_interval(a.diff, a.slot_off, lq, pr.qs, pr.qe, a.reso)) err_min(a.err_range, RAFTK_E_RANGE, rec);
        }
        // chop.hpp:171-184: record k >= 1 mirrors record 0
        if (r0[6] && (rec != 0 || !first_is_local) && r0[0] == pr.tid && r0[1] == pr.qid && r0[2] == pr.ts &&
            r0[3] == pr.te && r0[4] == pr.qs && r0[5] == pr.qe)
            *a.sym_flag = 1;
    }
    if (tid == 0 && tile == a.n_tiles - 1) *a.n_records_out = a.rec_base + (int64_t)prefix + n_valid;
}

// First record of the text (single thread; the first line is a record in any sane PAF).
__global__ void k_paf_peek(const uint8_t* text, int64_t nbytes, NameTable nt, int* rec0, ErrState* err, int tail_flag)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    rec0[7] = tail_flag;
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

void launch_paf_peek(const uint8_t* text, int64_t nbytes, const NameTable& nt, int* rec0, ErrState* err, cudaStream_t st, int tail_flag)
{
    k_paf_peek<<<1, 32, 0, st>>>(text, nbytes, nt, rec0, err, tail_flag);
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
