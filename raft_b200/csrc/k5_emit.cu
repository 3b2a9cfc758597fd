// k5_emit.cu — K5: text and sequence emitters.  Each output file of the reference is a byte
// stream of known length that is materialised on the device one window [w0, w1) at a time:
//
//  K5a coverage.txt   "read " i " " then k*reso "," cov " " per bin, then "\n"   (repeat.hpp:105-108)
//  K5c long_repeats   "read " i ", " then s "," e "    " per repeat, then "\n"   (repeat.hpp:180-203)
//  K5b reads.fasta    ">read=" num "," name ",pos_on_original_read=" a "-" b "\n" bases[a:b] "\n"
//                                                                                (chop.hpp:261-265,314-318)
//
// K5a is slot-parallel (a slot is one bin, or the per-read sentinel that carries the newline):
// a tile of 1024 slots is formatted into shared memory at the same 16-byte phase as its
// destination and then stored with aligned 128-bit writes.  K5b is output-tile-parallel: every CTA
// owns 16 KiB of the output file, finds the fragments that intersect it by binary search over the
// record offsets, generates header bytes on the fly and gathers sequence bytes with aligned
// 128-bit loads + funnel shifts + aligned 128-bit stores (stream-compacted: bytes land in their
// final file order).
#include <cstdlib>

#include "covtext.cuh"

namespace raftk {

// ================================================================ K5a coverage.txt
constexpr int CE_THREADS = 256;
constexpr int CE_PER = COV_TILE_SLOTS / CE_THREADS; // 4 slots per thread
constexpr int CE_MAX_SLOT_BYTES = 40;               // "read 2147483647 " (16) + "2147483647,-2147483648 " (23)
constexpr int CE_CAP = 16384;                       // shared-memory text buffer; tiles with more text (tiny reads) use the direct path
constexpr int CE_SMEM = CE_CAP + 32;

// text of this thread's slots at p (shared or local memory after inlining); returns the end
__device__ __forceinline__ uint8_t* cov_format_slots(uint8_t* p, SlotWalk w, int nmine, const int* cv, const int* dg, int reso, int64_t own_first)
{
#pragma unroll
    for (int k = 0; k < CE_PER; k++) {
        if (k < nmine) {
            if (w.bin == 0) {
                p[0] = 'r'; p[1] = 'e'; p[2] = 'a'; p[3] = 'd'; p[4] = ' '; p += 5;
                uint64_t id = (uint64_t)(own_first + w.r);
                int      nd = dec_digits64(id);
                for (int d = nd - 1; d >= 0; d--) { p[d] = (uint8_t)('0' + (unsigned)(id % 10ull)); id /= 10ull; }
                p += nd; *p++ = ' ';
            }
            if (w.left == 1) {
                *p++ = '\n';
            } else {
                p = put_u32_nd(p, (uint32_t)w.bin * (uint32_t)reso, dg[k] & 15); *p++ = ',';
                if (dg[k] >> 8) *p++ = '-';
                p = put_u32_nd(p, cv[k] < 0 ? (unsigned)(-(int64_t)cv[k]) : (unsigned)cv[k], (dg[k] >> 4) & 15); *p++ = ' ';
            }
            w.next();
        }
    }
    return p;
}

template <bool EMIT>
__device__ __forceinline__ void cov_text_tile(const CovEmitArgs& a, const int64_t tile, uint8_t* sbuf, int* ws)
{
    const int64_t  g0 = tile * COV_TILE_SLOTS + (int64_t)threadIdx.x * CE_PER;
    const int      nmine = g0 >= a.n_slots ? 0 : (a.n_slots - g0 < CE_PER ? (int)(a.n_slots - g0) : CE_PER);
    int            mine = 0;
    int            cv[CE_PER], dg[CE_PER];
    SlotWalk       w0;
    if (nmine) {
        // the tile map bounds the search to the reads that intersect this tile
        int64_t lo = a.tile_read[tile], hi = (int64_t)a.tile_read[tile + 1] + 1;
        while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (a.slot_off[mid] <= g0) lo = mid; else hi = mid; }
        w0.init(a.slot_off, lo, g0);
    }
    if (nmine == CE_PER) { // 4 consecutive ints, 16-byte aligned
        int4 v = *reinterpret_cast<const int4*>(a.cov + g0);
        cv[0] = v.x; cv[1] = v.y; cv[2] = v.z; cv[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < CE_PER; k++) cv[k] = (k < nmine) ? a.cov[g0 + k] : 0;
    }
    {
        SlotWalk w = w0;
#pragma unroll
        for (int k = 0; k < CE_PER; k++) {
            dg[k] = 0;
            if (k < nmine) {
                if (w.bin == 0) mine += 5 + dec_digits64((uint64_t)(a.own_first + w.r)) + 1;
                if (w.left == 1) mine += 1;
                else { dg[k] = slot_digits(w.bin, a.reso, cv[k]); mine += (dg[k] & 15) + ((dg[k] >> 4) & 15) + (dg[k] >> 8) + 2; }
                w.next();
            }
        }
    }
    int tot;
    int ex = block_exclusive_sum<int, CE_THREADS>(mine, ws, &tot);
    if (!EMIT) {
        if (threadIdx.x == 0) a.tile_bytes[tile] = tot;
        return;
    }
    const int64_t o0 = a.tile_off[tile], o1 = o0 + tot;
    const int64_t c0 = o0 > a.w0 ? o0 : a.w0, c1 = o1 < a.w1 ? o1 : a.w1;
    if (c0 >= c1) return;
    const uintptr_t gdst0 = (uintptr_t)a.dst + (uintptr_t)(o0 - a.w0); // address of stream byte o0 (may precede dst)
    const int       phase = (int)(gdst0 & 15);
    if (tot > a.text_cap) { // rare (thousands of tiny reads in one tile): format privately, store byte-wise with clipping
        uint8_t  loc[CE_PER * CE_MAX_SLOT_BYTES];
        uint8_t* e = cov_format_slots(loc, w0, nmine, cv, dg, a.reso, a.own_first);
        int64_t  x = o0 + ex;
        for (uint8_t* q = loc; q < e; q++, x++) if (x >= a.w0 && x < a.w1) a.dst[x - a.w0] = *q;
        return;
    }
    cov_format_slots(sbuf + phase + ex, w0, nmine, cv, dg, a.reso, a.own_first);
    __syncthreads();
    // store [c0, c1): head bytes, aligned 128-bit body, tail bytes
    const uintptr_t ga0 = gdst0 + (uintptr_t)(c0 - o0), ga1 = gdst0 + (uintptr_t)(c1 - o0);
    uintptr_t       fa = (ga0 + 15) & ~(uintptr_t)15, la = ga1 & ~(uintptr_t)15;
    if (fa > la) { fa = ga1; la = ga1; }
    const uint8_t* sb = sbuf + phase; // sb[x - gdst0] is the byte for address x
    for (uintptr_t x = ga0 + threadIdx.x; x < fa; x += CE_THREADS) *reinterpret_cast<uint8_t*>(x) = sb[x - gdst0];
    for (uintptr_t x = fa + (uintptr_t)threadIdx.x * 16; x < la; x += (uintptr_t)CE_THREADS * 16)
        stg_stream(reinterpret_cast<uint4*>(x), *reinterpret_cast<const uint4*>(sb + (x - gdst0)));
    for (uintptr_t x = la + threadIdx.x; x < ga1; x += CE_THREADS) *reinterpret_cast<uint8_t*>(x) = sb[x - gdst0];
}

// Persistent over tiles [tile_first, tile_first + n_tiles): the grid is sized by the launcher (all SM slots when the
// emitter runs alone, two CTAs per SM when it shares the GPU with the gather kernel on the other emit stream).
template <bool EMIT>
__global__ void __launch_bounds__(CE_THREADS, EMIT ? 6 : 8) k_cov_text(CovEmitArgs a, int64_t n_tiles)
{
    extern __shared__ __align__(16) uint8_t sbuf[];
    __shared__ int ws[34];
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        cov_text_tile<EMIT>(a, a.tile_first + t, sbuf, ws);
        __syncthreads(); // the shared text buffer and the scan scratch are reused by the next tile
    }
}

// one thread per read: every tile whose first slot lies in this read's slot range points at it
__global__ void __launch_bounds__(256) k_cov_tile_index(const int64_t* __restrict__ slot_off, int64_t m, int64_t n_tiles, int32_t* tile_read)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int64_t s0 = slot_off[i], s1 = slot_off[i + 1];
    for (int64_t T = (s0 + COV_TILE_SLOTS - 1) / COV_TILE_SLOTS; T * COV_TILE_SLOTS < s1 && T < n_tiles; T++) tile_read[T] = (int32_t)i;
    if (i == m - 1) tile_read[n_tiles] = (int32_t)(m - 1);
}
void launch_cov_tile_index(const int64_t* slot_off, int64_t m, int64_t n_slots, int32_t* tile_read, cudaStream_t st)
{
    int64_t T = (n_slots + COV_TILE_SLOTS - 1) / COV_TILE_SLOTS;
    if (m > 0) k_cov_tile_index<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(slot_off, m, T, tile_read);
}

int  cov_tiles(int64_t n_slots) { return (int)((n_slots + COV_TILE_SLOTS - 1) / COV_TILE_SLOTS); }
void launch_cov_sizes(const CovEmitArgs& a, cudaStream_t st)
{
    int t = cov_tiles(a.n_slots);
    if (t > 0) k_cov_text<false><<<t, CE_THREADS, 0, st>>>(a, (int64_t)t);
}
void launch_cov_emit(const CovEmitArgs& a_in, int64_t n_tiles_launch, cudaStream_t st)
{
    if (n_tiles_launch <= 0) return;
    CovEmitArgs a = a_in;
    a.text_cap = CE_CAP;
    if (const char* e = getenv("RAFT_B200_COV_CAP")) { int v = atoi(e); if (v >= 0 && v < CE_CAP) a.text_cap = v; } // test knob: force the direct path
    cudaFuncSetAttribute(k_cov_text<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CE_SMEM);
    // ctas_per_sm == 0: one CTA per tile (the loop in the kernel runs once)
    const int64_t grid = a.ctas_per_sm > 0 && n_tiles_launch > 148 * (int64_t)a.ctas_per_sm ? 148 * (int64_t)a.ctas_per_sm : n_tiles_launch;
    k_cov_text<true><<<(unsigned)grid, CE_THREADS, CE_SMEM, st>>>(a, n_tiles_launch);
}

// ================================================================ K5c long_repeats.txt
struct WinWriter {
    uint8_t* dst;
    int64_t  w0, w1, x; // x: stream position of the next byte
    __device__ __forceinline__ void put(uint8_t c) { if (x >= w0 && x < w1) dst[x - w0] = c; x++; }
    __device__ __forceinline__ void put_u64(uint64_t v)
    {
        int nd = dec_digits64(v);
        for (int k = 0; k < nd; k++) put(dec_digit_at(v, nd, k));
    }
    __device__ __forceinline__ void put_i32(int32_t v)
    {
        if (v < 0) { put('-'); put_u64((uint64_t)(-(int64_t)v)); } else put_u64((uint64_t)v);
    }
};

__global__ void __launch_bounds__(256) k_rep_emit(RepEmitArgs a)
{
    int64_t i = a.read_first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.read_last) return;
    int64_t o0 = a.line_off[i], o1 = a.line_off[i + 1];
    if (o1 <= a.w0 || o0 >= a.w1) return;
    WinWriter w{a.dst, a.w0, a.w1, o0};
    w.put('r'); w.put('e'); w.put('a'); w.put('d'); w.put(' ');
    w.put_u64((uint64_t)(a.own_first + i));
    w.put(','); w.put(' ');
    const int2* r = a.rep + a.rep_cap_off[i];
    for (int q = 0; q < a.rep_cnt[i]; q++) {
        w.put_i32(r[q].x); w.put(','); w.put_i32(r[q].y);
        w.put(' '); w.put(' '); w.put(' '); w.put(' ');
    }
    w.put('\n');
}
void launch_rep_emit(const RepEmitArgs& a, cudaStream_t st)
{
    int64_t cnt = a.read_last - a.read_first;
    if (cnt > 0) k_rep_emit<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(a);
}

// ---- long_repeats.bed (repeat.hpp:187-199): per repeat of a forward read  chr \t start+s \t start+e \n,
// of a reverse read  chr \t end-e \t end-s \n; other orientations write nothing.
__device__ __forceinline__ void bed_numbers(const SimInfo& si, int2 r, int* x, int* y)
{
    if (si.flags & 1) { *x = si.start_pos + r.x; *y = si.start_pos + r.y; }
    else { *x = si.end_pos - r.y; *y = si.end_pos - r.x; }
}
__global__ void __launch_bounds__(256) k_bed_sizes(const int32_t* rep_cnt, const int64_t* rep_cap_off, const int2* rep, const SimInfo* sim,
                                                    const int64_t* name_off, int64_t own_first, int64_t m, int32_t* line_size)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const SimInfo si = sim[i];
    int           sz = 0;
    if (si.flags) {
        const int   chr_len = (int)(name_off[own_first + i + 1] - name_off[own_first + i]) - si.tail_off - 1;
        const int2* r = rep + rep_cap_off[i];
        for (int q = 0; q < rep_cnt[i]; q++) { int x, y; bed_numbers(si, r[q], &x, &y); sz += chr_len + 1 + dec_len_i32(x) + 1 + dec_len_i32(y) + 1; }
    }
    line_size[i] = sz;
}
void launch_bed_sizes(const int32_t* rep_cnt, const int64_t* rep_cap_off, const int2* rep, const SimInfo* sim, const int64_t* name_off,
                      int64_t own_first, int64_t m, int32_t* line_size, cudaStream_t st)
{
    if (m > 0) k_bed_sizes<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(rep_cnt, rep_cap_off, rep, sim, name_off, own_first, m, line_size);
}
__global__ void __launch_bounds__(256) k_bed_emit(RepEmitArgs a)
{
    int64_t i = a.read_first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.read_last) return;
    int64_t o0 = a.line_off[i], o1 = a.line_off[i + 1];
    if (o1 <= a.w0 || o0 >= a.w1 || o1 == o0) return;
    const SimInfo  si = a.sim[i];
    const int64_t  nm0 = a.name_off[a.own_first + i];
    const int      chr_len = (int)(a.name_off[a.own_first + i + 1] - nm0) - si.tail_off - 1;
    const uint8_t* chr = a.names + nm0 + si.tail_off + 1;
    WinWriter      w{a.dst, a.w0, a.w1, o0};
    const int2*    r = a.rep + a.rep_cap_off[i];
    for (int q = 0; q < a.rep_cnt[i]; q++) {
        int x, y;
        bed_numbers(si, r[q], &x, &y);
        for (int k = 0; k < chr_len; k++) w.put(chr[k]);
        w.put('\t'); w.put_i32(x); w.put('\t'); w.put_i32(y); w.put('\n');
    }
}
void launch_bed_emit(const RepEmitArgs& a, cudaStream_t st)
{
    int64_t cnt = a.read_last - a.read_first;
    if (cnt > 0) k_bed_emit<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(a);
}

// ================================================================ K5b reads.fasta
// Output-tile-parallel gather, persistent and warp-specialised.  A tile is 16 KiB of the reads.fasta stream.
// Warp 0 is the producer.  For each tile its 32 lanes load 32 consecutive FragDesc entries starting at
// tile_frag[T] (one coalesced request; the entries of the NEXT tile are already in flight while the current
// one is laid out), a ballot + warp scan turns the records that intersect the tile into a piece list in the
// next free pipeline slot, and every lane issues the 1-D TMA bulk copy of its own piece from the read arena
// into that slot's stage buffer (a 16-byte aligned superset of the piece).  A lane whose record's header
// starts in the tile also fetches the name extent, so the consumers never chase name_off.  Warps 1..7 are
// consumers: they wait on the slot's `full` mbarrier (producer arrive + TMA transaction bytes), generate the
// few header bytes, realign the staged bases (two aligned 128-bit shared loads + funnel shifts) into aligned
// 128-bit streaming stores, and release the slot through its `empty` mbarrier.  Three slots keep two tiles
// of loads in flight per CTA behind the tile being stored.
constexpr int FE_THREADS = 256;
constexpr int FE_CONSUMERS = FE_THREADS - 32;       // 7 warps
constexpr int FE_SLOTS = 3;
constexpr int FE_MAXP = 32;                         // pieces per slot (one producer round)
constexpr int FE_STAGE = FASTA_TILE + 512;          // staged source bytes per slot (alignment slack; overflow spills to the next slot)

struct FePiece {
    long long O;       // stream offset of the record
    const uint8_t* src; // global address of the first sequence byte of the piece
    long long nm0;     // offset of the read's name in `names`
    int       frag;    // record index
    int       n;       // sequence bytes of the piece inside the tile (0: header / newline only)
    int       n_stage; // leading bytes available in shared memory (the rest is read from global)
    int       soff;    // offset of the first byte in the stage buffer
    int       h, len, read, fa;
    int       nl;      // name length
    int       pad;
};
static_assert(sizeof(FePiece) == 64, "FePiece layout");

struct __align__(16) FeSmem {
    uint8_t   stage[FE_SLOTS][FE_STAGE];
    FePiece   piece[FE_SLOTS][FE_MAXP];
    long long x0[FE_SLOTS], x1[FE_SLOTS];
    int       np[FE_SLOTS]; // -1: no more work
    uint64_t  full[FE_SLOTS], empty[FE_SLOTS];
};
static_assert(4 * (sizeof(FeSmem) + 1024) <= 228 * 1024, "four CTAs per SM");

// consumer threads (index ct of FE_CONSUMERS): n bytes from shared memory (any alignment) to global (any alignment)
__device__ __forceinline__ void consumers_copy_from_smem(uint8_t* __restrict__ dst, const uint8_t* sp, int n, int ct)
{
    int head = (int)((16 - ((uintptr_t)dst & 15)) & 15);
    if (head > n) head = n;
    for (int k = ct; k < head; k += FE_CONSUMERS) dst[k] = sp[k];
    const int nbody = (n - head) >> 4;
    uint4*    d16 = reinterpret_cast<uint4*>(dst + head);
    const uint8_t* s = sp + head;
    for (int c = ct; c < nbody; c += FE_CONSUMERS) stg_stream(d16 + c, lds_unaligned16(s + (c << 4)));
    const int done = head + (nbody << 4);
    for (int k = done + ct; k < n; k += FE_CONSUMERS) dst[k] = sp[k];
}

__global__ void __launch_bounds__(256) k_fasta_tile_index(const int64_t* __restrict__ frag_off, int64_t G, int64_t n_tiles, int32_t* tile_frag)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    int64_t o0 = frag_off[g], o1 = frag_off[g + 1];
    for (int64_t T = (o0 + FASTA_TILE - 1) / FASTA_TILE; T * FASTA_TILE < o1 && T < n_tiles; T++) tile_frag[T] = (int32_t)g;
}
void launch_fasta_tile_index(const int64_t* frag_off, int64_t G, int32_t* tile_frag, cudaStream_t st)
{
    // the caller sizes tile_frag for ceil(total/FASTA_TILE) tiles; every tile start lies inside exactly one record
    if (G > 0) k_fasta_tile_index<<<(unsigned)((G + 255) / 256), 256, 0, st>>>(frag_off, G, (int64_t)1 << 62, tile_frag);
}

__global__ void __launch_bounds__(256) k_frag_desc(const int32_t* __restrict__ frag_read, const int32_t* __restrict__ frag_a,
                                                   const int32_t* __restrict__ frag_b, const int32_t* __restrict__ frag_size,
                                                   const int64_t* __restrict__ frag_off, const int64_t* __restrict__ seq_off, int64_t G, FragDesc* desc)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g > G) return;
    FragDesc d{};
    d.out_off = frag_off[g];
    if (g < G) {
        int i = frag_read[g], fa = frag_a[g], fb = frag_b[g];
        d.src_off = seq_off[i] + fa; d.len = fb - fa; d.hdr_len = frag_size[g] - (fb - fa) - 1; d.read = i; d.a = fa;
    }
    desc[g] = d;
}
void launch_frag_desc(const int32_t* frag_read, const int32_t* frag_a, const int32_t* frag_b, const int32_t* frag_size, const int64_t* frag_off,
                      const int64_t* seq_off, int64_t G, FragDesc* desc, cudaStream_t st)
{
    k_frag_desc<<<(unsigned)((G + 1 + 255) / 256), 256, 0, st>>>(frag_read, frag_a, frag_b, frag_size, frag_off, seq_off, G, desc);
}

__global__ void k_frag_sample(const FragDesc* __restrict__ desc, int64_t G, int step, int64_t* out2)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t g = k * step;
    if (g > G) return;
    out2[2 * k] = desc[g].out_off; out2[2 * k + 1] = desc[g].src_off;
}
void launch_frag_sample(const FragDesc* desc, int64_t G, int step, int64_t* out2, cudaStream_t st)
{
    int64_t cnt = G / step + 1;
    k_frag_sample<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(desc, G, step, out2);
}

// Header bytes of one record (all modes).  q indexes the header, 0 = '>'.
struct FeHeader {
    uint64_t       num;      // read= number
    const uint8_t* name;     // read name
    int            nl, dn;
    unsigned       fa, fb, kk;
    int            da, db, dk;
    int            vx, vy, vl, dx, dy, dl, align_off, align_len, tail_off, tl;
};
__device__ __forceinline__ uint8_t fe_signed_char_at(int v, int nd, int q)
{
    if (v < 0) { if (q == 0) return '-'; return dec_digit_at32((uint32_t)(-(int64_t)v), nd - 1, q - 1); }
    return dec_digit_at32((uint32_t)v, nd, q);
}
template <int MODE> // 0 real, 1 sim, 2 split_naive
__device__ __forceinline__ uint8_t fe_header_char(const FeHeader& H, int q)
{
    if (MODE == 2) {
        // ">" name "_" k "\n"   (split_naive.cpp:32)
        if (q < 1) return '>';
        if ((q -= 1) < H.nl) return H.name[q];
        if ((q -= H.nl) < 1) return '_';
        if ((q -= 1) < H.dk) return dec_digit_at32(H.kk, H.dk, q);
        return '\n';
    }
    if (q < 6) return (uint8_t)(">read="[q]);
    if ((q -= 6) < H.dn) return (H.num >> 32) ? dec_digit_at(H.num, H.dn, q) : dec_digit_at32((uint32_t)H.num, H.dn, q);
    if ((q -= H.dn) < 1) return ',';
    q -= 1;
    if (MODE == 0) {
        // ">read=" num "," name ",pos_on_original_read=" a "-" b "\n"   (chop.hpp:261-265, 314-318)
        if (q < H.nl) return H.name[q];
        if ((q -= H.nl) < 22) return (uint8_t)(",pos_on_original_read="[q]);
        if ((q -= 22) < H.da) return dec_digit_at32(H.fa, H.da, q);
        if ((q -= H.da) < 1) return '-';
        if ((q -= 1) < H.db) return dec_digit_at32(H.fb, H.db, q);
        return '\n';
    }
    // ">read=" num "," align ",position=" x "-" y ",length=" ln tail "\n"   (chop.hpp:252-258, 293-310)
    if (q < H.align_len) return H.name[H.align_off + q];
    if ((q -= H.align_len) < 10) return (uint8_t)(",position="[q]);
    if ((q -= 10) < H.dx) return fe_signed_char_at(H.vx, H.dx, q);
    if ((q -= H.dx) < 1) return '-';
    if ((q -= 1) < H.dy) return fe_signed_char_at(H.vy, H.dy, q);
    if ((q -= H.dy) < 8) return (uint8_t)(",length="[q]);
    if ((q -= 8) < H.dl) return fe_signed_char_at(H.vl, H.dl, q);
    if ((q -= H.dl) < H.tl) return H.name[H.tail_off + q];
    return '\n';
}

template <int MODE>
__global__ void __launch_bounds__(FE_THREADS, 4) k_fasta_emit(FastaEmitArgs a, int64_t n_tiles)
{
    extern __shared__ __align__(16) uint8_t fe_raw[];
    FeSmem& s = *reinterpret_cast<FeSmem*>(fe_raw);
    if (threadIdx.x == 0) {
        for (int k = 0; k < FE_SLOTS; k++) { mbar_init(&s.full[k], 1); mbar_init(&s.empty[k], FE_CONSUMERS / 32); }
    }
    __syncthreads();
    const int64_t T0 = a.w0 / FASTA_TILE;

    if (threadIdx.x < 32) {
        // ================= producer warp =================
        const int      lane = threadIdx.x;
        const unsigned lt = (1u << lane) - 1u;
        int            slot = 0;
        unsigned       ph = 0;
        const int64_t  step = gridDim.x;
        auto load_desc = [&](int64_t g) {
            FragDesc d;
            if (g <= a.G) d = a.desc[g];                       // entry G: out_off = stream length (a terminator)
            else { d.out_off = INT64_MAX / 2; d.src_off = 0; d.len = 0; d.hdr_len = 0; d.read = 0; d.a = 0; }
            return d;
        };
        // software pipeline: tile_frag two tiles ahead, descriptors one tile ahead
        int64_t  t = blockIdx.x;
        int64_t  g_cur = t < n_tiles ? a.tile_frag[T0 + t] : 0;
        int64_t  g_nxt = t + step < n_tiles ? a.tile_frag[T0 + t + step] : 0;
        FragDesc d_cur = load_desc(t < n_tiles ? g_cur + lane : INT64_MAX);
        for (;; t += step) {
            const bool    done = t >= n_tiles;
            const int64_t T = T0 + t, xs = T * FASTA_TILE;
            const int64_t x0 = xs > a.w0 ? xs : a.w0;
            const int64_t x1 = (xs + FASTA_TILE) < a.w1 ? (xs + FASTA_TILE) : a.w1;
            int64_t       g = g_cur;                    // record containing the tile's first byte
            FragDesc      d = d_cur;
            g_cur = g_nxt;
            g_nxt = t + 2 * step < n_tiles ? a.tile_frag[T0 + t + 2 * step] : 0;
            d_cur = load_desc(t + step < n_tiles ? g_cur + lane : INT64_MAX);
            bool more = true, first = true;
            while (more) {
                if (!first) d = load_desc(g + lane);
                first = false;
                mbar_wait(&s.empty[slot], ph ^ 1);   // the consumers are done with this slot's previous contents
                if (done) {
                    if (lane == 0) { s.np[slot] = -1; mbar_arrive(&s.full[slot]); }
                    break;
                }
                const int64_t O = d.out_off;
                const bool    valid = O < x1;
                const int64_t s0 = O + d.hdr_len, s1 = s0 + d.len;
                const bool    emit = valid && s1 + 1 > x0;                    // else the window starts after this record
                const int64_t p0 = s0 > x0 ? s0 : x0, p1 = s1 < x1 ? s1 : x1;
                const int     n = emit && p1 > p0 ? (int)(p1 - p0) : 0;
                const int64_t srcoff = d.src_off + (p0 - s0);                 // offset in the arena
                const int64_t al = srcoff & ~(int64_t)15;
                int64_t       end = (srcoff + n + 15) & ~(int64_t)15;
                if (end > a.seq_safe_end) end = a.seq_safe_end;               // never read past what the arena guarantees
                const int      bytes = n > 0 && end > al ? (int)(end - al) : 0;
                const int      incl = warp_inclusive_sum(bytes);
                const unsigned emit_m = __ballot_sync(0xffffffffu, emit);
                const int      idx = __popc(emit_m & lt);
                const unsigned nofit_m = __ballot_sync(0xffffffffu, emit && (idx >= FE_MAXP || incl > FE_STAGE));
                const unsigned inval_m = __ballot_sync(0xffffffffu, !valid);
                const int      cut = nofit_m ? __ffs(nofit_m) - 1 : 32;       // lane 0 always fits
                const int      fin = inval_m ? __ffs(inval_m) - 1 : 32;
                const int64_t  gb = g;
                int            limit;
                if (fin < 32 && fin <= cut) { limit = fin; more = false; }
                else { limit = cut; g += cut; }                               // another slot for the rest of this tile
                const unsigned lim_m = limit >= 32 ? 0xffffffffu : ((1u << limit) - 1u);
                const int      used = limit > 0 ? __shfl_sync(0xffffffffu, incl, limit - 1) : 0;
                const bool     mine = emit && lane < limit;
                if (mine) {
                    FePiece& pc = s.piece[slot][idx];
                    const int soff = incl - bytes + (int)(srcoff - al);
                    const int avail = bytes - (int)(srcoff - al);
                    pc.O = O; pc.src = a.seq + srcoff; pc.frag = (int)(gb + lane); pc.n = n;
                    pc.n_stage = avail < 0 ? 0 : (avail > n ? n : avail); pc.soff = soff;
                    pc.h = d.hdr_len; pc.len = d.len; pc.read = d.read; pc.fa = d.a;
                    long long nm0 = 0;
                    int       nl = 0;
                    if (O < x1 && s0 > x0) {                                  // some header byte lies in the window
                        const int64_t gid = a.own_first + d.read;
                        nm0 = a.name_off[gid];
                        nl = (int)(a.name_off[gid + 1] - nm0);
                    }
                    pc.nm0 = nm0; pc.nl = nl;
                }
                if (lane == 0) { s.np[slot] = __popc(emit_m & lim_m); s.x0[slot] = x0; s.x1[slot] = x1; }
                __syncwarp();
                if (lane == 0) {
                    if (used > 0) mbar_expect_tx(&s.full[slot], (uint32_t)used); // arrive (release: the list above is visible) + expected bytes
                    else mbar_arrive(&s.full[slot]);
                }
                __syncwarp();
                if (mine && bytes > 0) tma_load_1d(s.stage[slot] + (incl - bytes), a.seq + al, (uint32_t)bytes, &s.full[slot]);
                if (++slot == FE_SLOTS) { slot = 0; ph ^= 1; }
            }
            if (done) break;
        }
        return;
    }

    // ================= consumers =================
    const int ct = threadIdx.x - 32;
    int       slot = 0;
    unsigned  ph = 0;
    for (;;) {
        mbar_wait(&s.full[slot], ph);
        const int np = s.np[slot];
        if (np < 0) break;
        const int64_t x0 = s.x0[slot], x1 = s.x1[slot];
        for (int k = 0; k < np; k++) {
            const FePiece& pc = s.piece[slot][k];
            const int64_t  O = pc.O;
            const int      h = pc.h;
            const int64_t  hp0 = O > x0 ? O : x0, hp1 = (O + h) < x1 ? (O + h) : x1;
            // the header byte this thread owns is computed first (its name load is in flight during the copy
            // below) and stored last; headers longer than the consumer group are finished on the spot
            uint8_t hc = 0;
            bool    hv = false;
            if (hp0 + ct < hp1) {
                FeHeader H;
                H.num = (uint64_t)(a.read_num_base + pc.frag + 1);
                H.name = a.names + pc.nm0; H.nl = pc.nl; H.dn = dec_digits64(H.num);
                if (MODE == 2) {
                    H.kk = (unsigned)(pc.fa / a.split_len + 1); H.dk = dec_digits(H.kk);
                } else if (MODE == 0) {
                    H.fa = (unsigned)pc.fa; H.fb = (unsigned)(pc.fa + pc.len); H.da = dec_digits(H.fa); H.db = dec_digits(H.fb);
                } else {
                    const SimInfo si = a.sim[pc.read];
                    const int     L = (int)(a.seq_off[pc.read + 1] - a.seq_off[pc.read]);
                    sim_header_numbers(si, pc.len == L && pc.fa == 0, pc.fa, pc.fa + pc.len, L, &H.vx, &H.vy, &H.vl);
                    H.dx = dec_len_i32(H.vx); H.dy = dec_len_i32(H.vy); H.dl = dec_len_i32(H.vl);
                    H.align_off = si.align_off; H.align_len = si.align_len; H.tail_off = si.tail_off; H.tl = pc.nl - si.tail_off;
                }
                hc = fe_header_char<MODE>(H, (int)(hp0 + ct - O));
                hv = true;
                for (int64_t x = hp0 + ct + FE_CONSUMERS; x < hp1; x += FE_CONSUMERS) a.dst[x - a.w0] = fe_header_char<MODE>(H, (int)(x - O));
            }
            const int64_t s0 = O + h, s1 = s0 + pc.len;
            if (ct == 0 && s1 >= x0 && s1 < x1) a.dst[s1 - a.w0] = '\n';
            if (pc.n > 0) {
                uint8_t* d = a.dst + ((s0 > x0 ? s0 : x0) - a.w0);
                consumers_copy_from_smem(d, s.stage[slot] + pc.soff, pc.n_stage, ct);
                for (int q = pc.n_stage + ct; q < pc.n; q += FE_CONSUMERS) d[q] = pc.src[q]; // arena tail not covered by the bulk copy
            }
            if (hv) a.dst[hp0 + ct - a.w0] = hc;
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&s.empty[slot]); // this warp no longer reads the slot
        if (++slot == FE_SLOTS) { slot = 0; ph ^= 1; }
    }
}
void launch_fasta_emit(const FastaEmitArgs& a, cudaStream_t st)
{
    if (a.w1 <= a.w0 || a.G <= 0) return;
    int64_t tiles = (a.w1 - 1) / FASTA_TILE - a.w0 / FASTA_TILE + 1;
    int64_t grid = tiles < 148 * 4 ? tiles : 148 * 4;
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FeSmem));
        kern<<<(unsigned)grid, FE_THREADS, sizeof(FeSmem), st>>>(a, tiles);
    };
    if (a.split_len > 0) go(k_fasta_emit<2>);
    else if (a.sim) go(k_fasta_emit<1>);
    else go(k_fasta_emit<0>);
}

__global__ void k_sample_i64(const int64_t* __restrict__ src, int64_t n, int step, int64_t cnt, int64_t* dst)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt) return;
    int64_t i = k * step;
    dst[k] = src[i < n ? i : n - 1];
}
void launch_sample_i64(const int64_t* src, int64_t n, int step, int64_t* dst, cudaStream_t st)
{
    int64_t cnt = (n - 1) / step + 2;
    k_sample_i64<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(src, n, step, cnt, dst);
}

// ================================================================ digest
__global__ void __launch_bounds__(256) k_digest(const uint8_t* __restrict__ buf, int64_t n, int64_t abs_off, unsigned long long* acc)
{
    unsigned long long d = 0;
    const int64_t      stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        d += mix64((uint64_t)(abs_off + i) * 257ull + buf[i] + 1ull);
    d = warp_sum(d);
    __shared__ unsigned long long ws[8];
    if (lane_id() == 0) ws[warp_id()] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; w++) t += ws[w];
        atomicAdd(acc, t);
    }
}
void launch_digest(const uint8_t* buf, int64_t n, int64_t abs_off, unsigned long long* acc, cudaStream_t st)
{
    if (n <= 0) return;
    int64_t blocks = (n + 256 * 16 - 1) / (256 * 16);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_digest<<<(unsigned)blocks, 256, 0, st>>>(buf, n, abs_off, acc);
}

} // namespace raftk
