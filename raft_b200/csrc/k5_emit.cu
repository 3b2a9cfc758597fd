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
    const int64_t per_sm = a.ctas_per_sm > 0 ? a.ctas_per_sm : 6;
    const int64_t grid = n_tiles_launch < 148 * per_sm ? n_tiles_launch : 148 * per_sm;
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
// Warp 0 (one lane) is the producer: for each tile it walks the records that intersect it (tile_frag[T] gives
// the first, FragDesc has everything else in one 32-byte load), writes a piece list into the next free
// pipeline slot and issues one 1-D TMA bulk copy per sequence piece from the read arena into that slot's
// stage buffer (a 16-byte aligned superset of the piece).  Warps 1..7 are consumers: they wait on the
// slot's `full` mbarrier (producer arrive + TMA transaction bytes), generate the few header bytes,
// realign the staged bases (two aligned 128-bit shared loads + funnel shifts) into aligned 128-bit
// streaming stores, and release the slot through its `empty` mbarrier.  With two slots the walk and the
// TMA latency of tile i+1 hide behind the stores of tile i.
constexpr int FE_THREADS = 256;
constexpr int FE_CONSUMERS = FE_THREADS - 32;       // 7 warps
constexpr int FE_SLOTS = 2;
constexpr int FE_MAXP = 24;                         // pieces per slot
constexpr int FE_STAGE = FASTA_TILE + FE_MAXP * 32; // staged source bytes per slot

struct FePiece {
    long long frag;   // record index
    long long O;      // stream offset of the record
    long long p0;     // stream position of the first sequence byte of the piece
    const uint8_t* src; // global address of that byte
    int       n;      // sequence bytes of the piece inside the tile (0: header / newline only)
    int       n_stage; // leading bytes available in shared memory (the rest is read from global)
    int       soff;   // offset of the first byte in the stage buffer
    int       bulk;   // bytes moved by the piece's TMA copy
    int       h, len, read, fa;
};

struct __align__(16) FeSmem {
    uint8_t   stage[FE_SLOTS][FE_STAGE];
    FePiece   piece[FE_SLOTS][FE_MAXP];
    long long x0[FE_SLOTS], x1[FE_SLOTS];
    int       np[FE_SLOTS]; // -1: no more work
    uint64_t  full[FE_SLOTS], empty[FE_SLOTS];
};

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
        // ================= producer (one lane) =================
        if (threadIdx.x != 0) return;
        int      slot = 0;
        unsigned ph = 0;
        for (int64_t t = blockIdx.x;; t += gridDim.x) {
            const bool    done = t >= n_tiles;
            const int64_t T = T0 + t, xs = T * FASTA_TILE;
            const int64_t x0 = xs > a.w0 ? xs : a.w0;
            const int64_t x1 = (xs + FASTA_TILE) < a.w1 ? (xs + FASTA_TILE) : a.w1;
            int64_t       g = done ? 0 : a.tile_frag[T]; // record containing the tile's first byte
            bool          more = true;
            while (more) {
                mbar_wait(&s.empty[slot], ph ^ 1);   // the consumers are done with this slot's previous contents
                int np = 0, used = 0;
                more = false;
                if (done) {
                    np = -1;
                } else {
                    FePiece* pl = s.piece[slot];
                    for (; g < a.G; g++) {
                        const FragDesc d = a.desc[g];
                        const int64_t  O = d.out_off;
                        if (O >= x1) break;
                        const int64_t s0 = O + d.hdr_len, s1 = s0 + d.len;
                        if (s1 + 1 <= x0) continue;                               // the window starts after this record
                        if (np == FE_MAXP) { more = true; break; }
                        const int64_t p0 = s0 > x0 ? s0 : x0, p1 = s1 < x1 ? s1 : x1;
                        FePiece&      pc = pl[np];
                        pc.frag = g; pc.O = O; pc.p0 = p0; pc.n = p1 > p0 ? (int)(p1 - p0) : 0; pc.n_stage = 0; pc.soff = 0; pc.bulk = 0; pc.src = nullptr;
                        pc.h = d.hdr_len; pc.len = d.len; pc.read = d.read; pc.fa = d.a;
                        if (pc.n > 0) {
                            const int64_t srcoff = d.src_off + (p0 - s0);         // offset in the arena
                            const int64_t al = srcoff & ~(int64_t)15;
                            int64_t       end = (srcoff + pc.n + 15) & ~(int64_t)15;
                            if (end > a.seq_safe_end) end = a.seq_safe_end;       // never read past what the arena guarantees
                            const int bytes = end > al ? (int)(end - al) : 0;
                            if (used + bytes > FE_STAGE) { more = true; break; }  // another slot for the rest of this tile
                            pc.src = a.seq + srcoff;
                            pc.soff = used + (int)(srcoff - al);
                            const int avail = bytes - (int)(srcoff - al);
                            pc.n_stage = avail < 0 ? 0 : (avail > pc.n ? pc.n : avail);
                            pc.bulk = bytes;
                            used += bytes;
                        }
                        np++;
                    }
                }
                s.np[slot] = np; s.x0[slot] = x0; s.x1[slot] = x1;
                if (used > 0) {
                    mbar_expect_tx(&s.full[slot], (uint32_t)used); // arrive (release: the list above is visible) + expected bytes
                    int off = 0;
                    for (int k = 0; k < np; k++) {
                        const FePiece& pc = s.piece[slot][k];
                        if (pc.bulk > 0) {
                            tma_load_1d(s.stage[slot] + off, pc.src - (pc.soff - off), (uint32_t)pc.bulk, &s.full[slot]);
                            off += pc.bulk;
                        }
                    }
                } else {
                    mbar_arrive(&s.full[slot]);
                }
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
            if (hp0 < hp1) {
                const int64_t  gid = a.own_first + pc.read;
                const uint64_t num = (uint64_t)(a.read_num_base + pc.frag + 1);
                const int64_t  nm0 = a.name_off[gid];
                const int      nl = (int)(a.name_off[gid + 1] - nm0);
                const int      dn = dec_digits64(num);
                if (a.split_len > 0) {
                    // ">" name "_" k "\n"   (split_naive.cpp:32)
                    const unsigned kk = (unsigned)(pc.fa / a.split_len + 1);
                    const int      dk = dec_digits(kk);
                    for (int64_t x = hp0 + ct; x < hp1; x += FE_CONSUMERS) {
                        int     q = (int)(x - O);
                        uint8_t c;
                        if (q < 1) c = '>';
                        else if ((q -= 1) < nl) c = a.names[nm0 + q];
                        else if ((q -= nl) < 1) c = '_';
                        else if ((q -= 1) < dk) c = dec_digit_at32(kk, dk, q);
                        else c = '\n';
                        a.dst[x - a.w0] = c;
                    }
                } else if (!a.sim) {
                    const unsigned fa = (unsigned)pc.fa, fb = (unsigned)(pc.fa + pc.len);
                    const int      da = dec_digits(fa), db = dec_digits(fb);
                    for (int64_t x = hp0 + ct; x < hp1; x += FE_CONSUMERS) {
                        int     q = (int)(x - O);
                        uint8_t c;
                        if (q < 6) c = (uint8_t)(">read="[q]);
                        else if ((q -= 6) < dn) c = (num >> 32) ? dec_digit_at(num, dn, q) : dec_digit_at32((uint32_t)num, dn, q);
                        else if ((q -= dn) < 1) c = ',';
                        else if ((q -= 1) < nl) c = a.names[nm0 + q];
                        else if ((q -= nl) < 22) c = (uint8_t)(",pos_on_original_read="[q]);
                        else if ((q -= 22) < da) c = dec_digit_at32(fa, da, q);
                        else if ((q -= da) < 1) c = '-';
                        else if ((q -= 1) < db) c = dec_digit_at32(fb, db, q);
                        else c = '\n';
                        a.dst[x - a.w0] = c;
                    }
                } else {
                    // ">read=" num "," align ",position=" x "-" y ",length=" ln tail "\n"   (chop.hpp:252-258, 293-310)
                    const SimInfo si = a.sim[pc.read];
                    const int     L = (int)(a.seq_off[pc.read + 1] - a.seq_off[pc.read]);
                    int           vx, vy, vl;
                    sim_header_numbers(si, pc.len == L && pc.fa == 0, pc.fa, pc.fa + pc.len, L, &vx, &vy, &vl);
                    const int dx = dec_len_i32(vx), dy = dec_len_i32(vy), dl = dec_len_i32(vl), tl = nl - si.tail_off;
                    auto signed_char_at = [](int v, int nd, int q) -> uint8_t {
                        if (v < 0) { if (q == 0) return '-'; return dec_digit_at32((uint32_t)(-(int64_t)v), nd - 1, q - 1); }
                        return dec_digit_at32((uint32_t)v, nd, q);
                    };
                    for (int64_t x = hp0 + ct; x < hp1; x += FE_CONSUMERS) {
                        int     q = (int)(x - O);
                        uint8_t c;
                        if (q < 6) c = (uint8_t)(">read="[q]);
                        else if ((q -= 6) < dn) c = (num >> 32) ? dec_digit_at(num, dn, q) : dec_digit_at32((uint32_t)num, dn, q);
                        else if ((q -= dn) < 1) c = ',';
                        else if ((q -= 1) < si.align_len) c = a.names[nm0 + si.align_off + q];
                        else if ((q -= si.align_len) < 10) c = (uint8_t)(",position="[q]);
                        else if ((q -= 10) < dx) c = signed_char_at(vx, dx, q);
                        else if ((q -= dx) < 1) c = '-';
                        else if ((q -= 1) < dy) c = signed_char_at(vy, dy, q);
                        else if ((q -= dy) < 8) c = (uint8_t)(",length="[q]);
                        else if ((q -= 8) < dl) c = signed_char_at(vl, dl, q);
                        else if ((q -= dl) < tl) c = a.names[nm0 + si.tail_off + q];
                        else c = '\n';
                        a.dst[x - a.w0] = c;
                    }
                }
            }
            const int64_t s1 = O + h + pc.len;
            if (ct == 0 && s1 >= x0 && s1 < x1) a.dst[s1 - a.w0] = '\n';
            if (pc.n > 0) {
                uint8_t* d = a.dst + (pc.p0 - a.w0);
                consumers_copy_from_smem(d, s.stage[slot] + pc.soff, pc.n_stage, ct);
                for (int q = pc.n_stage + ct; q < pc.n; q += FE_CONSUMERS) d[q] = pc.src[q]; // arena tail not covered by the bulk copy
            }
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
    cudaFuncSetAttribute(k_fasta_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FeSmem));
    k_fasta_emit<<<(unsigned)grid, FE_THREADS, sizeof(FeSmem), st>>>(a, tiles);
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
