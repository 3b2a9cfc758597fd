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
#include "kernels.h"

namespace raftk {

// ================================================================ K5a coverage.txt
constexpr int CE_THREADS = 128;
constexpr int CE_PER = COV_TILE_SLOTS / CE_THREADS; // 8 slots per thread
constexpr int CE_MAX_SLOT_BYTES = 40;               // "read 2147483647 " (16) + "2147483647,-2147483648 " (23)
constexpr int CE_SMEM = COV_TILE_SLOTS * CE_MAX_SLOT_BYTES + 32;

__device__ __forceinline__ int64_t find_read(const int64_t* __restrict__ slot_off, int64_t m, int64_t g)
{ // largest i in [0,m) with slot_off[i] <= g
    int64_t lo = 0, hi = m;
    while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (slot_off[mid] <= g) lo = mid; else hi = mid; }
    return lo;
}

template <bool EMIT>
__global__ void __launch_bounds__(CE_THREADS) k_cov_text(CovEmitArgs a)
{
    extern __shared__ __align__(16) uint8_t sbuf[];
    __shared__ int ws[34];
    const int64_t  tile = a.tile_first + blockIdx.x;
    const int64_t  g0 = tile * COV_TILE_SLOTS + (int64_t)threadIdx.x * CE_PER;
    int            sizes[CE_PER];
    int            mine = 0;
    int64_t        ri = 0, rs = 0, re = 0; // current read, its first slot, one past its last slot
    if (g0 < a.n_slots) { ri = find_read(a.slot_off, a.m, g0); rs = a.slot_off[ri]; re = a.slot_off[ri + 1]; }
    {
        int64_t r_ = ri, rs_ = rs, re_ = re;
#pragma unroll
        for (int k = 0; k < CE_PER; k++) {
            int64_t g = g0 + k;
            int     sz = 0;
            if (g < a.n_slots) {
                while (g >= re_) { r_++; rs_ = re_; re_ = a.slot_off[r_ + 1]; }
                int64_t bin = g - rs_;
                if (bin == 0) sz += 5 + dec_digits64((uint64_t)(a.own_first + r_)) + 1;
                if (g == re_ - 1) sz += 1;
                else sz += dec_digits((uint32_t)(bin * a.reso)) + 1 + dec_len_i32(a.cov[g]) + 1;
            }
            sizes[k] = sz; mine += sz;
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
    {
        uint8_t* p = sbuf + phase + ex;
        int64_t  r_ = ri, rs_ = rs, re_ = re;
#pragma unroll
        for (int k = 0; k < CE_PER; k++) {
            int64_t g = g0 + k;
            if (g < a.n_slots) {
                while (g >= re_) { r_++; rs_ = re_; re_ = a.slot_off[r_ + 1]; }
                int64_t bin = g - rs_;
                if (bin == 0) {
                    p[0] = 'r'; p[1] = 'e'; p[2] = 'a'; p[3] = 'd'; p[4] = ' '; p += 5;
                    uint64_t id = (uint64_t)(a.own_first + r_);
                    int      nd = dec_digits64(id);
                    for (int d = nd - 1; d >= 0; d--) { p[d] = (uint8_t)('0' + (unsigned)(id % 10ull)); id /= 10ull; }
                    p += nd; *p++ = ' ';
                }
                if (g == re_ - 1) {
                    *p++ = '\n';
                } else {
                    p = put_i32(p, (int32_t)(bin * a.reso)); *p++ = ',';
                    p = put_i32(p, a.cov[g]); *p++ = ' ';
                }
            }
        }
    }
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

int  cov_tiles(int64_t n_slots) { return (int)((n_slots + COV_TILE_SLOTS - 1) / COV_TILE_SLOTS); }
void launch_cov_sizes(const CovEmitArgs& a, cudaStream_t st)
{
    int t = cov_tiles(a.n_slots);
    if (t > 0) k_cov_text<false><<<t, CE_THREADS, 0, st>>>(a);
}
void launch_cov_emit(const CovEmitArgs& a, int64_t n_tiles_launch, cudaStream_t st)
{
    if (n_tiles_launch <= 0) return;
    cudaFuncSetAttribute(k_cov_text<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CE_SMEM);
    k_cov_text<true><<<(unsigned)n_tiles_launch, CE_THREADS, CE_SMEM, st>>>(a);
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

// ================================================================ K5b reads.fasta
constexpr int FE_THREADS = 256;

// 16 bytes starting at arbitrary address p (all inside the source arena)
__device__ __forceinline__ uint4 load_unaligned16(const uint8_t* p)
{
    const unsigned sa = (unsigned)((uintptr_t)p & 15);
    const uint4*   b = reinterpret_cast<const uint4*>(p - sa);
    uint4          q0 = b[0];
    if (sa == 0) return q0;
    uint4          q1 = b[1];
    const unsigned bs = (sa & 3u) * 8u;
    uint4          o;
    switch (sa >> 2) {
    case 0:
        o.x = __funnelshift_r(q0.x, q0.y, bs); o.y = __funnelshift_r(q0.y, q0.z, bs);
        o.z = __funnelshift_r(q0.z, q0.w, bs); o.w = __funnelshift_r(q0.w, q1.x, bs);
        break;
    case 1:
        o.x = __funnelshift_r(q0.y, q0.z, bs); o.y = __funnelshift_r(q0.z, q0.w, bs);
        o.z = __funnelshift_r(q0.w, q1.x, bs); o.w = __funnelshift_r(q1.x, q1.y, bs);
        break;
    case 2:
        o.x = __funnelshift_r(q0.z, q0.w, bs); o.y = __funnelshift_r(q0.w, q1.x, bs);
        o.z = __funnelshift_r(q1.x, q1.y, bs); o.w = __funnelshift_r(q1.y, q1.z, bs);
        break;
    default:
        o.x = __funnelshift_r(q0.w, q1.x, bs); o.y = __funnelshift_r(q1.x, q1.y, bs);
        o.z = __funnelshift_r(q1.y, q1.z, bs); o.w = __funnelshift_r(q1.z, q1.w, bs);
        break;
    }
    return o;
}

// all threads of the block copy n bytes; dst and src arbitrarily aligned
__device__ __forceinline__ void block_copy(uint8_t* dst, const uint8_t* src, int64_t n)
{
    int64_t head = (int64_t)((16 - ((uintptr_t)dst & 15)) & 15);
    if (head > n) head = n;
    for (int64_t k = threadIdx.x; k < head; k += FE_THREADS) dst[k] = src[k];
    const int64_t nbody = (n - head) >> 4;
    uint4*        d16 = reinterpret_cast<uint4*>(dst + head);
    const uint8_t* s = src + head;
    for (int64_t c = threadIdx.x; c < nbody; c += FE_THREADS) stg_stream(d16 + c, load_unaligned16(s + (c << 4)));
    const int64_t done = head + (nbody << 4);
    for (int64_t k = done + threadIdx.x; k < n; k += FE_THREADS) dst[k] = src[k];
}

__global__ void __launch_bounds__(FE_THREADS) k_fasta_emit(FastaEmitArgs a)
{
    const int64_t lead = (int64_t)((uintptr_t)a.dst & 15);
    const int64_t xs = a.w0 - lead + (int64_t)blockIdx.x * FASTA_TILE;
    const int64_t x0 = xs > a.w0 ? xs : a.w0;
    const int64_t x1 = (xs + FASTA_TILE) < a.w1 ? (xs + FASTA_TILE) : a.w1;
    if (x0 >= x1) return;
    // first record intersecting the tile: largest g with frag_off[g] <= x0
    int64_t lo = 0, hi = a.G;
    while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (a.frag_off[mid] <= x0) lo = mid; else hi = mid; }
    for (int64_t g = lo; g < a.G; g++) {
        const int64_t O = a.frag_off[g];
        if (O >= x1) break;
        const int64_t  O1 = a.frag_off[g + 1];
        const int64_t  i = a.frag_read[g];
        const int      fa = a.frag_a[g], fb = a.frag_b[g];
        const int64_t  gid = a.own_first + i;
        const int64_t  nm0 = a.name_off[gid];
        const int      nl = (int)(a.name_off[gid + 1] - nm0);
        const uint64_t num = (uint64_t)(a.read_num_base + g + 1);
        const int      dn = dec_digits64(num), da = dec_digits((uint32_t)fa), db = dec_digits((uint32_t)fb);
        const int      h = 6 + dn + 1 + nl + 22 + da + 1 + db + 1;
        const int64_t  len = (int64_t)fb - fa;
        // header bytes [O, O+h)
        {
            int64_t p0 = O > x0 ? O : x0, p1 = (O + h) < x1 ? (O + h) : x1;
            for (int64_t x = p0 + threadIdx.x; x < p1; x += FE_THREADS) {
                int     k = (int)(x - O);
                uint8_t c;
                if (k < 6) c = (uint8_t)(">read="[k]);
                else if ((k -= 6) < dn) c = dec_digit_at(num, dn, k);
                else if ((k -= dn) < 1) c = ',';
                else if ((k -= 1) < nl) c = a.names[nm0 + k];
                else if ((k -= nl) < 22) c = (uint8_t)(",pos_on_original_read="[k]);
                else if ((k -= 22) < da) c = dec_digit_at((uint64_t)fa, da, k);
                else if ((k -= da) < 1) c = '-';
                else if ((k -= 1) < db) c = dec_digit_at((uint64_t)fb, db, k);
                else c = '\n';
                a.dst[x - a.w0] = c;
            }
        }
        // sequence bytes [O+h, O+h+len)
        {
            int64_t s0 = O + h, s1 = s0 + len;
            int64_t p0 = s0 > x0 ? s0 : x0, p1 = s1 < x1 ? s1 : x1;
            if (p0 < p1) block_copy(a.dst + (p0 - a.w0), a.seq + a.seq_off[i] + fa + (p0 - s0), p1 - p0);
            if (threadIdx.x == 0 && s1 >= x0 && s1 < x1) a.dst[s1 - a.w0] = '\n';
        }
        (void)O1;
    }
}
void launch_fasta_emit(const FastaEmitArgs& a, cudaStream_t st)
{
    if (a.w1 <= a.w0 || a.G <= 0) return;
    int64_t lead = (int64_t)((uintptr_t)a.dst & 15);
    int64_t tiles = (a.w1 - a.w0 + lead + FASTA_TILE - 1) / FASTA_TILE;
    k_fasta_emit<<<(unsigned)tiles, FE_THREADS, 0, st>>>(a);
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
