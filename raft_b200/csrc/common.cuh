// common.cuh — device utilities shared by the raft_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace raftk {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// ------------------------------------------------------------------ warp / block scans
template <typename T>
__device__ __forceinline__ T warp_inclusive_sum(T v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(FULL, v, d);
        if (lane_id() >= d) v += o;
    }
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

// Block-wide exclusive sum over one value per thread. `ws` is shared scratch of >= 33 T.
// Returns the exclusive prefix; *total receives the block sum. Contains two __syncthreads().
template <typename T, int NT>
__device__ __forceinline__ T block_exclusive_sum(T v, T* ws, T* total)
{
    constexpr int NW = NT / 32;
    T inc = warp_inclusive_sum(v);
    if (lane_id() == 31) ws[warp_id()] = inc;
    __syncthreads();
    if (warp_id() == 0) {
        T w = lane_id() < NW ? ws[lane_id()] : T(0);
        T wi = warp_inclusive_sum(w);
        if (lane_id() < NW) ws[lane_id()] = wi - w;
        if (lane_id() == NW - 1) ws[32] = wi;
    }
    __syncthreads();
    T res = inc - v + ws[warp_id()];
    *total = ws[32];
    return res;
}

// ------------------------------------------------------------------ decoupled look-back
// One 64-bit status word per tile: bits 63..62 = flag (0 empty, 1 aggregate, 2 inclusive prefix),
// low 62 bits = value (mod 2^62).  Tiles must obtain their index from an atomic ticket so that
// every predecessor is already resident (forward progress).
constexpr uint64_t LB_MASK = (1ull << 62) - 1;
constexpr uint64_t LB_AGG = 1ull << 62;
constexpr uint64_t LB_PREFIX = 2ull << 62;

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by ALL lanes of one warp. Returns the sum of aggregates of tiles [0, tile) (mod 2^62).
__device__ __forceinline__ uint64_t lookback_exclusive(const uint64_t* status, int tile)
{
    uint64_t sum = 0;
    int      look = tile - 1;
    while (look >= 0) {
        int      idx = look - lane_id();
        uint64_t w = LB_PREFIX; // tiles before 0: prefix 0
        if (idx >= 0) {
            w = ld_relaxed_u64(status + idx);
            while ((w >> 62) == 0) {
                __nanosleep(20);
                w = ld_relaxed_u64(status + idx);
            }
        }
        unsigned pm = __ballot_sync(FULL, (w >> 62) == 2);
        int      stop = pm ? (__ffs(pm) - 1) : 31;
        uint64_t c = (lane_id() <= stop) ? (w & LB_MASK) : 0ull;
        sum += warp_sum(c);
        if (pm) break;
        look -= 32;
    }
    return sum & LB_MASK;
}

// Thread 0 of the block publishes `agg`, warp 0 looks back, thread 0 publishes the inclusive prefix.
// Returns the exclusive prefix to every thread (through shared `*bcast`). Contains __syncthreads().
__device__ __forceinline__ uint64_t lookback_block(uint64_t* status, int tile, uint64_t agg, uint64_t* bcast)
{
    if (threadIdx.x == 0) st_relaxed_u64(status + tile, (tile == 0 ? LB_PREFIX : LB_AGG) | (agg & LB_MASK));
    if (warp_id() == 0) {
        uint64_t ex = (tile == 0) ? 0ull : lookback_exclusive(status, tile);
        if (lane_id() == 0) {
            if (tile != 0) st_relaxed_u64(status + tile, LB_PREFIX | ((ex + agg) & LB_MASK));
            *bcast = ex;
        }
    }
    __syncthreads();
    return *bcast;
}

// Split form: publish the aggregate early, wait for the prefix later (work can sit in between).
__device__ __forceinline__ void lookback_publish(uint64_t* status, int tile, uint64_t agg)
{ // one thread
    st_relaxed_u64(status + tile, (tile == 0 ? LB_PREFIX : LB_AGG) | (agg & LB_MASK));
}
__device__ __forceinline__ uint64_t lookback_wait(uint64_t* status, int tile, uint64_t agg, uint64_t* bcast)
{ // all threads; contains __syncthreads()
    if (warp_id() == 0) {
        uint64_t ex = (tile == 0) ? 0ull : lookback_exclusive(status, tile);
        if (lane_id() == 0) {
            if (tile != 0) st_relaxed_u64(status + tile, LB_PREFIX | ((ex + agg) & LB_MASK));
            *bcast = ex;
        }
    }
    __syncthreads();
    return *bcast;
}

__device__ __forceinline__ int64_t lb_signed(uint64_t v)
{ // sign-extend a 62-bit two's complement value
    return (int64_t)(v << 2) >> 2;
}

// ------------------------------------------------------------------ decimal text
static __device__ __constant__ unsigned c_pow10[10] = {1u, 10u, 100u, 1000u, 10000u, 100000u, 1000000u, 10000000u, 100000000u, 1000000000u};
// decimal digits of v: bit length * log10(2) (1233/4096), corrected by one table compare
__device__ __forceinline__ int dec_digits(uint32_t v)
{
    v |= 1u;
    int t = ((32 - __clz(v)) * 1233) >> 12;
    return t + (v >= c_pow10[t]);
}
__device__ __forceinline__ int dec_digits64(uint64_t v)
{
    if (v < 4000000000ull) return dec_digits((uint32_t)v);
    int d = 10;
    uint64_t p = 10000000000ull;
    while (d < 20 && v >= p) { d++; p *= 10ull; }
    return d;
}
// length of operator<<(int): '-' for negatives
__device__ __forceinline__ int dec_len_i32(int32_t v)
{
    return v < 0 ? 1 + dec_digits((uint32_t)(-(int64_t)v)) : dec_digits((uint32_t)v);
}
// writes the decimal text of v ending just before `end`; returns pointer to first char
__device__ __forceinline__ uint8_t* put_dec_back(uint8_t* end, uint32_t v)
{
    do {
        uint32_t q = v / 10u;
        *--end = (uint8_t)('0' + (v - q * 10u));
        v = q;
    } while (v);
    return end;
}
// v has exactly nd decimal digits
__device__ __forceinline__ uint8_t* put_u32_nd(uint8_t* p, uint32_t v, int nd)
{
    put_dec_back(p + nd, v);
    return p + nd;
}
__device__ __forceinline__ uint8_t* put_i32(uint8_t* p, int32_t v)
{
    uint32_t u = (uint32_t)v;
    if (v < 0) { *p++ = '-'; u = (uint32_t)(-(int64_t)v); }
    int n = dec_digits(u);
    put_dec_back(p + n, u);
    return p + n;
}
// digit `k` (0 = most significant) of v which has nd digits
__device__ __forceinline__ uint8_t dec_digit_at(uint64_t v, int nd, int k)
{
    for (int i = nd - 1 - k; i > 0; --i) v /= 10ull;
    return (uint8_t)('0' + (uint32_t)(v % 10ull));
}

__device__ __forceinline__ uint8_t dec_digit_at32(uint32_t v, int nd, int k)
{
    for (int i = nd - 1 - k; i > 0; --i) v /= 10u;
    return (uint8_t)('0' + (v % 10u));
}

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

// ------------------------------------------------------------------ streaming loads / stores
__device__ __forceinline__ uint4 ldg_stream(const uint4* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(uint4* p, const uint4& v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ------------------------------------------------------------------ TMA 1-D bulk copy (global -> shared) + mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// dst (shared, 16B aligned), src (global, 16B aligned), bytes multiple of 16
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 16 bytes starting at an arbitrarily aligned shared-memory address: two aligned 128-bit loads + funnel shifts
__device__ __forceinline__ uint4 lds_unaligned16(const uint8_t* sp)
{
    const unsigned sa = (unsigned)(smem_u32(sp) & 15u);
    const uint4*   b = reinterpret_cast<const uint4*>(sp - sa);
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


} // namespace raftk
