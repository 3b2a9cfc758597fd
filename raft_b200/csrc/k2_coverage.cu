// k2_coverage.cu — K2: per-read bin coverage.
//
// Replaces profileCoverage (repeat.hpp:28-79).  The reference sorts a read's intervals and
// increments every covered bin; here each contributing interval adds +1 at its first bin and -1
// one past its last bin in a per-read difference array, and one device-wide inclusive scan turns
// the array into coverage.  Every read owns nb+1 slots (nb = ceil(L/reso), repeat.hpp:32-37) so
// each read's slots sum to zero and a plain (unsegmented) scan restarts at 0 on every read
// boundary: no head flags, no segmented operator.
//
// Which intervals contribute (repeat.hpp:48-58, chop.hpp:165-169): the query side of every record;
// the target side too when overlaps are not symmetric and target != query.
// Bin range of [s, e) (repeat.hpp:62-77): lo = max(s,0)/reso, bins lo..(e-1)/reso when e-1 >= lo*reso.
#include "coverage.cuh"

namespace raftk {

// ---------------------------------------------------------------- per-read layout
__global__ void k_read_layout(const int64_t* seq_off, int64_t m, int reso, int p, int P, int l, int32_t* slots, int32_t* rep_cap,
                              int32_t* cut_cap)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int64_t L = seq_off[i + 1] - seq_off[i];
    int64_t nb = (L + reso - 1) / reso;
    slots[i] = (int32_t)(nb + 1);
    // a repeat needs a run of >= ceil(p/reso) bins (repeat.hpp:125) and runs are separated by >= 1 bin
    int64_t minbins = ((int64_t)p + reso - 1) / reso;
    if (minbins < 1) minbins = 1;
    rep_cap[i] = (int32_t)((nb + 1) / (minbins + 1));
    // stars: 0, P, ..., floor(L/P)*P (+ L) (chop.hpp:209-223); fragments <= ceil((stars-1)/div) (chop.hpp:270-276)
    int64_t nstars = L / P + 1 + (L % P != 0);
    int64_t div = l / P;
    int64_t fmax = (nstars - 1 + div - 1) / div;
    if (fmax < 1) fmax = 1;
    cut_cap[i] = (int32_t)(fmax - 1);
}
void launch_read_layout(const int64_t* seq_off, int64_t m, int reso, int p, int P, int l, int32_t* slots, int32_t* rep_cap, int32_t* cut_cap,
                        cudaStream_t st)
{
    if (m > 0) k_read_layout<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(seq_off, m, reso, p, P, l, slots, rep_cap, cut_cap);
}

// ---------------------------------------------------------------- small scan: int32[n] -> exclusive int64[n+1]
constexpr int SS_THREADS = 256;
constexpr int SS_ITEMS = 8;
constexpr int SS_TILE = SS_THREADS * SS_ITEMS;

__global__ void __launch_bounds__(SS_THREADS) k_scan_i32_to_i64(const int32_t* __restrict__ in, int64_t* __restrict__ out, int64_t n,
                                                                uint64_t* status, int* ticket)
{
    __shared__ long long ws[34];
    __shared__ uint64_t  bcast;
    __shared__ int       s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1);
    __syncthreads();
    const int     tile = s_tile;
    const int64_t base = (int64_t)tile * SS_TILE + (int64_t)threadIdx.x * SS_ITEMS;
    long long     v[SS_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < SS_ITEMS; k++) { v[k] = (base + k < n) ? (long long)in[base + k] : 0ll; sum += v[k]; }
    long long tot;
    long long ex = block_exclusive_sum<long long, SS_THREADS>(sum, ws, &tot);
    uint64_t  pre = lookback_block(status, tile, (uint64_t)tot, &bcast);
    long long run = lb_signed(pre) + ex;
#pragma unroll
    for (int k = 0; k < SS_ITEMS; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
        if (base + k == n - 1) out[n] = run;
    }
    if (n == 0 && tile == 0 && threadIdx.x == 0) out[0] = 0;
}
int  scan_tiles_small(int64_t n) { return (int)((n + SS_TILE - 1) / SS_TILE) + (n == 0); }
void launch_scan_i32_to_i64(const int32_t* in, int64_t* out, int64_t n, uint64_t* status, int* ticket, cudaStream_t st)
{
    int tiles = scan_tiles_small(n);
    cudaMemsetAsync(status, 0, sizeof(uint64_t) * tiles, st);
    cudaMemsetAsync(ticket, 0, sizeof(int), st);
    k_scan_i32_to_i64<<<tiles, SS_THREADS, 0, st>>>(in, out, n, status, ticket);
}

// ---------------------------------------------------------------- coverage scan: in-place inclusive, int32
// Tile = 256 threads x 16 ints; warp w owns 512 consecutive ints read as 4 rounds of fully
// coalesced 128-bit loads (lane l of round r holds ints [w*512 + r*128 + 4l, +4)).
constexpr int CS_THREADS = 256;
constexpr int CS_ROUNDS = 4;
constexpr int CS_TILE = CS_THREADS * 4 * CS_ROUNDS; // 4096 ints = 16 KiB

// Sum over bins b in [b0, b1) of the decimal length of b*reso (b*reso < 2^31): every bin has one digit, and one more
// for every power of ten it reaches.
__device__ __forceinline__ int64_t pos_digit_sum(int64_t b0, int64_t b1, int reso)
{
    int64_t sum = b1 - b0, p = 10;
    for (int d = 1; d <= 9; d++, p *= 10) {
        const int64_t thr = (p + reso - 1) / reso; // first bin with b*reso >= 10^d
        if (thr >= b1) break;
        sum += b1 - (b0 > thr ? b0 : thr);
    }
    return sum;
}
__global__ void __launch_bounds__(256) k_cov_static(const int64_t* __restrict__ slot_off, int64_t m, int64_t own_first, int reso, int32_t* tile_static)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int64_t s0 = slot_off[i], s1 = slot_off[i + 1]; // bins s0 .. s1-2, sentinel s1-1
    const int     pre = 5 + dec_digits64((uint64_t)(own_first + i)) + 1;
    for (int64_t T = s0 / COV_TILE_SLOTS; T * COV_TILE_SLOTS < s1; T++) {
        const int64_t lo = T * COV_TILE_SLOTS > s0 ? T * COV_TILE_SLOTS : s0;
        const int64_t hi = (T + 1) * COV_TILE_SLOTS < s1 ? (T + 1) * COV_TILE_SLOTS : s1;
        const int64_t b1 = (hi < s1 ? hi : s1 - 1) - s0;
        int64_t       sz = b1 > lo - s0 ? pos_digit_sum(lo - s0, b1, reso) : 0;
        if (lo == s0) sz += pre;
        if (hi == s1) sz -= 2; // the sentinel is one newline; the scan counts digits(0) + 2 for it
        atomicAdd(tile_static + T, (int32_t)sz);
    }
}
void launch_cov_static_sizes(const int64_t* slot_off, int64_t m, int64_t n_slots, int64_t own_first, int reso, int32_t* tile_static, cudaStream_t st)
{
    const int64_t T = (n_slots + COV_TILE_SLOTS - 1) / COV_TILE_SLOTS;
    if (T <= 0 || m <= 0) return;
    cudaMemsetAsync(tile_static, 0, sizeof(int32_t) * (size_t)T, st);
    k_cov_static<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(slot_off, m, own_first, reso, tile_static);
}

// The scan is memory-bound with most issue slots idle, so it also sizes coverage.txt: every thread knows the final
// coverage of its 16 slots and adds digits(cov) + 2 per slot into the 1024-slot tile counters the emitter (K5a) uses
// (the counters start from the coverage-independent part, launch_cov_static_sizes).
__device__ __forceinline__ int cov_dyn_bytes(int c)
{
    const unsigned u = c < 0 ? (unsigned)(-(int64_t)c) : (unsigned)c;
    return (u < 100u ? (u < 10u ? 3 : 4) : dec_digits(u) + 2) + (c < 0);
}
// The slots of every read sum to zero (each +1 has its -1 at or before the read's sentinel), so the running sum
// entering a tile is just the sum of the differences from the start of the read that spans the tile boundary
// (tile_read gives that read) up to the tile: on average half a read (~1 KiB) re-read per 16 KiB tile, and no
// communication between tiles.  Only when that read starts more than CS_BACK slots before the tile (reads longer
// than ~200 kbp at -r 50) does the tile wait for its predecessor's running sum, which every tile publishes as soon
// as it knows it; tile ids come from an atomic ticket so a waited-for tile is always resident or finished.  The
// scan is out of place (diff -> cov) because the backward re-read must see differences, not coverages.
constexpr int CS_BACK = CS_TILE; // longest backward re-read

__global__ void __launch_bounds__(CS_THREADS) k_scan_cov(const int32_t* __restrict__ diff, int32_t* __restrict__ cov, int64_t n, uint64_t* status,
                                                         int* ticket, CovSizeArgs cs)
{
    __shared__ int     ws[34], wp[34];
    __shared__ int     s_tile;
    __shared__ int64_t s_rs;
    if (threadIdx.x == 0) {
        const int t = atomicAdd(ticket, 1);
        s_tile = t;
        s_rs = cs.slot_off[cs.tile_read[(int64_t)t * (CS_TILE / COV_TILE_SLOTS)]]; // first slot of the read holding the tile's first slot
    }
    const int lane = lane_id(), warp = warp_id();
    __syncthreads();
    const int     tile = s_tile;
    const int64_t t0 = (int64_t)tile * CS_TILE, tbase = t0 + (int64_t)warp * (128 * CS_ROUNDS);
    const int64_t rs = s_rs;
    const bool    local = t0 - rs <= CS_BACK; // uniform
    int4          v[CS_ROUNDS];
    int           rsum[CS_ROUNDS];
    const bool    full = (int64_t)(tile + 1) * CS_TILE <= n;
#pragma unroll
    for (int r = 0; r < CS_ROUNDS; r++) {
        int64_t i = tbase + r * 128 + lane * 4;
        if (full) {
            v[r] = *reinterpret_cast<const int4*>(diff + i);
        } else {
            v[r].x = i + 0 < n ? diff[i + 0] : 0; v[r].y = i + 1 < n ? diff[i + 1] : 0;
            v[r].z = i + 2 < n ? diff[i + 2] : 0; v[r].w = i + 3 < n ? diff[i + 3] : 0;
        }
        rsum[r] = v[r].x + v[r].y + v[r].z + v[r].w;
    }
    int part = 0; // this thread's share of the differences between the spanning read's start and the tile
    if (local)
        for (int64_t i = rs + threadIdx.x; i < t0; i += CS_THREADS) part += diff[i];
    // lane-exclusive prefix inside each round, rounds chained
    int lane_ex[CS_ROUNDS], wtot = 0;
#pragma unroll
    for (int r = 0; r < CS_ROUNDS; r++) {
        int inc = warp_inclusive_sum(rsum[r]);
        lane_ex[r] = wtot + inc - rsum[r];
        wtot += __shfl_sync(FULL, inc, 31);
    }
    // block: exclusive prefix over the warp totals, and the sum of the carry parts, in one pass
    part = warp_sum(part);
    if (lane == 0) { ws[warp] = wtot; wp[warp] = part; }
    __syncthreads();
    if (warp == 0) {
        constexpr int NW = CS_THREADS / 32;
        int w = lane < NW ? ws[lane] : 0, q = lane < NW ? wp[lane] : 0;
        int wi = warp_inclusive_sum(w);
        q = warp_sum(q);
        if (lane < NW) ws[lane] = wi - w;
        if (lane == NW - 1) {
            int carry = q;
            if (!local) { // the spanning read started long before: take the predecessor's running sum
                uint64_t x;
                while (((x = ld_relaxed_u64(status + tile - 1)) >> 62) == 0) __nanosleep(40);
                carry = (int)lb_signed(x & LB_MASK);
            }
            st_relaxed_u64(status + tile, LB_PREFIX | ((uint64_t)(int64_t)(carry + wi) & LB_MASK)); // running sum after this tile
            wp[32] = carry;
        }
    }
    __syncthreads();
    int p0 = wp[32] + ws[warp];
    int text_bytes = 0;
#pragma unroll
    for (int r = 0; r < CS_ROUNDS; r++) {
        int64_t i = tbase + r * 128 + lane * 4;
        int     a = p0 + lane_ex[r];
        int4    o;
        o.x = a + v[r].x; o.y = o.x + v[r].y; o.z = o.y + v[r].z; o.w = o.z + v[r].w;
        if (cs.tile_bytes) {
            if (full) text_bytes += cov_dyn_bytes(o.x) + cov_dyn_bytes(o.y) + cov_dyn_bytes(o.z) + cov_dyn_bytes(o.w);
            else text_bytes += (i + 0 < n ? cov_dyn_bytes(o.x) : 0) + (i + 1 < n ? cov_dyn_bytes(o.y) : 0) +
                               (i + 2 < n ? cov_dyn_bytes(o.z) : 0) + (i + 3 < n ? cov_dyn_bytes(o.w) : 0);
        }
        if (full) {
            *reinterpret_cast<int4*>(cov + i) = o;
        } else {
            if (i + 0 < n) cov[i + 0] = o.x; if (i + 1 < n) cov[i + 1] = o.y;
            if (i + 2 < n) cov[i + 2] = o.z; if (i + 3 < n) cov[i + 3] = o.w;
        }
    }
    if (cs.tile_bytes) { // a warp's 512 slots lie inside one 1024-slot text tile
        text_bytes = warp_sum(text_bytes);
        if (lane == 0 && text_bytes) atomicAdd(cs.tile_bytes + tbase / COV_TILE_SLOTS, text_bytes);
    }
}
int  scan_tiles_cov(int64_t n) { return (int)((n + CS_TILE - 1) / CS_TILE); }
void launch_scan_cov(const int32_t* diff, int32_t* cov, int64_t n, uint64_t* status, int* ticket, const CovSizeArgs& cs, cudaStream_t st)
{
    int tiles = scan_tiles_cov(n);
    if (tiles == 0) return;
    cudaMemsetAsync(status, 0, sizeof(uint64_t) * tiles, st);
    cudaMemsetAsync(ticket, 0, sizeof(int), st);
    if (cs.tile_bytes)
        cudaMemcpyAsync(cs.tile_bytes, cs.tile_static, sizeof(int32_t) * (size_t)((n + COV_TILE_SLOTS - 1) / COV_TILE_SLOTS), cudaMemcpyDeviceToDevice, st);
    k_scan_cov<<<tiles, CS_THREADS, 0, st>>>(diff, cov, n, status, ticket, cs);
}

// ---------------------------------------------------------------- scatter
__global__ void __launch_bounds__(256) k_scatter_records(ScatterArgs a)
{
    const bool    sym = *a.sym_flag != 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < a.n_rec; k += stride) {
        int     q = a.qid[k], t = a.tid[k];
        int64_t lq = (int64_t)q - a.own_first;
        if (!a.skip_query && lq >= 0 && lq < a.own_count)
            if (!add_interval(a.diff, a.slot_off, lq, a.qs[k], a.qe[k], a.reso)) err_min(a.err, RAFTK_E_RANGE, k);
        if (!sym && t != q) {
            int64_t lt = (int64_t)t - a.own_first;
            if (lt >= 0 && lt < a.own_count)
                if (!add_interval(a.diff, a.slot_off, lt, a.ts[k], a.te[k], a.reso)) err_min(a.err, RAFTK_E_RANGE, k);
        }
    }
}
void launch_scatter_records(const ScatterArgs& a, cudaStream_t st)
{
    if (a.n_rec <= 0) return;
    int64_t blocks = (a.n_rec + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_scatter_records<<<(unsigned)blocks, 256, 0, st>>>(a);
}

__global__ void __launch_bounds__(256) k_scatter_endpoints(const int32_t* __restrict__ ep, int64_t n, const int64_t* slot_off, int32_t* diff,
                                                            int reso, int64_t own_first, int64_t own_count, ErrState* err)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        int64_t lr = (int64_t)ep[3 * k] - own_first;
        if (lr < 0 || lr >= own_count) { err_min(err, RAFTK_E_RANGE, k); continue; }
        if (!add_interval(diff, slot_off, lr, ep[3 * k + 1], ep[3 * k + 2], reso)) err_min(err, RAFTK_E_RANGE, k);
    }
}
void launch_scatter_endpoints(const int32_t* ep, int64_t n, const int64_t* slot_off, int32_t* diff, int reso, int64_t own_first,
                              int64_t own_count, ErrState* err, cudaStream_t st)
{
    if (n <= 0) return;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_scatter_endpoints<<<(unsigned)blocks, 256, 0, st>>>(ep, n, slot_off, diff, reso, own_first, own_count, err);
}

// ---------------------------------------------------------------- multi-GPU routing
// Every contributing interval becomes one endpoint (global read id, s, e) for the rank owning the read.
__device__ __forceinline__ int owner_of(const int64_t* __restrict__ bounds, int nranks, int64_t id)
{
    int lo = 0, hi = nranks; // bounds[lo] <= id < bounds[hi]
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (id >= bounds[mid]) lo = mid; else hi = mid; }
    return lo;
}

// Shared-memory counters per destination are hit by every endpoint of a block: with 2..64 destinations that is a handful of
// addresses under 256 threads.  The lanes of a warp that go to the same destination are therefore grouped first
// (__match_any_sync): one atomic per group, the lanes take consecutive positions behind its result.
// All 32 lanes call it; inactive lanes pass dest = -1.  Returns this lane's position (old value + rank in its group).
__device__ __forceinline__ unsigned warp_grouped_add(unsigned* counters, int dest)
{
    const unsigned grp = __match_any_sync(FULL, dest);
    const int      leader = __ffs(grp) - 1;
    unsigned       old = 0;
    if (dest >= 0 && lane_id() == leader) old = atomicAdd(&counters[dest], (unsigned)__popc(grp));
    old = __shfl_sync(FULL, old, leader);
    return old + (unsigned)__popc(grp & ((1u << lane_id()) - 1u));
}

// Two-pass packing straight from the records (used when the collected list of k_route_collect overflowed): every block
// counts its chunk's endpoints per destination, reserves a range per destination, then writes.
__global__ void __launch_bounds__(256) k_route_pack(ScatterArgs a, int nranks, const int64_t* __restrict__ bounds, unsigned long long* counters,
                                                    int32_t* sendbuf)
{
    extern __shared__ unsigned long long s_base[]; // nranks reserved bases, then nranks 32-bit counts / cursors
    unsigned*  s_cnt = reinterpret_cast<unsigned*>(s_base + nranks);
    const bool sym = *a.sym_flag != 0;
    for (int r = threadIdx.x; r < nranks; r += blockDim.x) s_cnt[r] = 0;
    __syncthreads();
    // each block owns a contiguous chunk of records so that packing is deterministic per block
    int64_t per = (a.n_rec + gridDim.x - 1) / gridDim.x;
    int64_t k0 = (int64_t)blockIdx.x * per, k1 = k0 + per < a.n_rec ? k0 + per : a.n_rec;
    // endpoints of reads this rank owns never leave it (raftgpu_accumulate_local scatters them directly)
    const int64_t own_lo = a.own_first, own_hi = a.own_first + a.own_count;
    const int64_t kend = k0 + ((k1 - k0 + blockDim.x - 1) / blockDim.x) * blockDim.x; // whole warps walk the chunk together
    // pass 1: count
    for (int64_t k = k0 + threadIdx.x; k < kend; k += blockDim.x) {
        int dq = -1, dt = -1;
        if (k < k1) {
            const int q = a.qid[k], t = a.tid[k];
            if (q < own_lo || q >= own_hi) dq = owner_of(bounds, nranks, q);
            if (!sym && t != q && (t < own_lo || t >= own_hi)) dt = owner_of(bounds, nranks, t);
        }
        warp_grouped_add(s_cnt, dq);
        warp_grouped_add(s_cnt, dt);
    }
    __syncthreads();
    // pass 2: reserve a range per destination, then write
    for (int r = threadIdx.x; r < nranks; r += blockDim.x) {
        s_base[r] = s_cnt[r] ? atomicAdd(&counters[r], (unsigned long long)s_cnt[r]) : 0ull;
        s_cnt[r] = 0;
    }
    __syncthreads();
    for (int64_t k = k0 + threadIdx.x; k < kend; k += blockDim.x) {
        int dq = -1, dt = -1, q = 0, t = 0;
        if (k < k1) {
            q = a.qid[k]; t = a.tid[k];
            if (q < own_lo || q >= own_hi) dq = owner_of(bounds, nranks, q);
            if (!sym && t != q && (t < own_lo || t >= own_hi)) dt = owner_of(bounds, nranks, t);
        }
        const unsigned pq = warp_grouped_add(s_cnt, dq), pt = warp_grouped_add(s_cnt, dt);
        if (dq >= 0) {
            const unsigned long long slot = s_base[dq] + pq;
            sendbuf[3 * slot] = q; sendbuf[3 * slot + 1] = a.qs[k]; sendbuf[3 * slot + 2] = a.qe[k];
        }
        if (dt >= 0) {
            const unsigned long long slot = s_base[dt] + pt;
            sendbuf[3 * slot] = t; sendbuf[3 * slot + 1] = a.ts[k]; sendbuf[3 * slot + 2] = a.te[k];
        }
    }
}
// Count pass that also collects: blocks that saw an endpoint for another rank reserve a range of the bounded list with one
// atomic and write (read, start, end, destination) there, so that packing only touches the collected endpoints -- with
// query-grouped symmetric PAF they are a few thousand out of 10^8 records.  counters[0..nranks) = endpoints per destination,
// *list_n = endpoints collected (may exceed cap: then the list is incomplete and the caller packs with k_route_pack).
__global__ void __launch_bounds__(256) k_route_collect(ScatterArgs a, int nranks, const int64_t* __restrict__ bounds, unsigned long long* counters,
                                                       int4* list, unsigned long long* list_n, unsigned long long cap)
{
    extern __shared__ unsigned long long s_base[]; // [0] = base of this block's range in the list, then nranks + 1 32-bit counters
    unsigned*  s_cnt = reinterpret_cast<unsigned*>(s_base + 1); // [0..nranks) per destination, [nranks] = block total / cursor
    const bool sym = *a.sym_flag != 0;
    for (int r = threadIdx.x; r < nranks + 1; r += blockDim.x) s_cnt[r] = 0;
    __syncthreads();
    // a block owns a contiguous chunk of records (a multiple of 1024, so 128-bit loads of the id columns stay aligned)
    const int64_t per = (((a.n_rec + gridDim.x - 1) / gridDim.x) + 1023) & ~(int64_t)1023;
    const int64_t k0 = (int64_t)blockIdx.x * per, k1 = k0 + per < a.n_rec ? k0 + per : a.n_rec;
    const int64_t own_lo = a.own_first, own_hi = a.own_first + a.own_count;
    unsigned      mine = 0; // endpoints this thread saw (summed per warp, one atomic per warp at the end)
    auto count_one = [&](int q, int t, bool real) { // all lanes of the warp call it together; !real: padding behind the chunk
        int dq = -1, dt = -1;
        if (real && (q < own_lo || q >= own_hi)) dq = owner_of(bounds, nranks, q);
        if (real && !sym && t != q && (t < own_lo || t >= own_hi)) dt = owner_of(bounds, nranks, t);
        const unsigned any = __ballot_sync(FULL, dq >= 0 || dt >= 0);
        if (!any) return; // the usual case with query-grouped symmetric PAF: nothing leaves the rank
        warp_grouped_add(s_cnt, dq);
        warp_grouped_add(s_cnt, dt);
        mine += (unsigned)(dq >= 0) + (unsigned)(dt >= 0);
    };
    const int64_t step = 4 * (int64_t)blockDim.x;
    const int64_t kend = k0 + ((k1 - k0 + step - 1) / step) * step;
    for (int64_t k = k0 + 4 * (int64_t)threadIdx.x; k < kend; k += step) {
        int4 q4 = make_int4(0, 0, 0, 0), t4 = q4;
        if (k + 4 <= k1) { // four records per thread and iteration: two 128-bit loads in flight
            q4 = *reinterpret_cast<const int4*>(a.qid + k); t4 = *reinterpret_cast<const int4*>(a.tid + k);
        } else {
            if (k < k1) { q4.x = a.qid[k]; t4.x = a.tid[k]; }
            if (k + 1 < k1) { q4.y = a.qid[k + 1]; t4.y = a.tid[k + 1]; }
            if (k + 2 < k1) { q4.z = a.qid[k + 2]; t4.z = a.tid[k + 2]; }
        }
        count_one(q4.x, t4.x, k < k1); count_one(q4.y, t4.y, k + 1 < k1); count_one(q4.z, t4.z, k + 2 < k1); count_one(q4.w, t4.w, k + 3 < k1);
    }
    mine = warp_sum(mine);
    if (lane_id() == 0 && mine) atomicAdd(&s_cnt[nranks], mine);
    __syncthreads();
    const unsigned total = s_cnt[nranks];
    if (total == 0) return;
    for (int r = threadIdx.x; r < nranks; r += blockDim.x)
        if (s_cnt[r]) atomicAdd(&counters[r], (unsigned long long)s_cnt[r]);
    if (threadIdx.x == 0) { s_base[0] = atomicAdd(list_n, (unsigned long long)total); s_cnt[nranks] = 0; }
    __syncthreads();
    const unsigned long long base = s_base[0];
    if (base + total > cap) return; // the list cannot hold this block's endpoints: the caller falls back to the two-pass packing
    for (int64_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
        int q = a.qid[k], t = a.tid[k];
        if (q < own_lo || q >= own_hi)
            list[base + atomicAdd(&s_cnt[nranks], 1u)] = make_int4(q, a.qs[k], a.qe[k], owner_of(bounds, nranks, q));
        if (!sym && t != q && (t < own_lo || t >= own_hi))
            list[base + atomicAdd(&s_cnt[nranks], 1u)] = make_int4(t, a.ts[k], a.te[k], owner_of(bounds, nranks, t));
    }
}
// buckets the collected endpoints by destination: cursors[d] = first free endpoint slot of destination d in sendbuf
__global__ void __launch_bounds__(256) k_route_pack_list(const int4* __restrict__ list, int64_t n, unsigned long long* cursors, int32_t* sendbuf)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int4               e = list[k];
    const unsigned long long slot = atomicAdd(&cursors[e.w], 1ull);
    sendbuf[3 * slot] = e.x; sendbuf[3 * slot + 1] = e.y; sendbuf[3 * slot + 2] = e.z;
}
static unsigned route_blocks(int64_t n) { int64_t b = (n + 4095) / 4096; if (b < 1) b = 1; if (b > 148 * 8) b = 148 * 8; return (unsigned)b; }
void launch_route_collect(const ScatterArgs& a, int nranks, const int64_t* bounds, unsigned long long* counts, int4* list,
                          unsigned long long* list_n, unsigned long long cap, cudaStream_t st)
{
    if (a.n_rec <= 0) return;
    k_route_collect<<<route_blocks(a.n_rec), 256, sizeof(unsigned long long) + sizeof(unsigned) * (nranks + 2), st>>>(a, nranks, bounds, counts, list, list_n, cap);
}
void launch_route_pack_list(const int4* list, int64_t n, unsigned long long* cursors, int32_t* sendbuf, cudaStream_t st)
{
    if (n > 0) k_route_pack_list<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(list, n, cursors, sendbuf);
}
void launch_route_pack(const ScatterArgs& a, int nranks, const int64_t* bounds, unsigned long long* cursors, int32_t* sendbuf, cudaStream_t st)
{
    if (a.n_rec <= 0) return;
    k_route_pack<<<route_blocks(a.n_rec), 256, (sizeof(unsigned long long) + sizeof(unsigned)) * nranks, st>>>(a, nranks, bounds, cursors, sendbuf);
}

__global__ void k_pick_rec0(const int* __restrict__ gathered, int nranks, int rank, int* rec0)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int k = 0; k < 8; k++) rec0[k] = 0;
    for (int r = 0; r < nranks; r++)
        if (gathered[8 * r + 6]) {
            for (int k = 0; k < 6; k++) rec0[k] = gathered[8 * r + k];
            rec0[6] = 1; rec0[7] = r == rank;
            return;
        }
}
void launch_pick_rec0(const int* gathered, int nranks, int rank, int* rec0, cudaStream_t st) { k_pick_rec0<<<1, 32, 0, st>>>(gathered, nranks, rank, rec0); }

struct Extra4 { long long v[4]; };
__global__ void k_pack_counts(const unsigned long long* counts, const unsigned long long* list_n, const int* sym_flag, Extra4 ex, int nranks, long long* out)
{
    const int k = threadIdx.x;
    if (k < nranks) out[k] = (long long)counts[k];
    if (k == 0) { out[nranks] = (long long)*list_n; out[nranks + 1] = *sym_flag; }
    if (k < 4) out[nranks + 2 + k] = ex.v[k];
}
void launch_pack_counts(const unsigned long long* counts, const unsigned long long* list_n, const int* sym_flag, const long long extra[4],
                        int nranks, long long* out, cudaStream_t st)
{
    Extra4 ex{{extra[0], extra[1], extra[2], extra[3]}};
    k_pack_counts<<<1, 64, 0, st>>>(counts, list_n, sym_flag, ex, nranks, out);
}

struct ScalarPtrs { const long long* p[8]; };
__global__ void k_pack_scalars(ScalarPtrs s, int n, const ErrState* err, const ErrState* err_range, long long* out)
{
    const int k = threadIdx.x;
    if (k < n) out[k] = *s.p[k];
    if (k == 0) { out[n] = (long long)err->packed; out[n + 1] = (long long)err_range->packed; }
}
void launch_pack_scalars(const long long* const* src_host, int n, const ErrState* err, const ErrState* err_range, long long* out, cudaStream_t st)
{
    ScalarPtrs s{};
    for (int k = 0; k < n && k < 8; k++) s.p[k] = src_host[k];
    k_pack_scalars<<<1, 32, 0, st>>>(s, n, err, err_range, out);
}

} // namespace raftk
