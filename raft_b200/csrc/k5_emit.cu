// k5_emit.cu — K5: text and sequence emitters.  Each output file of the reference is a byte
// stream of known length that is materialised on the device one window [w0, w1) at a time:
//
//  K5a coverage.txt   "read " i " " then k*reso "," cov " " per bin, then "\n"   (repeat.hpp:105-108)
//  K5c long_repeats   "read " i ", " then s "," e "    " per repeat, then "\n"   (repeat.hpp:180-203)
//  K5b reads.fasta    ">read=" num "," name ",pos_on_original_read=" a "-" b "\n" bases[a:b] "\n"
//                                                                                (chop.hpp:261-265,314-318)
//
// K5a is slot-parallel (a slot is one bin, or the per-read sentinel that carries the newline):
// one warp per tile of 1024 slots formats 32 slots at a time into shared memory at the same 16-byte
// phase as its destination and stores them with aligned 128-bit writes.  K5b is output-tile-parallel: every CTA
// owns 16 KiB of the output file, finds the fragments that intersect it by binary search over the
// record offsets, generates header bytes on the fly and gathers sequence bytes with aligned
// 128-bit loads + funnel shifts + aligned 128-bit stores (stream-compacted: bytes land in their
// final file order).
#include <cstdlib>

#include "kernels.h"

namespace raftk {

// ================================================================ K5a coverage.txt
// One warp per 1024-slot tile (a slot is one bin, or the per-read sentinel that carries the newline); the tile's
// stream offset comes from the scan of the per-tile text sizes.  The warp walks its tile in 8 chunks of 128 slots,
// four consecutive slots per lane: one 128-bit coverage load per lane (the next chunk's load is already in
// flight), per-lane text size, warp scan, SWAR decimal conversion (four digits per multiply chain) into a per-warp
// shared buffer laid out at the 16-byte phase of its destination, then aligned 128-bit stores.  No block-level
// barrier and no per-thread search: the read a slot belongs to is warp-uniform state that only moves when a chunk
// reaches the end of the current read.
constexpr int CW_WARPS = 8;
constexpr int CW_THREADS = CW_WARPS * 32;
constexpr int CW_PER = 4;             // slots per lane per chunk (8 doubles the time: the registers spill)
constexpr int CW_CHUNK = 32 * CW_PER;
constexpr int CW_BUF = 512 * CW_PER;  // per-warp text buffer
constexpr int CW_CAP = CW_BUF - 16;   // chunks with more text (tiny reads: many "read i " prefixes) take the direct path
constexpr int CW_MAX_SLOT_BYTES = 40; // "read 2147483647 " (16) + "2147483647,-2147483648 " (23)
// Fast lanes (all four slots are ordinary bins of one read): the text of a bin is   <k*reso> ',' <cov> ' '   and both halves
// come from tables -- the position text of bin k does not depend on the data (one 8-byte entry per k: the digits and the
// comma in the low bytes, their count in the top byte; built once per context), the coverage text of values below 1000
// sits in shared memory.  A slot's <= 11 bytes are composed in three registers, shifted to the lane's running byte phase
// and stored as whole 32-bit words; only the lane's first and last partial word go out byte by byte.  Lanes that hold a
// read boundary (a sentinel, a "read i " prefix), a coverage outside [0, 1000) or a bin beyond the table take the generic
// digit-by-digit path below.
constexpr int CT_COV = 1024;          // coverage-text table entries (values 0..999 are used)

// ASCII of the four decimal digits of n < 10000, most significant digit in the low byte
__device__ __forceinline__ unsigned ascii4(unsigned n)
{
    const unsigned q = (n * 5243u) >> 19;                 // n / 100
    const unsigned x = q + ((n - q * 100u) << 16);        // two two-digit lanes
    const unsigned t = ((x * 103u) >> 10) & 0x000F000Fu;  // tens of each lane
    return (t + ((x - t * 10u) << 8)) | 0x30303030u;
}
// the nd decimal digits of v at p; every branch is on nd, which is nearly warp-uniform for bin positions
__device__ __forceinline__ void put_dec(uint8_t* p, unsigned v, int nd)
{
    if (nd <= 2) { // v < 100
        const unsigned t = (v * 103u) >> 10, o = v - t * 10u;
        if (nd == 2) { p[0] = (uint8_t)('0' + t); p[1] = (uint8_t)('0' + o); }
        else p[0] = (uint8_t)('0' + o);
    } else if (nd <= 4) {
        const unsigned A = ascii4(v) >> (nd == 3 ? 8 : 0);
        p[0] = (uint8_t)A; p[1] = (uint8_t)(A >> 8); p[2] = (uint8_t)(A >> 16);
        if (nd == 4) p[3] = (uint8_t)(A >> 24);
    } else if (nd <= 8) {
        const unsigned hi = v / 10000u, A = ascii4(hi), B = ascii4(v - hi * 10000u);
        const unsigned sh = 8u * (8u - (unsigned)nd);     // 0, 8, 16, 24: drop the leading zeros of the 8-digit field
        const unsigned lo = __funnelshift_r(A, B, sh), up = B >> sh;
        p[0] = (uint8_t)lo; p[1] = (uint8_t)(lo >> 8); p[2] = (uint8_t)(lo >> 16); p[3] = (uint8_t)(lo >> 24);
        p[4] = (uint8_t)up;
        if (nd > 5) p[5] = (uint8_t)(up >> 8);
        if (nd > 6) p[6] = (uint8_t)(up >> 16);
        if (nd > 7) p[7] = (uint8_t)(up >> 24);
    } else {
        put_u32_nd(p, v, nd);
    }
}
// text of one slot at p (shared or local memory after inlining)
__device__ __forceinline__ uint8_t* cov_slot_format(uint8_t* p, bool first, bool sentinel, bool neg, unsigned pos, int dpos, unsigned ucov, int dcov,
                                                    unsigned read_id)
{
    if (first) {
        p[0] = 'r'; p[1] = 'e'; p[2] = 'a'; p[3] = 'd'; p[4] = ' '; p += 5;
        const int nd = dec_digits(read_id); // read ids are ints in the reference (repeat.hpp:105)
        put_dec(p, read_id, nd);
        p += nd; *p++ = ' ';
    }
    if (sentinel) { *p++ = '\n'; return p; }
    put_dec(p, pos, dpos); p[dpos] = ','; p += dpos + 1;
    if (neg) *p++ = '-';
    put_dec(p, ucov, dcov); p[dcov] = ' ';
    return p + dcov + 1;
}
// rare (dozens of tiny reads in one chunk): format privately, store byte-wise with clipping
__device__ __noinline__ int cov_slot_direct(uint8_t* gbase, int x, int wlo, int whi, bool first, bool sentinel, bool neg, unsigned pos, unsigned ucov,
                                            unsigned read_id)
{
    uint8_t  loc[CW_MAX_SLOT_BYTES];
    uint8_t* e = cov_slot_format(loc, first, sentinel, neg, pos, dec_digits(pos), ucov, dec_digits(ucov), read_id);
    for (uint8_t* q = loc; q < e; q++, x++) if (x >= wlo && x < whi) gbase[x] = *q;
    return x;
}

__global__ void __launch_bounds__(CW_THREADS, 4) k_cov_text(CovEmitArgs a, int64_t n_tiles)
{
    __shared__ __align__(16) uint8_t sbuf[CW_WARPS][CW_BUF];
    __shared__ unsigned s_cov[CT_COV];
    for (int i = threadIdx.x; i < CT_COV; i += CW_THREADS) s_cov[i] = a.cov_tab[i];
    __syncthreads();
    const int     lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tl = (int64_t)blockIdx.x * CW_WARPS + warp;
    if (tl >= n_tiles) return;
    const int64_t tile = a.tile_first + tl;
    const int64_t oc0 = a.tile_off[tile];
    if (a.tile_off[tile + 1] <= a.w0 || oc0 >= a.w1) return;
    const int64_t g0 = tile * COV_TILE_SLOTS;
    // read holding the tile's first slot: the tile map bounds the search to the reads that intersect the tile
    int64_t r;
    {
        int64_t lo = a.tile_read[tile], hi = (int64_t)a.tile_read[tile + 1] + 1;
        while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (a.slot_off[mid] <= g0) lo = mid; else hi = mid; }
        r = lo;
    }
    // everything inside the tile is 32-bit and relative: slots to g0, text bytes to oc0
    constexpr int FAR = 1 << 20;
    auto rel = [&](int64_t slot) { const int64_t d = slot - g0; return d > FAR ? FAR : (int)d; }; // a read has < 2^31 slots
    // slot_off carries one entry past the end (2^62), so stepping past the last read's sentinel needs no bounds test
    int            rs = rel(a.slot_off[r]), re = rel(a.slot_off[r + 1]), mr = 0; // warp-uniform: read of the chunk's first slot (= r + mr)
    const int      nvalid = a.n_slots - g0 < COV_TILE_SLOTS ? (int)(a.n_slots - g0) : COV_TILE_SLOTS;
    const int64_t  dlo = a.w0 - oc0, dhi = a.w1 - oc0;
    const int      wlo = dlo < 0 ? 0 : (dlo > (1 << 30) ? (1 << 30) : (int)dlo), whi = dhi > (1 << 30) ? (1 << 30) : (int)dhi;
    uint8_t* const gbase = a.dst + (oc0 - a.w0); // byte 0 of the tile's text (only dereferenced inside the window)
    const int      gl = (int)((uintptr_t)gbase & 15);
    const int32_t* covp = a.cov + g0;
    uint8_t*       wb = sbuf[warp];
    const unsigned reso = (unsigned)a.reso, rid0 = (unsigned)(a.own_first + r); // global id of read r
    auto load4 = [&](int s) { // coverage of slots s .. s+3 of the tile
        if (s + 3 < nvalid) return *reinterpret_cast<const int4*>(covp + s);
        int4 v;
        v.x = s < nvalid ? covp[s] : 0; v.y = s + 1 < nvalid ? covp[s + 1] : 0; v.z = s + 2 < nvalid ? covp[s + 2] : 0; v.w = 0;
        return v;
    };
    int  ro = 0; // text offset of the chunk inside the tile
    constexpr int NQ = CW_PER / 4;
    int4 cvn[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) cvn[q] = load4(CW_PER * lane + 4 * q);
    for (int sl0 = 0; sl0 < nvalid; sl0 += CW_CHUNK) {
        const int  s0 = sl0 + CW_PER * lane;
        int cv[CW_PER];
#pragma unroll
        for (int q = 0; q < NQ; q++) { cv[4 * q] = cvn[q].x; cv[4 * q + 1] = cvn[q].y; cv[4 * q + 2] = cvn[q].z; cv[4 * q + 3] = cvn[q].w; }
        if (sl0 + CW_CHUNK < nvalid) {
#pragma unroll
            for (int q = 0; q < NQ; q++) cvn[q] = load4(s0 + CW_CHUNK + 4 * q);
        }
        // read of this lane's first slot
        int        lrs = rs, lre = re, lmr = mr;
        const bool crossing = sl0 + CW_CHUNK - 1 >= re - 1; // the chunk reaches this read's sentinel
        if (crossing) while (s0 < nvalid && s0 >= lre) { lmr++; lrs = lre; lre = rel(a.slot_off[r + lmr + 1]); }
        const int frs = lrs, fre = lre, fmr = lmr; // the format pass restarts here
        // pass 1: sizes.  Three kinds of lanes:
        //  plain    -- four ordinary bins of one read (not its first slot, none its sentinel), all inside the tables (position
        //              < 10^6, coverage in [0, 1000)): the specialised table path;
        //  boundary -- all four slots exist and are inside the tables, but some are sentinels ("\n") or open a read
        //              ("read <id> " in front): the same table path through a general append;
        //  generic  -- everything else (table misses, the ragged end of the last tile): digit by digit.
        const int          k0 = s0 - lrs;
        const bool         plain = s0 + CW_PER - 1 < nvalid && k0 > 0 && s0 + CW_PER - 1 < lre - 1 && k0 + CW_PER - 1 < a.tab_n;
        bool               fast = s0 + CW_PER - 1 < nvalid;
        unsigned long long ek[CW_PER];
        int                size = 0, meta[CW_PER];
        if (plain) {
#pragma unroll
            for (int k = 0; k < CW_PER; k++) {
                ek[k] = __ldg(a.pos_tab + k0 + k);
                const unsigned c = (unsigned)cv[k];
                fast = fast && (ek[k] >> 56) != 0ull && c < 1000u;
                size += (int)(ek[k] >> 56) + 2 + (int)(c > 9u) + (int)(c > 99u);
            }
        } else if (fast) {
#pragma unroll
            for (int k = 0; k < CW_PER; k++) {
                const int  s = s0 + k;
                const bool sentinel = s == lre - 1;
                if (s == lrs) { const unsigned id = rid0 + (unsigned)lmr; fast = fast && id < 100000000u; size += 6 + dec_digits(id); }
                if (!sentinel) {
                    const int      kk = s - lrs;
                    const unsigned c = (unsigned)cv[k];
                    ek[k] = kk < a.tab_n ? __ldg(a.pos_tab + kk) : 0ull;
                    fast = fast && (ek[k] >> 56) != 0ull && c < 1000u;
                    size += (int)(ek[k] >> 56) + 2 + (int)(c > 9u) + (int)(c > 99u);
                } else {
                    ek[k] = 0ull;
                    size += 1;
                    lmr++; lrs = lre; lre = rel(a.slot_off[r + lmr + 1]); // the next slot starts the next read
                }
            }
        }
        if (!fast) {
            size = 0; lrs = frs; lre = fre; lmr = fmr;
#pragma unroll
            for (int k = 0; k < CW_PER; k++) {
                const int      s = s0 + k;
                const bool     sentinel = s == lre - 1, neg = cv[k] < 0;
                const unsigned ucov = neg ? (unsigned)(-(int64_t)cv[k]) : (unsigned)cv[k];
                const int      dpos = dec_digits((unsigned)(s - lrs) * reso), dcov = ucov < 100u ? 1 + (int)(ucov > 9u) : dec_digits(ucov);
                meta[k] = dpos | (dcov << 4);
                if (s < nvalid) {
                    size += sentinel ? 1 : dpos + dcov + 2 + (int)neg;
                    if (s == lrs) size += 6 + dec_digits(rid0 + (unsigned)lmr);
                    if (sentinel) { lmr++; lrs = lre; lre = rel(a.slot_off[r + lmr + 1]); } // the next slot starts the next read
                }
            }
        }
        const int incl = warp_inclusive_sum(size), ex = incl - size;
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        if (crossing) { mr = __shfl_sync(0xffffffffu, lmr, 31); rs = __shfl_sync(0xffffffffu, lrs, 31); re = __shfl_sync(0xffffffffu, lre, 31); }
        const int c0 = ro > wlo ? ro : wlo, c1 = ro + tot < whi ? ro + tot : whi;
        if (c0 < c1) {
            lrs = frs; lre = fre; lmr = fmr;
            if (tot <= a.text_cap) {
                const int phase = (gl + ro) & 15;
                if (fast && plain) {
                    const int o = phase + ex;                    // the lane's text starts at byte o of the warp buffer
                    unsigned* wp = reinterpret_cast<unsigned*>(wb + (o & ~3));
                    int       fill = o & 3;                      // bytes of the first word that belong to the previous lane
                    unsigned  acc = 0;
#pragma unroll
                    for (int k = 0; k < CW_PER; k++) {
                        const unsigned long long e = ek[k];
                        const unsigned c = (unsigned)cv[k], ct = s_cov[c];      // digits + ' ' (2..4 bytes)
                        const int      aa = (int)(e >> 56), bb = 2 + (int)(c > 9u) + (int)(c > 99u);
                        const unsigned lo = (unsigned)e, hi = (unsigned)(e >> 32) & 0x00FFFFFFu, sh = (unsigned)(aa & 3) * 8u;
                        const unsigned c_lo = ct << sh, c_hi = __funnelshift_l(ct, 0u, sh);
                        unsigned       W0, W1, W2;                              // the slot's text, first byte in the lowest byte of W0
                        if (aa < 4) { W0 = lo | c_lo; W1 = c_hi; W2 = 0u; } else { W0 = lo; W1 = hi | c_lo; W2 = c_hi; }
                        const unsigned s8 = (unsigned)fill * 8u;
                        const unsigned X0 = acc | (W0 << s8), X1 = __funnelshift_l(W0, W1, s8), X2 = __funnelshift_l(W1, W2, s8),
                                       X3 = __funnelshift_l(W2, 0u, s8);
                        const int      total = fill + aa + bb, nfull = total >> 2;
                        if (k == 0) { // a bin has >= 4 bytes: the first word always completes; its leading `fill` bytes are not this lane's
                            uint8_t* bp = reinterpret_cast<uint8_t*>(wp);
                            if (fill == 0) wp[0] = X0;
                            else { if (fill <= 1) bp[1] = (uint8_t)(X0 >> 8); if (fill <= 2) bp[2] = (uint8_t)(X0 >> 16); bp[3] = (uint8_t)(X0 >> 24); }
                        } else if (nfull > 0) wp[0] = X0;
                        if (nfull > 1) wp[1] = X1;
                        if (nfull > 2) wp[2] = X2;
                        wp += nfull;
                        acc = (nfull & 2) ? ((nfull & 1) ? X3 : X2) : ((nfull & 1) ? X1 : X0);
                        fill = total & 3;
                    }
                    uint8_t* bp = reinterpret_cast<uint8_t*>(wp); // the last partial word shares its other bytes with the next lane
                    if (fill > 0) bp[0] = (uint8_t)acc;
                    if (fill > 1) bp[1] = (uint8_t)(acc >> 8);
                    if (fill > 2) bp[2] = (uint8_t)(acc >> 16);
                } else if (fast) {
                    const int o = phase + ex;                    // the lane's text starts at byte o of the warp buffer
                    unsigned* wp = reinterpret_cast<unsigned*>(wb + (o & ~3));
                    const int fill0 = o & 3;                     // bytes of the first word that belong to the previous lane
                    int       fill = fill0;
                    unsigned  acc = 0;
                    bool      head = true;                       // the lane's first word has not gone out yet
                    // appends L <= 12 bytes (first byte in the lowest byte of W0): whole words go out as words, the rest waits in acc
                    auto append = [&](unsigned W0, unsigned W1, unsigned W2, int L) {
                        const unsigned s8 = (unsigned)fill * 8u;
                        const unsigned X0 = acc | (W0 << s8), X1 = __funnelshift_l(W0, W1, s8), X2 = __funnelshift_l(W1, W2, s8),
                                       X3 = __funnelshift_l(W2, 0u, s8);
                        const int      total = fill + L, nfull = total >> 2;
                        if (nfull > 0) {
                            if (head && fill0) { // only the bytes from fill0 on are this lane's
                                uint8_t* bp = reinterpret_cast<uint8_t*>(wp);
                                if (fill0 <= 1) bp[1] = (uint8_t)(X0 >> 8);
                                if (fill0 <= 2) bp[2] = (uint8_t)(X0 >> 16);
                                bp[3] = (uint8_t)(X0 >> 24);
                            } else wp[0] = X0;
                            head = false;
                        }
                        if (nfull > 1) wp[1] = X1;
                        if (nfull > 2) wp[2] = X2;
                        wp += nfull;
                        acc = (nfull & 2) ? ((nfull & 1) ? X3 : X2) : ((nfull & 1) ? X1 : X0);
                        fill = total & 3;
                    };
#pragma unroll
                    for (int k = 0; k < CW_PER; k++) {
                        const int  s = s0 + k;
                        const bool sentinel = s == lre - 1;
                        if (s == lrs) { // "read <id> ": the five letters, then the id's digits and a blank (id < 10^8 on this path)
                            const unsigned id = rid0 + (unsigned)lmr;
                            const int      nd = dec_digits(id);
                            const unsigned hi4 = id / 10000u, A = ascii4(hi4), B = ascii4(id - hi4 * 10000u); // 8 digits, leading zeros included
                            const unsigned sh = 8u * (unsigned)((8 - nd) & 3);
                            unsigned       D0, D1;                                                       // the nd digits, first in the lowest byte
                            if (nd > 4) { D0 = __funnelshift_r(A, B, sh); D1 = B >> sh; } else { D0 = B >> sh; D1 = 0u; }
                            // a blank after the last digit
                            if (nd < 4) D0 |= 0x20u << (8 * nd); else if (nd == 4) D1 = 0x20u; else if (nd < 8) D1 |= 0x20u << (8 * (nd - 4));
                            append(0x64616572u, 0x20u, 0u, 5);
                            append(D0, D1, nd == 8 ? 0x20u : 0u, nd + 1);
                        }
                        if (!sentinel) {
                            const unsigned long long e = ek[k];
                            const unsigned c = (unsigned)cv[k], ct = s_cov[c];      // digits + ' ' (2..4 bytes)
                            const int      aa = (int)(e >> 56), bb = 2 + (int)(c > 9u) + (int)(c > 99u);
                            const unsigned lo = (unsigned)e, hi = (unsigned)(e >> 32) & 0x00FFFFFFu, sh = (unsigned)(aa & 3) * 8u;
                            const unsigned c_lo = ct << sh, c_hi = __funnelshift_l(ct, 0u, sh);
                            if (aa < 4) append(lo | c_lo, c_hi, 0u, aa + bb); else append(lo, hi | c_lo, c_hi, aa + bb);
                        } else {
                            append(0x0Au, 0u, 0u, 1);
                            lmr++; lrs = lre; lre = rel(a.slot_off[r + lmr + 1]);
                        }
                    }
                    uint8_t*  bp = reinterpret_cast<uint8_t*>(wp); // the last partial word shares its other bytes with the next lane
                    const int from = head ? fill0 : 0;
                    if (fill > 0 && from <= 0) bp[0] = (uint8_t)acc;
                    if (fill > 1 && from <= 1) bp[1] = (uint8_t)(acc >> 8);
                    if (fill > 2 && from <= 2) bp[2] = (uint8_t)(acc >> 16);
                } else {
                    uint8_t* p = wb + phase + ex;
#pragma unroll
                    for (int k = 0; k < CW_PER; k++) {
                        const int s = s0 + k;
                        if (s < nvalid) {
                            const bool sentinel = s == lre - 1, neg = cv[k] < 0;
                            p = cov_slot_format(p, s == lrs, sentinel, neg, (unsigned)(s - lrs) * reso, meta[k] & 15,
                                                neg ? (unsigned)(-(int64_t)cv[k]) : (unsigned)cv[k], meta[k] >> 4, rid0 + (unsigned)lmr);
                            if (sentinel) { lmr++; lrs = lre; lre = rel(a.slot_off[r + lmr + 1]); }
                        }
                    }
                }
                __syncwarp();
                int fa = c0 + ((16 - ((gl + c0) & 15)) & 15), la = c1 - ((gl + c1) & 15); // 16-byte aligned part of [c0, c1)
                if (fa > la) { fa = c1; la = c1; }
                const uint8_t* sb = wb + (phase - ro); // sb[x] is byte x of the tile's text
                if (c0 + lane < fa) gbase[c0 + lane] = sb[c0 + lane];
#pragma unroll 1
                for (int x = fa + lane * 16; x < la; x += 512) stg_stream(reinterpret_cast<uint4*>(gbase + x), *reinterpret_cast<const uint4*>(sb + x));
                if (la + lane < c1) gbase[la + lane] = sb[la + lane];
                __syncwarp(); // the buffer is rewritten by the next chunk
            } else {
                int x = ro + ex;
                for (int k = 0; k < CW_PER; k++) {
                    const int s = s0 + k;
                    if (s < nvalid) {
                        const bool sentinel = s == lre - 1, neg = cv[k] < 0;
                        x = cov_slot_direct(gbase, x, wlo, whi, s == lrs, sentinel, neg, (unsigned)(s - lrs) * reso,
                                            neg ? (unsigned)(-(int64_t)cv[k]) : (unsigned)cv[k], rid0 + (unsigned)lmr);
                        if (sentinel) { lmr++; lrs = lre; lre = rel(a.slot_off[r + lmr + 1]); }
                    }
                }
            }
        }
        ro += tot;
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

// position text of bin k ("<k*reso>,": at most 6 digits + comma in the low 7 bytes, their count in the top byte; 0 = not
// representable, take the generic path) and coverage text ("<c> " for c < 1000, first digit in the lowest byte)
__global__ void __launch_bounds__(256) k_cov_tables(unsigned long long* pos_tab, int n, unsigned reso, unsigned* cov_tab)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        const unsigned long long pos = (unsigned long long)k * reso;
        unsigned long long       e = 0;
        if (pos < 1000000ull) {
            unsigned  v = (unsigned)pos;
            const int nd = dec_digits(v);
            for (int i = nd - 1; i >= 0; i--) { e |= (unsigned long long)('0' + v % 10u) << (8 * i); v /= 10u; }
            e |= (unsigned long long)',' << (8 * nd);
            e |= (unsigned long long)(nd + 1) << 56;
        }
        pos_tab[k] = e;
    }
    if (k < CT_COV) {
        unsigned  v = (unsigned)k, t = 0;
        const int nd = dec_digits(v);
        if (nd <= 3) {
            for (int i = nd - 1; i >= 0; i--) { t |= ('0' + v % 10u) << (8 * i); v /= 10u; }
            t |= (unsigned)' ' << (8 * nd);
        }
        cov_tab[k] = t;
    }
}
void launch_cov_tables(unsigned long long* pos_tab, int n, int reso, unsigned* cov_tab, cudaStream_t st)
{
    const int cnt = n > CT_COV ? n : CT_COV;
    k_cov_tables<<<(cnt + 255) / 256, 256, 0, st>>>(pos_tab, n, (unsigned)reso, cov_tab);
}

int  cov_tiles(int64_t n_slots) { return (int)((n_slots + COV_TILE_SLOTS - 1) / COV_TILE_SLOTS); }
void launch_cov_emit(const CovEmitArgs& a_in, int64_t n_tiles_launch, cudaStream_t st)
{
    if (n_tiles_launch <= 0) return;
    CovEmitArgs a = a_in;
    a.text_cap = CW_CAP;
    if (const char* e = getenv("RAFT_B200_COV_CAP")) { int v = atoi(e); if (v >= 0 && v < CW_CAP) a.text_cap = v; } // test knob: force the direct path
    k_cov_text<<<(unsigned)((n_tiles_launch + CW_WARPS - 1) / CW_WARPS), CW_THREADS, 0, st>>>(a, n_tiles_launch);
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
// 128-bit streaming stores, and release the slot through its `empty` mbarrier.  Two slots (one tile loading
// while one is stored, four CTAs per SM) measured faster than three: 33.0 vs 35.2 ms per pass at full scale -- the
// extra 74 KiB per SM serve as L1 for the descriptor, name and header traffic.
constexpr int FE_THREADS = 256;
constexpr int FE_CONSUMERS = FE_THREADS - 32;       // 7 warps
constexpr int FE_SLOTS = 2;
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
