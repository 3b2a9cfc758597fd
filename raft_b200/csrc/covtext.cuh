// covtext.cuh — slot walk and text sizing of coverage.txt, shared by the scan (fused sizing) and the emitter.
// coverage.txt (repeat.hpp:105-108): "read " i " " then k*reso "," cov " " per bin, then "\n".
#pragma once
#include "kernels.h"

namespace raftk {

// Per-thread walk over a few consecutive slots.  A slot is one bin ("pos,cov ") or the read's sentinel ("\n");
// the first slot of a read is preceded by "read i ".  The walk keeps (bin, slots left in the read) in 32-bit
// registers and only touches slot_off when it crosses into the next read.
struct SlotWalk {
    const int64_t* __restrict__ slot_off;
    int64_t r;     // current read (local index)
    int64_t re;    // one past its last slot
    int     bin;   // index of the current slot inside the read
    int     left;  // slots left in the read including the current one (clamped)
    __device__ __forceinline__ void init(const int64_t* so, int64_t read, int64_t g)
    {
        slot_off = so; r = read;
        const int64_t rs = so[read];
        re = so[read + 1];
        bin = (int)(g - rs);
        const int64_t rem = re - g;
        left = rem > (1 << 30) ? (1 << 30) : (int)rem;
    }
    __device__ __forceinline__ void next()
    {
        bin++;
        if (--left == 0) {
            r++;
            const int64_t nre = slot_off[r + 1];
            const int64_t rem = nre - re;
            left = rem > (1 << 30) ? (1 << 30) : (int)rem;
            re = nre; bin = 0;
        }
    }
};

// digit counts of one slot packed as pos | cov << 4 | neg << 8 (cov may be negative only on invalid input)
__device__ __forceinline__ int slot_digits(int bin, int reso, int cov)
{
    const unsigned ucov = cov < 0 ? (unsigned)(-(int64_t)cov) : (unsigned)cov;
    return dec_digits((uint32_t)bin * (uint32_t)reso) | (dec_digits(ucov) << 4) | ((cov < 0) << 8);
}


// bytes of coverage.txt contributed by `cnt` (<= 4) consecutive slots starting at global slot g with coverages cv[]
__device__ __forceinline__ int cov_text_size4(const int64_t* __restrict__ slot_off, const int32_t* __restrict__ tile_read, int64_t g, int cnt,
                                              const int* cv, int reso, int64_t own_first)
{
    const int64_t ct = g / COV_TILE_SLOTS;
    int64_t       lo = tile_read[ct], hi = (int64_t)tile_read[ct + 1] + 1;
    while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (slot_off[mid] <= g) lo = mid; else hi = mid; }
    SlotWalk w;
    w.init(slot_off, lo, g);
    int sz = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (k < cnt) {
            if (w.bin == 0) sz += 5 + dec_digits64((uint64_t)(own_first + w.r)) + 1;
            if (w.left == 1) sz += 1;
            else { int d = slot_digits(w.bin, reso, cv[k]); sz += (d & 15) + ((d >> 4) & 15) + (d >> 8) + 2; }
            w.next();
        }
    }
    return sz;
}

} // namespace raftk
