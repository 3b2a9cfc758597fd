// covtext.cuh — slot walk and per-slot digit counts of coverage.txt (emitter K5a).
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


} // namespace raftk
