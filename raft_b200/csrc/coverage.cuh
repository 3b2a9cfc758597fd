// coverage.cuh — interval -> difference-array update shared by the tokenizer (query sides, fused) and K2.
#pragma once
#include "kernels.h"

namespace raftk {

// Adds interval [s,e) of owned local read lr: +1 at its first bin, -1 one past its last bin (repeat.hpp:62-77:
// lo = max(s,0)/reso, bins lo..(e-1)/reso when e-1 >= lo*reso).  Returns false when it would leave the read's bins
// (the reference writes out of bounds there).
__device__ __forceinline__ bool add_interval(int32_t* diff, const int64_t* __restrict__ slot_off, int64_t lr, int s, int e, int reso)
{
    int64_t base = slot_off[lr];
    int64_t nb = slot_off[lr + 1] - base - 1;
    // s, e are 32-bit and reso >= 1: both quotients fit 32-bit unsigned division
    int64_t lo = (int64_t)((unsigned)(s < 0 ? 0 : s) / (unsigned)reso);
    int64_t em = (int64_t)e - 1;
    if (em < lo * reso) return true; // nothing covered (repeat.hpp:69 never true)
    int64_t hi = (int64_t)((unsigned)em / (unsigned)reso); // 0 <= lo*reso <= em < 2^31 here
    if (hi >= nb) return false;
    atomicAdd(diff + base + lo, 1);
    atomicAdd(diff + base + hi + 1, -1);
    return true;
}

} // namespace raftk
