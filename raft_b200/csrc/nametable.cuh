// nametable.cuh — device hash table read-name -> dense id.
// Replaces std::unordered_map<std::string,int> of chop.hpp:73-85 (addStringToMap) / chop.hpp:162-163.
//
// Keys are 64-bit hashes of the full name bytes.  After the build, k_name_verify proves that every
// read probes back to its own id; two reads landing on the same key are either byte-identical
// names (RAFTGPU_E_DUP_NAME) or a true 64-bit collision (the host re-seeds and rebuilds).  With all
// stored keys distinct, every PAF name that IS a read name resolves exactly; a name that is not a
// read name is reported as unknown unless it collides in 64 bits with a stored key (p ~ n/2^64).
#pragma once
#include "common.cuh"

namespace raftk {

struct NameSlot {
    unsigned long long key; // 0 = empty
    int                id;
    int                pad;
};

struct NameTable {
    NameSlot*          slots;
    unsigned long long mask; // capacity - 1 (power of two)
    unsigned long long seed;
};

// Incremental hash: bytes are consumed as little-endian 32-bit words (zero padded), one multiply
// per word; the byte count is folded in at the end.
struct NameHasher {
    unsigned long long h;
    unsigned           w;
    unsigned           n;
    __device__ __forceinline__ void init(unsigned long long seed) { h = seed ^ 0x9E3779B97F4A7C15ull; w = 0; n = 0; }
    __device__ __forceinline__ void add(unsigned c)
    {
        w |= c << ((n & 3u) * 8u);
        n++;
        if ((n & 3u) == 0) { h = (h ^ w) * 0xD6E8FEB86659FD93ull; h ^= h >> 29; w = 0; }
    }
    __device__ __forceinline__ unsigned long long finish()
    {
        if (n & 3u) { h = (h ^ w) * 0xD6E8FEB86659FD93ull; h ^= h >> 29; }
        unsigned long long r = mix64(h ^ ((unsigned long long)n << 32));
        return r ? r : 1ull;
    }
};

__device__ __forceinline__ unsigned long long hash_name_global(const uint8_t* s, int64_t len, unsigned long long seed)
{
    NameHasher hs;
    hs.init(seed);
    for (int64_t i = 0; i < len; i++) hs.add(s[i]);
    return hs.finish();
}

__device__ __forceinline__ int nametable_find(const NameTable& t, unsigned long long key)
{
    unsigned long long i = key & t.mask;
    for (;;) {
        // one 16-byte load per probe
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(t.slots + i));
        unsigned long long k = (unsigned long long)v.x | ((unsigned long long)v.y << 32);
        if (k == key) return (int)v.z;
        if (k == 0ull) return -1;
        i = (i + 1) & t.mask;
    }
}

} // namespace raftk
