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

// Incremental hash: bytes are consumed as little-endian 32-bit words (zero padded) by two independent
// 32-bit multiplicative streams (3 integer instructions per word); the byte count is folded in at
// the end and the 64 bits are finalised with mix64.  hash_name_smem in k1_paf.cu computes the same
// value word-wise from shared memory.
constexpr unsigned NH_K1 = 0x9E3779B1u, NH_K2 = 0x85EBCA77u;
__device__ __forceinline__ unsigned long long name_hash_finish(unsigned a, unsigned b, unsigned n)
{
    unsigned long long r = mix64((((unsigned long long)a << 32) | b) ^ ((unsigned long long)n * 0x9E3779B97F4A7C15ull));
    return r ? r : 1ull;
}
struct NameHasher {
    unsigned a, b, w, n;
    __device__ __forceinline__ void init(unsigned long long seed) { a = (unsigned)seed ^ 0x9E3779B9u; b = (unsigned)(seed >> 32) ^ 0x85EBCA6Bu; w = 0; n = 0; }
    __device__ __forceinline__ void word(unsigned x) { a = (a ^ x) * NH_K1; b = b * NH_K2 + x; }
    __device__ __forceinline__ void add(unsigned c)
    {
        w |= c << ((n & 3u) * 8u);
        n++;
        if ((n & 3u) == 0) { word(w); w = 0; }
    }
    __device__ __forceinline__ unsigned long long finish()
    {
        if (n & 3u) word(w);
        return name_hash_finish(a, b, n);
    }
};

// Same value as NameHasher fed byte by byte, from global memory four bytes at a time: aligned 32-bit loads of the words that
// hold the name (never a word past its last byte) and a funnel shift for names that do not start on a word boundary.
__device__ __forceinline__ unsigned long long hash_name_global(const uint8_t* s, int64_t len, unsigned long long seed)
{
    NameHasher hs;
    hs.init(seed);
    const unsigned  mis = (unsigned)((uintptr_t)s & 3u), sh = mis * 8u;
    const unsigned* q = reinterpret_cast<const unsigned*>(s - mis);
    const int64_t   nw = ((int64_t)mis + len + 3) >> 2; // aligned words covering [s, s + len)
    unsigned        lo = nw > 0 ? q[0] : 0u;
    int64_t         k = 0, wi = 0;
    for (; k + 4 <= len; k += 4) {
        const unsigned hi = wi + 1 < nw ? q[wi + 1] : 0u;
        hs.word(__funnelshift_r(lo, hi, sh));
        lo = hi; wi++;
    }
    const int rem = (int)(len - k);
    if (rem) {
        const unsigned hi = wi + 1 < nw ? q[wi + 1] : 0u;
        hs.word(__funnelshift_r(lo, hi, sh) & ((1u << (8 * rem)) - 1u));
    }
    return name_hash_finish(hs.a, hs.b, (unsigned)len);
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
