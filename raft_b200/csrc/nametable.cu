// nametable.cu — build + verify the device name -> id table (see nametable.cuh).
#include "kernels.h"

namespace raftk {

__global__ void k_name_insert(NameTable t, const uint8_t* names, const int64_t* name_off, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t            a = name_off[i];
    unsigned long long key = hash_name_global(names + a, name_off[i + 1] - a, t.seed);
    unsigned long long s = key & t.mask;
    for (;;) {
        unsigned long long old = atomicCAS(&t.slots[s].key, 0ull, key);
        if (old == 0ull || old == key) { atomicMin(&t.slots[s].id, (int)i); return; } // smallest id wins (first occurrence)
        s = (s + 1) & t.mask;
    }
}

// every read must probe back to itself; otherwise two names share a key
__global__ void k_name_verify(NameTable t, const uint8_t* names, const int64_t* name_off, int64_t n, ErrState* err)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t            a = name_off[i], la = name_off[i + 1] - a;
    unsigned long long key = hash_name_global(names + a, la, t.seed);
    int                j = nametable_find(t, key);
    if (j == (int)i) return;
    bool same = false;
    if (j >= 0) {
        int64_t b = name_off[j], lb = name_off[j + 1] - b;
        same = (la == lb);
        for (int64_t k = 0; same && k < la; k++) same = names[a + k] == names[b + k];
    }
    int       code = same ? RAFTK_E_DUP_NAME : RAFTK_E_HASH_COLLISION;
    err_min(err, code, (long long)i);
}

__global__ void k_name_clear(NameSlot* slots, unsigned long long cap)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) { slots[i].key = 0ull; slots[i].id = 0x7fffffff; slots[i].pad = 0; }
}

cudaError_t launch_name_build(NameTable* desc, void* slots, unsigned long long capacity, unsigned long long seed, const uint8_t* names,
                              const int64_t* name_off, int64_t n, ErrState* err, cudaStream_t st)
{
    desc->slots = reinterpret_cast<NameSlot*>(slots);
    desc->mask = capacity - 1;
    desc->seed = seed;
    k_name_clear<<<(unsigned)((capacity + 255) / 256), 256, 0, st>>>(desc->slots, capacity);
    if (n > 0) {
        unsigned g = (unsigned)((n + 255) / 256);
        k_name_insert<<<g, 256, 0, st>>>(*desc, names, name_off, n);
        k_name_verify<<<g, 256, 0, st>>>(*desc, names, name_off, n, err);
    }
    return cudaGetLastError();
}

} // namespace raftk
