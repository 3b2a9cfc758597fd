// synth_gen.cu — device-side generator of synthetic bench inputs (libraft_synth.so).
// BENCH/TEST INPUT GENERATION ONLY: builds reads + PAF text of the shapes in SURVEY.md §8.C directly
// in HBM so that bench.py can run human-scale inputs without a host-side generator.  Not part of the
// fragmentation path and not part of the C ABI in include/raft_b200.h.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

using namespace raftk;

namespace {

__device__ __forceinline__ uint64_t h2(uint64_t seed, uint64_t stream, uint64_t idx)
{
    return mix64(idx * 0x9E3779B97F4A7C15ull + mix64(seed * 0x9E3779B97F4A7C15ull + stream * 0xD1B54A32D192ED03ull));
}

// UUID-like 36-byte name of read i: 8-4-4-4-12 hex, last 8 hex digits = i (unique)
__device__ __forceinline__ uint8_t name_char(uint64_t seed, int64_t i, int c)
{
    if (c == 8 || c == 13 || c == 18 || c == 23) return '-';
    const char* hexd = "0123456789abcdef";
    if (c >= 28) return hexd[((uint64_t)i >> (4 * (35 - c))) & 15];
    int      k = c - (c > 8) - (c > 13) - (c > 18) - (c > 23);
    uint64_t a = h2(seed, 20 + (k >> 4), (uint64_t)i);
    return hexd[(a >> (4 * (k & 15))) & 15];
}

__global__ void k_names(uint8_t* names, int64_t n, uint64_t seed)
{
    int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n * 36) return;
    names[x] = name_char(seed, x / 36, (int)(x % 36));
}

__device__ __forceinline__ uint8_t genome_base(uint64_t seed, int64_t pos) { return "ACGT"[h2(seed, 1, (uint64_t)pos) & 3]; }

// one thread per 16 output bases
__global__ void k_seq(uint8_t* seq, const int64_t* __restrict__ seq_off, const int64_t* __restrict__ start,
                      const int8_t* __restrict__ strand, int64_t n, int64_t total, uint64_t seed)
{
    int64_t x0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (x0 >= total) return;
    int64_t lo = 0, hi = n;
    while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (seq_off[mid] <= x0) lo = mid; else hi = mid; }
    int64_t  r = lo, rs = seq_off[r], re = seq_off[r + 1];
    uint32_t w[4] = {0, 0, 0, 0};
    int      cnt = 0;
    for (int k = 0; k < 16 && x0 + k < total; k++, cnt++) {
        int64_t x = x0 + k;
        while (x >= re) { r++; rs = re; re = seq_off[r + 1]; }
        int64_t off = x - rs, L = re - rs;
        uint8_t b;
        if (strand[r]) { // reverse complement
            uint8_t g = genome_base(seed, start[r] + (L - 1 - off));
            b = g == 'A' ? 'T' : g == 'C' ? 'G' : g == 'G' ? 'C' : 'A';
        } else b = genome_base(seed, start[r] + off);
        w[k >> 2] |= (uint32_t)b << ((k & 3) * 8);
    }
    if (cnt == 16) *reinterpret_cast<uint4*>(seq + x0) = make_uint4(w[0], w[1], w[2], w[3]);
    else for (int k = 0; k < cnt; k++) seq[x0 + k] = (uint8_t)(w[k >> 2] >> ((k & 3) * 8));
}

// unwrapped FASTA text of reads [r0, r0+n): ">" name(36) "\n" bases "\n"; one thread per 16 output bytes
__global__ void k_fasta_text(uint8_t* text, const int64_t* __restrict__ seq_off, const int64_t* __restrict__ start,
                             const int8_t* __restrict__ strand, int64_t r0, int64_t n, int64_t total, uint64_t seed)
{
    int64_t x0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (x0 >= total) return;
    const int64_t base = seq_off[r0] + 39 * r0;        // text offset of record r0 in the whole file
    auto rec_off = [&](int64_t i) { return seq_off[i] + 39 * i - base; };
    int64_t lo = r0, hi = r0 + n;
    while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (rec_off(mid) <= x0) lo = mid; else hi = mid; }
    int64_t r = lo, ro = rec_off(r), L = seq_off[r + 1] - seq_off[r];
    for (int k = 0; k < 16 && x0 + k < total; k++) {
        int64_t x = x0 + k;
        while (x >= ro + 39 + L) { r++; ro = rec_off(r); L = seq_off[r + 1] - seq_off[r]; }
        int64_t off = x - ro;
        uint8_t b;
        if (off == 0) b = '>';
        else if (off <= 36) b = name_char(seed, r, (int)off - 1);
        else if (off == 37 || off == 38 + L) b = '\n';
        else {
            int64_t q = off - 38;
            if (strand[r]) { uint8_t g = genome_base(seed, start[r] + (L - 1 - q)); b = g == 'A' ? 'T' : g == 'C' ? 'G' : g == 'G' ? 'C' : 'A'; }
            else b = genome_base(seed, start[r] + q);
        }
        text[x] = b;
    }
}

struct PafCols {
    const int64_t *q, *t, *qs, *qe, *ts, *te;
    const int8_t*  rev;
    const int64_t* len; // read lengths
    int64_t        N;
    uint64_t       seed;
};

// qn ql qs qe strand tn tl ts te nmatch blen 255\n  with 36-byte names
__device__ __forceinline__ int paf_line_size(const PafCols& c, int64_t k)
{
    int64_t bl = max(c.qe[k] - c.qs[k], c.te[k] - c.ts[k]);
    return 36 + 36 + dec_digits64(c.len[c.q[k]]) + dec_digits64(c.qs[k]) + dec_digits64(c.qe[k]) + dec_digits64(c.len[c.t[k]]) +
           dec_digits64(c.ts[k]) + dec_digits64(c.te[k]) + 2 * dec_digits64(bl) + 11 + 1 + 3 + 1;
}
__global__ void k_paf_sizes(PafCols c, int32_t* sizes)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < c.N) sizes[k] = paf_line_size(c, k);
}
__device__ __forceinline__ uint8_t* put_u64(uint8_t* p, uint64_t v)
{
    int nd = dec_digits64(v);
    for (int d = nd - 1; d >= 0; d--) { p[d] = (uint8_t)('0' + (unsigned)(v % 10ull)); v /= 10ull; }
    return p + nd;
}
__global__ void k_paf_write(PafCols c, const int64_t* __restrict__ off, uint8_t* text)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= c.N) return;
    uint8_t* p = text + off[k];
    int64_t  q = c.q[k], t = c.t[k], bl = max(c.qe[k] - c.qs[k], c.te[k] - c.ts[k]);
    for (int j = 0; j < 36; j++) *p++ = name_char(c.seed, q, j);
    *p++ = '\t'; p = put_u64(p, c.len[q]); *p++ = '\t'; p = put_u64(p, c.qs[k]); *p++ = '\t'; p = put_u64(p, c.qe[k]);
    *p++ = '\t'; *p++ = c.rev[k] ? '-' : '+'; *p++ = '\t';
    for (int j = 0; j < 36; j++) *p++ = name_char(c.seed, t, j);
    *p++ = '\t'; p = put_u64(p, c.len[t]); *p++ = '\t'; p = put_u64(p, c.ts[k]); *p++ = '\t'; p = put_u64(p, c.te[k]);
    *p++ = '\t'; p = put_u64(p, bl); *p++ = '\t'; p = put_u64(p, bl); *p++ = '\t'; *p++ = '2'; *p++ = '5'; *p++ = '5'; *p++ = '\n';
}

} // namespace

extern "C" {
int synth_names(void* names, int64_t n, uint64_t seed, void* stream)
{
    if (n > 0) k_names<<<(unsigned)((n * 36 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((uint8_t*)names, n, seed);
    return (int)cudaGetLastError();
}
int synth_seq(void* seq, const void* seq_off, const void* start, const void* strand, int64_t n, int64_t total, uint64_t seed, void* stream)
{
    int64_t thr = (total + 15) / 16;
    if (thr > 0)
        k_seq<<<(unsigned)((thr + 255) / 256), 256, 0, (cudaStream_t)stream>>>((uint8_t*)seq, (const int64_t*)seq_off, (const int64_t*)start,
                                                                                   (const int8_t*)strand, n, total, seed);
    return (int)cudaGetLastError();
}
int synth_fasta_text(void* text, const void* seq_off, const void* start, const void* strand, int64_t r0, int64_t n, int64_t total, uint64_t seed,
                     void* stream)
{
    int64_t thr = (total + 15) / 16;
    if (thr > 0)
        k_fasta_text<<<(unsigned)((thr + 255) / 256), 256, 0, (cudaStream_t)stream>>>((uint8_t*)text, (const int64_t*)seq_off, (const int64_t*)start,
                                                                                          (const int8_t*)strand, r0, n, total, seed);
    return (int)cudaGetLastError();
}
int synth_paf_sizes(const void* q, const void* t, const void* qs, const void* qe, const void* ts, const void* te, const void* rev,
                    const void* len, int64_t N, uint64_t seed, void* sizes, void* stream)
{
    PafCols c{(const int64_t*)q, (const int64_t*)t, (const int64_t*)qs, (const int64_t*)qe, (const int64_t*)ts, (const int64_t*)te,
              (const int8_t*)rev, (const int64_t*)len, N, seed};
    if (N > 0) k_paf_sizes<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c, (int32_t*)sizes);
    return (int)cudaGetLastError();
}
int synth_paf_write(const void* q, const void* t, const void* qs, const void* qe, const void* ts, const void* te, const void* rev,
                    const void* len, int64_t N, uint64_t seed, const void* off, void* text, void* stream)
{
    PafCols c{(const int64_t*)q, (const int64_t*)t, (const int64_t*)qs, (const int64_t*)qe, (const int64_t*)ts, (const int64_t*)te,
              (const int8_t*)rev, (const int64_t*)len, N, seed};
    if (N > 0) k_paf_write<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c, (const int64_t*)off, (uint8_t*)text);
    return (int)cudaGetLastError();
}
}
