// k3_repeat_cut.cu — K3: repeat detection and cut points, one warp per read.
//
// Replaces the run-length state machine of repeat_annotate (repeat.hpp:111-170) and the star /
// fragment arithmetic of break_reads (chop.hpp:209-320).
//
//  * repeats: every maximal run of bins with coverage >= H whose length (k2-k1+1)*reso >= p gives
//    (max(0, k1*reso - f), min(L, (k2+1)*reso + f)) (repeat.hpp:119-168); runs are found from warp
//    ballots of (cov >= H), 32 bins at a time, with the open run carried between ballots.
//  * stars: 0, P, 2P, ..., floor(L/P)*P and L when L % P != 0 (chop.hpp:209-223); an interior star
//    inside any repeat [s,e] (both ends inclusive) is dropped, first and last always stay
//    (chop.hpp:225-246).
//  * fragments: div = l / P (chop.hpp:248); with n surviving stars there are
//    F = max(1, ceil((n-1)/div)) fragments and the boundaries are the surviving stars whose rank is
//    a positive multiple of div, except the last star (chop.hpp:250-291).  Only those boundaries
//    ("cuts") are stored; k_frag_expand turns them into (a,b) with the -v back-overlap.
#include "kernels.h"

namespace raftk {

constexpr int K3_THREADS = 256;
constexpr int K3_CHUNK = 8; // reads claimed per atomic

__global__ void __launch_bounds__(K3_THREADS) k_repeat_cut(RepeatCutArgs a)
{
    const int          lane = lane_id();
    const unsigned     lt_mask = (1u << lane) - 1u;
    unsigned long long sum_cov = 0, sum_raw = 0;
    const int          div = a.l / a.P;

    for (;;) {
        int64_t c0 = 0;
        if (lane == 0) c0 = (int64_t)atomicAdd(a.work_counter, K3_CHUNK);
        c0 = __shfl_sync(FULL, c0, 0);
        if (c0 >= a.m) break;
        int64_t c1 = c0 + K3_CHUNK < a.m ? c0 + K3_CHUNK : a.m;
        // offsets of the chunk's reads: four coalesced loads for up to eight reads, handed out by shuffles
        const int     cnt = (int)(c1 - c0);
        const int64_t m_slot = lane <= cnt ? a.slot_off[c0 + lane] : 0, m_seq = lane <= cnt ? a.seq_off[c0 + lane] : 0;
        const int64_t m_rep = lane < cnt ? a.rep_cap_off[c0 + lane] : 0, m_cut = lane < cnt ? a.cut_cap_off[c0 + lane] : 0;
        for (int j = 0; j < cnt; j++) {
            const int64_t  i = c0 + j;
            const int64_t  L = __shfl_sync(FULL, m_seq, j + 1) - __shfl_sync(FULL, m_seq, j);
            const int64_t  base = __shfl_sync(FULL, m_slot, j);
            const int      nb = (int)(__shfl_sync(FULL, m_slot, j + 1) - base - 1);
            const int32_t* cov = a.cov + base;
            int2*          rep = a.rep + __shfl_sync(FULL, m_rep, j);
            int            nrep = 0;
            int            run_start = -1;

            auto emit = [&](int k1, int kend) {
                long long raw = (long long)(kend - k1) * a.reso;
                if (raw >= a.p) {
                    sum_raw += (unsigned long long)raw; // identical on all lanes; lane 0's copy is used
                    long long s = (long long)k1 * a.reso - a.f, e = (long long)kend * a.reso + a.f;
                    if (s <= 0) s = 0;
                    if (e >= L) e = L;
                    if (lane == 0) rep[nrep] = make_int2((int)s, (int)e);
                    nrep++;
                }
            };

            int cn[4]; // next group's coverage, loaded one iteration ahead
#pragma unroll
            for (int u = 0; u < 4; u++) { const int k = u * 32 + lane; cn[u] = k < nb ? cov[k] : 0; }
            for (int k0 = 0; k0 < nb; k0 += 128) {
                // four independent loads in flight per lane; bins past the read's end count as 0 and never set a mask bit
                int      c[4];
                unsigned m[4];
                int      part = 0; // |cov| < 2^31 / 4 in any input whose record count fits an int
#pragma unroll
                for (int u = 0; u < 4; u++) c[u] = cn[u];
                if (k0 + 128 < nb) {
#pragma unroll
                    for (int u = 0; u < 4; u++) { const int k = k0 + 128 + u * 32 + lane; cn[u] = k < nb ? cov[k] : 0; }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const bool in = k0 + u * 32 + lane < nb;
                    part += c[u];
                    m[u] = __ballot_sync(FULL, in && c[u] >= a.H);
                }
                sum_cov += (unsigned long long)(long long)part;
                if ((m[0] | m[1] | m[2] | m[3]) == 0u && run_start < 0) continue; // no bin at or above H here and no open run
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int kk = k0 + u * 32;
                    if (kk >= nb) break;
                    const unsigned mu = m[u]; // bins past nb have their bit clear, which closes a run that reaches the read's end at nb
                    int            bit = 0;
                    while (bit < 32) {
                        if (run_start < 0) {
                            unsigned mm = mu & (0xFFFFFFFFu << bit);
                            if (!mm) break;
                            int s = __ffs(mm) - 1;
                            run_start = kk + s; bit = s + 1;
                        } else {
                            unsigned inv = ~mu & (0xFFFFFFFFu << bit);
                            if (!inv) break;
                            int e = __ffs(inv) - 1;
                            emit(run_start, kk + e);
                            run_start = -1; bit = e + 1;
                        }
                    }
                }
            }
            if (run_start >= 0) emit(run_start, nb);
            if (lane == 0) a.rep_cnt[i] = nrep;
            __syncwarp();

            // stars and cuts
            const int64_t parts = L / a.P;
            const int     nstars = (int)(parts + 1 + (L % a.P != 0));
            int32_t*      cuts = a.cuts + __shfl_sync(FULL, m_cut, j);
            int           nf = 0;
            for (int j0 = 0; j0 < nstars; j0 += 32) {
                int  j = j0 + lane;
                bool keep = false;
                int  x = 0;
                if (j < nstars) {
                    x = (j <= parts) ? (int)((int64_t)j * a.P) : (int)L;
                    keep = true;
                    if (j > 0 && j < nstars - 1)
                        for (int q = 0; q < nrep; q++) { int2 r = rep[q]; if (r.x <= x && x <= r.y) { keep = false; break; } }
                }
                unsigned km = __ballot_sync(FULL, keep);
                int      fi = nf + __popc(km & lt_mask);
                if (keep && fi > 0 && (fi % div) == 0 && j != nstars - 1) cuts[fi / div - 1] = x;
                nf += __popc(km);
            }
            if (lane == 0) {
                int F = (nf - 1 + div - 1) / div;
                a.frag_cnt[i] = F < 1 ? 1 : F;
            }
        }
    }
    // sum_cov: per-lane partials; sum_raw: replicated on all lanes
    unsigned long long tc = warp_sum(sum_cov);
    if (lane == 0) {
        if (tc) atomicAdd(a.stats + 0, tc);
        if (sum_raw) atomicAdd(a.stats + 1, sum_raw);
    }
}

void launch_repeat_cut(const RepeatCutArgs& a, cudaStream_t st)
{
    if (a.m <= 0) return;
    int64_t warps_needed = (a.m + K3_CHUNK - 1) / K3_CHUNK;
    int64_t blocks = (warps_needed + (K3_THREADS / 32) - 1) / (K3_THREADS / 32);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_repeat_cut<<<(unsigned)blocks, K3_THREADS, 0, st>>>(a);
}

// ---------------------------------------------------------------- simulated-read names
// chop.hpp:14-70: start = atoi after the first '=' that follows the first ','; end = atoi after the first '-';
// align = between the first and second ','; chr / tail = from the last ','.
__global__ void __launch_bounds__(256) k_sim_parse(const uint8_t* __restrict__ names, const int64_t* __restrict__ name_off, int64_t own_first,
                                                    int64_t m, SimInfo* out, ErrState* err)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int64_t  gid = own_first + i;
    const uint8_t* s = names + name_off[gid];
    const int      n = (int)(name_off[gid + 1] - name_off[gid]);
    int c1 = -1, c2 = -1, eq = -1, da = -1, d0 = -1, c3 = -1, lc = -1;
    for (int k = 0; k < n; k++) {
        uint8_t c = s[k];
        if (c == ',') { if (c1 < 0) c1 = k; else if (c2 < 0) c2 = k; if (d0 >= 0 && c3 < 0) c3 = k; lc = k; }
        if (c == '=' && c1 >= 0 && eq < 0) eq = k;
        if (c == '-') { if (d0 < 0) d0 = k; if (eq >= 0 && da < 0) da = k; }
    }
    SimInfo o{};
    if (c1 < 0 || c2 < 0 || eq < 0 || da < 0 || d0 < 0 || c3 < 0) {
        err_min(err, RAFTK_E_SIM_NAME, (long long)gid);
        out[i] = o;
        return;
    }
    auto atoi_range = [&](int a, int b) { long long v = 0; for (int k = a; k < b && s[k] >= '0' && s[k] <= '9'; k++) v = v * 10 + (s[k] - '0'); return (int)v; };
    o.start_pos = atoi_range(eq + 1, da);
    o.end_pos = atoi_range(d0 + 1, c3);
    o.align_off = c1 + 1; o.align_len = c2 - c1 - 1; o.tail_off = lc;
    auto is = [&](const char* w) { if (o.align_len != 7) return false; for (int k = 0; k < 7; k++) if (s[o.align_off + k] != (uint8_t)w[k]) return false; return true; };
    o.flags = (is("forward") ? 1 : 0) | (is("reverse") ? 2 : 0);
    out[i] = o;
}
void launch_sim_parse(const uint8_t* names, const int64_t* name_off, int64_t own_first, int64_t m, SimInfo* out, ErrState* err, cudaStream_t st)
{
    if (m > 0) k_sim_parse<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(names, name_off, own_first, m, out, err);
}

// ---------------------------------------------------------------- fragments
// One thread per read: (a,b) of every fragment, the read= number, and the byte size of its FASTA
// record  ">read=" num "," name ",pos_on_original_read=" a "-" b "\n" bases[a:b] "\n"  (chop.hpp:261-265,314-318).
__global__ void __launch_bounds__(256) k_frag_expand(FragExpandArgs a)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.m) return;
    const int64_t  L = a.seq_off[i + 1] - a.seq_off[i];
    const int      F = a.frag_cnt[i];
    const int64_t  g0 = a.frag_base[i];
    const int32_t* cuts = a.cuts + a.cut_cap_off[i];
    const int64_t  gid = a.own_first + i;
    const int      name_len = (int)(a.name_off[gid + 1] - a.name_off[gid]);
    for (int j = 0; j < F; j++) {
        int64_t fa = (j == 0) ? 0 : (int64_t)cuts[j - 1] - a.v;
        int64_t fb = (j == F - 1) ? L : (int64_t)cuts[j];
        if (fa < 0) { // std::string::substr would throw (chop.hpp:318)
            err_min(a.err, RAFTK_E_NEG_START, (long long)gid);
            fa = 0;
        }
        int64_t g = g0 + j;
        a.frag_read[g] = (int32_t)i; a.frag_a[g] = (int32_t)fa; a.frag_b[g] = (int32_t)fb;
        int64_t num = a.read_num_base + g + 1;
        int     hdr;
        if (!a.sim) {
            hdr = 6 + dec_digits64((uint64_t)num) + 1 + name_len + 22 + dec_digits((uint32_t)fa) + 1 + dec_digits((uint32_t)fb) + 1;
        } else { // chop.hpp:252-258 (whole read), 293-310 (fragments: a header only for forward / reverse)
            const SimInfo si = a.sim[i];
            int x, y, ln;
            sim_header_numbers(si, F == 1, (int)fa, (int)fb, (int)L, &x, &y, &ln);
            hdr = (F == 1 || si.flags) ? 6 + dec_digits64((uint64_t)num) + 1 + si.align_len + 10 + dec_len_i32(x) + 1 + dec_len_i32(y) + 8 +
                                             dec_len_i32(ln) + (name_len - si.tail_off) + 1
                                       : 0;
        }
        a.frag_size[g] = (int32_t)(hdr + (fb - fa) + 1);
    }
}
void launch_frag_expand(const FragExpandArgs& a, cudaStream_t st)
{
    if (a.m > 0) k_frag_expand<<<(unsigned)((a.m + 255) / 256), 256, 0, st>>>(a);
}

// ---------------------------------------------------------------- split_naive
__global__ void __launch_bounds__(256) k_split_counts(const int64_t* __restrict__ seq_off, int64_t m, int sublen, int32_t* cnt)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int64_t L = seq_off[i + 1] - seq_off[i];
    cnt[i] = (int32_t)((L + sublen - 1) / sublen); // for (i = 0; i < length; i += subreadLength): split_naive.cpp:27
}
void launch_split_counts(const int64_t* seq_off, int64_t m, int sublen, int32_t* cnt, cudaStream_t st)
{
    if (m > 0) k_split_counts<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(seq_off, m, sublen, cnt);
}
__global__ void __launch_bounds__(256) k_split_expand(const int64_t* __restrict__ seq_off, const int64_t* __restrict__ name_off, int64_t own_first,
                                                       int64_t m, int sublen, const int64_t* __restrict__ base, int32_t* frag_read,
                                                       int32_t* frag_a, int32_t* frag_b, int32_t* frag_size)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int64_t L = seq_off[i + 1] - seq_off[i];
    const int     name_len = (int)(name_off[own_first + i + 1] - name_off[own_first + i]);
    int64_t       g = base[i];
    for (int64_t a0 = 0, k = 1; a0 < L; a0 += sublen, k++, g++) {
        int64_t b0 = a0 + sublen < L ? a0 + sublen : L;
        frag_read[g] = (int32_t)i; frag_a[g] = (int32_t)a0; frag_b[g] = (int32_t)b0;
        frag_size[g] = (int32_t)(1 + name_len + 1 + dec_digits64((uint64_t)k) + 1 + (b0 - a0) + 1); // ">" name "_" k "\n" bases "\n"
    }
}
void launch_split_expand(const int64_t* seq_off, const int64_t* name_off, int64_t own_first, int64_t m, int sublen, const int64_t* base,
                         int32_t* frag_read, int32_t* frag_a, int32_t* frag_b, int32_t* frag_size, cudaStream_t st)
{
    if (m > 0) k_split_expand<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(seq_off, name_off, own_first, m, sublen, base, frag_read, frag_a, frag_b, frag_size);
}

// ---------------------------------------------------------------- repeats: text sizes + compaction
// long_repeats.txt line: "read " i ", " then s "," e "    " per repeat, then "\n" (repeat.hpp:180-203)
__global__ void __launch_bounds__(256) k_rep_sizes(const int32_t* rep_cnt, const int64_t* rep_cap_off, const int2* rep, int64_t m,
                                                    int64_t own_first, int32_t* line_size)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int         sz = 5 + dec_digits64((uint64_t)(own_first + i)) + 2 + 1;
    const int2* r = rep + rep_cap_off[i];
    for (int q = 0; q < rep_cnt[i]; q++) sz += dec_len_i32(r[q].x) + 1 + dec_len_i32(r[q].y) + 4;
    line_size[i] = sz;
}
void launch_rep_sizes(const int32_t* rep_cnt, const int64_t* rep_cap_off, const int2* rep, int64_t m, int64_t own_first, int32_t* line_size,
                      cudaStream_t st)
{
    if (m > 0) k_rep_sizes<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(rep_cnt, rep_cap_off, rep, m, own_first, line_size);
}
__global__ void __launch_bounds__(256) k_rep_compact(const int32_t* rep_cnt, const int64_t* rep_cap_off, const int64_t* rep_off, const int2* rep,
                                                      int64_t m, int32_t* out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int2* r = rep + rep_cap_off[i];
    int64_t     o = rep_off[i];
    for (int q = 0; q < rep_cnt[i]; q++) { out[2 * (o + q)] = r[q].x; out[2 * (o + q) + 1] = r[q].y; }
}
void launch_rep_compact(const int32_t* rep_cnt, const int64_t* rep_cap_off, const int64_t* rep_off, const int2* rep, int64_t m, int32_t* out,
                        cudaStream_t st)
{
    if (m > 0) k_rep_compact<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(rep_cnt, rep_cap_off, rep_off, rep, m, out);
}

} // namespace raftk
