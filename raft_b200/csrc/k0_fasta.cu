// k0_fasta.cu — K0: FASTA tokenizer on the device.
//
// Replaces the record reader behind loadFASTA (chop.hpp:88-131, kseq_read kseq.h:240-298) for plain FASTA text:
// a record starts at a line whose first byte is '>' or '@' (kseq.h:247,263); its name is the header up to the
// first isspace byte (kseq.h:254), the rest of the header line is a comment; the sequence is the
// concatenation of the following lines without their newlines, blank lines skipped (kseq.h:263-269).
//
// The sequence arena is therefore a pure stream compaction of the file: keep every byte that is neither a
// newline nor part of a header line.  One pass, 16 KiB tiles staged by 1-D TMA:
//   * newline mask by SWAR + dp4a; '>' '@' '+' '\r' are only tested where they matter (line starts / the byte
//     before a newline), so there is a single full-width mask;
//   * "am I inside a header line at the start of the tile" is a last-writer scan over tiles, resolved with a
//     look-back: a tile that contains a newline knows its own carry-out and publishes it immediately;
//   * kept-byte and record counts get their global offsets from two decoupled look-backs;
//   * kept bytes come in long runs (whole lines): warps copy run pieces straight from the staged tile to the arena with
//     realigning shared-memory loads and aligned 128-bit stores.
// Outputs: the arena, and per record its file position and its offset in the arena.  Names are cut out by two
// tiny per-record kernels afterwards.
//
// FASTQ (template parameter FQ) is taken in its strict four-line form: '@' header, one line of bases, '+' line, one
// line of qualities of the same length.  There the role of a line is its index in the file modulo 4, so the header
// carry is replaced by a third look-back (newlines before the tile) and the kept bytes are the role-1 lines; the
// length condition is checked per record afterwards (k_fastq_verify).
//
// Inputs this kernel does not take (the caller falls back to the host reader, raftgpu_load_fasta): FASTA with a '+'
// at a line start, FASTQ with wrapped or truncated sequence / quality lines, any '\r' before a newline (kseq's
// strip rule depends on the accumulated length, kseq.h:189-190), a file that does not begin with '>' or '@', a
// marker as the very last byte.
#include "kernels.h"

namespace raftk {

constexpr int F0_THREADS = 256;
constexpr int F0_TILE = 16384;
constexpr int F0_WORDS = F0_TILE / 32; // 512 mask words, two per thread
constexpr int F0_RUNCAP = 1024;        // run pieces listed per tile (more: byte-wise path)
constexpr int F0_PIECE = 1024;

struct __align__(16) F0Smem {
    uint8_t  text[F0_TILE + 16];
    unsigned nl[F0_WORDS + 1];  // (+1: all-ones sentinel)
    unsigned hdr[F0_WORDS];     // byte belongs to a header line (including its newline)
    uint64_t bar;
    uint64_t bcast[2];
    int      scan_ws[34];
    int      tile;
    int      last_nl;           // position of the last newline in the tile, -1 if none
    int      carry_in;
    int      n_runs;
    unsigned drop_sum[F0_WORDS / 32]; // bit j of word k: mask word 32k+j has a dropped byte
    uint16_t run_src[F0_RUNCAP], run_dst[F0_RUNCAP], run_len[F0_RUNCAP]; // kept runs, cut into pieces of <= 1 KiB
};

__device__ __forceinline__ unsigned f0_zero_flags(unsigned x) { return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u; }
__device__ __forceinline__ unsigned f0_eq_mask16(const uint4& v, unsigned c4)
{
    unsigned lo = __dp4a(f0_zero_flags(v.x ^ c4), 0x08040201u, 0u);
    lo = __dp4a(f0_zero_flags(v.y ^ c4), 0x80402010u, lo);
    unsigned hi = __dp4a(f0_zero_flags(v.z ^ c4), 0x08040201u, 0u);
    hi = __dp4a(f0_zero_flags(v.w ^ c4), 0x80402010u, hi);
    return (lo >> 7) | ((hi >> 7) << 8);
}
__device__ __forceinline__ int f0_next_bit(const unsigned* m, int from)
{ // first set bit at position >= from; the sentinel word makes it return >= F0_TILE when there is none
    int      w = from >> 5;
    unsigned x = m[w] & (0xFFFFFFFFu << (from & 31));
    while (!x) x = m[++w];
    return (w << 5) + __ffs(x) - 1;
}

// carry status word: bits 63..62 = 2 when resolved (value in bit 0); 0 = not yet published; 1 = transparent
// (the tile has no newline: its carry-out equals its carry-in)
constexpr uint64_t F0_RESOLVED = 2ull << 62, F0_TRANSPARENT = 1ull << 62;

__device__ __forceinline__ int f0_resolve_carry(const uint64_t* status, int tile)
{ // all lanes of one warp; nearest predecessor with a resolved value
    int look = tile - 1;
    while (look >= 0) {
        int      idx = look - lane_id();
        uint64_t w = F0_RESOLVED; // before the file: not in a header
        if (idx >= 0) {
            w = ld_relaxed_u64(status + idx);
            while ((w >> 62) == 0) { __nanosleep(20); w = ld_relaxed_u64(status + idx); }
        }
        unsigned pm = __ballot_sync(FULL, (w >> 62) == 2);
        if (pm) { int src = __ffs(pm) - 1; return (int)(__shfl_sync(FULL, (unsigned)(w & 1ull), src)); }
        look -= 32;
    }
    return 0;
}

template <bool FQ>
__global__ void __launch_bounds__(F0_THREADS, 8) k_fasta_tokenize(FastaTokArgs a)
{
    extern __shared__ __align__(16) uint8_t f0_raw[];
    F0Smem& s = *reinterpret_cast<F0Smem*>(f0_raw);
    const int tid = threadIdx.x, lane = lane_id();
    if (tid == 0) { s.tile = atomicAdd(a.ticket, 1); s.last_nl = -1; mbar_init(&s.bar, 1); }
    __syncthreads();
    const int     tile = s.tile;
    const int64_t t0 = (int64_t)tile * F0_TILE;
    const int64_t avail = a.nbytes - t0;
    const int     want = (int)(avail < F0_TILE ? avail : F0_TILE);
    const int     bulk = want & ~15;
    if (tid == 0 && bulk > 0) { mbar_expect_tx(&s.bar, (uint32_t)bulk); tma_load_1d(s.text, a.text + t0, (uint32_t)bulk, &s.bar); }
    for (int j = bulk + tid; j < F0_TILE + 16; j += F0_THREADS) s.text[j] = (j < want) ? a.text[t0 + j] : (uint8_t)'\n';
    // the byte before the tile decides whether its first byte starts a line
    const bool prev_nl = (t0 == 0) || a.text[t0 - 1] == '\n';
    if (bulk > 0) mbar_wait(&s.bar, 0);
    __syncthreads();

    // ---- newline mask (bytes past the end of the file were filled with '\n': they are dropped like any newline)
    const uint4* t16 = reinterpret_cast<const uint4*>(s.text);
    int          my_last = -1;
    for (int g = tid; g < F0_TILE / 16; g += F0_THREADS) {
        unsigned m = f0_eq_mask16(t16[g], 0x0A0A0A0Au);
        unsigned p = __shfl_down_sync(FULL, m, 1);
        if (!(lane & 1)) s.nl[g >> 1] = m | (p << 16);
        const int room = want - g * 16; // bytes of this group inside the file
        const unsigned mv = room >= 16 ? m : (room > 0 ? (m & ((1u << room) - 1u)) : 0u);
        if (mv) my_last = g * 16 + (31 - __clz(mv));
    }
    for (int w = tid; w < F0_WORDS; w += F0_THREADS) s.hdr[w] = 0;
    if (tid == 0) s.nl[F0_WORDS] = 0xFFFFFFFFu;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) my_last = max(my_last, __shfl_xor_sync(FULL, my_last, d));
    if (lane == 0 && my_last >= 0) atomicMax(&s.last_nl, my_last);
    __syncthreads();

    const int wlo = tid * 2;
    // ---- strict FASTQ: the role of a line is its index in the file modulo 4, so count the newlines before every line start
    int64_t L0 = 0;      // index of the line that holds the tile's first byte
    int     nl_before = 0; // newlines of the tile before my first word
    if (FQ) {
        int cnt[2];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int      lo = (wlo + i) << 5;
            const unsigned valid = lo + 32 <= want ? 0xFFFFFFFFu : (lo < want ? ((1u << (want - lo)) - 1u) : 0u); // bytes past the file end were filled with newlines
            cnt[i] = __popc(s.nl[wlo + i] & valid);
        }
        int tot_nl;
        nl_before = block_exclusive_sum<int, F0_THREADS>(cnt[0] + cnt[1], s.scan_ws, &tot_nl);
        if (tid == 0) lookback_publish(a.st_nl, tile, (uint64_t)tot_nl);
        const int64_t base_nl = (int64_t)lookback_wait(a.st_nl, tile, (uint64_t)tot_nl, &s.bcast[0]);
        if (tid == 0 && tile == a.n_tiles - 1) a.totals[2] = base_nl + tot_nl;
        L0 = a.line_base + base_nl;
        __syncthreads(); // scan_ws and bcast are reused below
    }

    // ---- line starts in my two words: markers open header lines; '+' or a CR before a newline are not handled here
    unsigned  mk[2] = {0, 0};
    int       flags = 0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int w = wlo + i;
        unsigned  prev = w ? (s.nl[w - 1] >> 31) : (prev_nl ? 1u : 0u);
        unsigned  ls = (s.nl[w] << 1) | prev; // byte p follows a newline (blank lines included: their byte is '\n', tested below)
        unsigned  nlw = s.nl[w];
        // a '\r' right before a newline
        unsigned m = nlw;
        while (m) {
            int bit = __ffs(m) - 1; m &= m - 1;
            int q = (w << 5) + bit;
            if (q < want) { uint8_t b = q ? s.text[q - 1] : (t0 ? a.text[t0 - 1] : (uint8_t)0); if (b == '\r') flags |= 1; }
        }
        m = ls;
        while (m) {
            int bit = __ffs(m) - 1; m &= m - 1;
            int p = (w << 5) + bit;
            if (p >= want) break;
            uint8_t c = s.text[p];
            bool    drop_line; // the whole line is left out of the arena
            if (FQ) {
                const int64_t ln = L0 + nl_before + (i ? __popc(s.nl[wlo]) : 0) + __popc(nlw & ((1u << bit) - 1u)); // bytes < want here: the fill is never counted
                const int     role = (int)(ln & 3);
                if (role == 0) {
                    if (c != '@') flags |= 16;
                    mk[i] |= 1u << bit;
                    if ((ln >> 2) < a.rec_gcap) a.rec_gpos[ln >> 2] = a.text_gbase + t0 + p;
                } else if (role == 1) {
                    if (c == '>' || c == '+' || c == '@') flags |= 16; // kseq would end the sequence here (kseq.h:263)
                } else if (role == 2) {
                    if (c != '+') flags |= 16;
                } else if ((ln >> 2) < a.rec_gcap) a.qual_gpos[ln >> 2] = a.text_gbase + t0 + p;
                drop_line = role != 1;
            } else {
                drop_line = c == '>' || c == '@';
                if (drop_line) {
                    mk[i] |= 1u << bit;
                    if (t0 + p == a.nbytes - 1) flags |= 8; // marker as the last byte of the file: kseq returns no record
                } else if (c == '+') flags |= 2;
            }
            if (drop_line) {
                // header line [p, e]: e = its newline (or the end of the tile)
                int e = f0_next_bit(s.nl, p);
                if (e >= F0_TILE) e = F0_TILE - 1;
                int w0 = p >> 5, w1 = e >> 5;
                for (int ww = w0; ww <= w1; ww++) {
                    unsigned lo = ww == w0 ? (0xFFFFFFFFu << (p & 31)) : 0xFFFFFFFFu;
                    unsigned hi = ww == w1 ? (0xFFFFFFFFu >> (31 - (e & 31))) : 0xFFFFFFFFu;
                    atomicOr(&s.hdr[ww], lo & hi);
                }
            }
        }
    }
    if (!FQ && t0 == 0 && tid == 0 && a.first_chunk && !(s.text[0] == '>' || s.text[0] == '@')) flags |= 4;
    if (tid == 0 && a.last_chunk && t0 + want == a.nbytes && s.text[want - 1] == '\r') flags |= 1; // CR at EOF: kseq's strip rule again
    if (flags) atomicOr(a.flags, flags);

    // ---- carry: is the tile's first byte inside a header line?  (last-writer scan over tiles; in FASTQ mode the role of
    // the line that continues into the tile says it)
    if (FQ) {
        if (tid == 0) s.carry_in = !prev_nl && (L0 & 3) != 1;
    } else if (tid == 0) {
        const int ql = s.last_nl;
        uint64_t  st;
        if (ql >= 0) { // the line open at the end of the tile starts at ql+1 (in this tile, or exactly at the next tile)
            int p = ql + 1;
            st = F0_RESOLVED | ((p < want && (s.text[p] == '>' || s.text[p] == '@')) ? 1ull : 0ull);
        } else if (prev_nl) {
            st = F0_RESOLVED | ((s.text[0] == '>' || s.text[0] == '@') ? 1ull : 0ull);
        } else st = F0_TRANSPARENT;
        st_relaxed_u64(a.st_carry + tile, st);
    }
    if (!FQ && tid < 32) {
        int cin = prev_nl ? 0 : f0_resolve_carry(a.st_carry, tile);
        if (lane == 0) {
            s.carry_in = cin;
            if (s.last_nl < 0 && !prev_nl) st_relaxed_u64(a.st_carry + tile, F0_RESOLVED | (uint64_t)cin); // transparent tile: now known
        }
    }
    __syncthreads();
    if (s.carry_in && tid == 0) { // the header line continues from the previous tile up to the first newline
        int e = f0_next_bit(s.nl, 0);
        if (e >= F0_TILE) e = F0_TILE - 1;
        for (int ww = 0; ww <= (e >> 5); ww++) atomicOr(&s.hdr[ww], ww == (e >> 5) ? (0xFFFFFFFFu >> (31 - (e & 31))) : 0xFFFFFFFFu);
    }
    __syncthreads();

    // ---- keep = neither newline nor header; counts and global offsets
    unsigned keep[2];
    int      nkeep = 0, nmk = 0;
#pragma unroll
    for (int i = 0; i < 2; i++) { keep[i] = ~(s.nl[wlo + i] | s.hdr[wlo + i]); nkeep += __popc(keep[i]); nmk += __popc(mk[i]); }
    int tot_keep, tot_mk;
    int ex_keep = block_exclusive_sum<int, F0_THREADS>(nkeep, s.scan_ws, &tot_keep);
    __syncthreads();
    int ex_mk = block_exclusive_sum<int, F0_THREADS>(nmk, s.scan_ws, &tot_mk);
    if (tid == 0) { lookback_publish(a.st_keep, tile, (uint64_t)tot_keep); lookback_publish(a.st_rec, tile, (uint64_t)tot_mk); }
    const int64_t base_keep = (int64_t)lookback_wait(a.st_keep, tile, (uint64_t)tot_keep, &s.bcast[0]);
    const int64_t base_rec = (int64_t)lookback_wait(a.st_rec, tile, (uint64_t)tot_mk, &s.bcast[1]);
    if (tid == 0 && tile == a.n_tiles - 1) { a.totals[0] = base_rec + tot_mk; a.totals[1] = base_keep + tot_keep; }

    // ---- records that start in my words
    {
        int kept_before = ex_keep, r = ex_mk;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            unsigned m = mk[i];
            while (m) {
                int     bit = __ffs(m) - 1; m &= m - 1;
                int64_t idx = base_rec + r++;
                if (idx < a.rec_cap) {
                    a.rec_pos[idx] = t0 + ((wlo + i) << 5) + bit;
                    a.seq_off[idx] = a.seq_off_base + base_keep + kept_before + __popc(keep[i] & ((1u << bit) - 1u)); // header bytes are not kept
                }
            }
            kept_before += __popc(keep[i]);
        }
    }
    // ---- kept bytes come in long runs (whole lines): list the runs (pieces of <= 1 KiB) and let each warp copy pieces with
    // lane-consecutive bytes; per-thread copying of its own 64 bytes would be a 16-way shared-memory bank conflict.
    const uintptr_t gdst0 = (uintptr_t)a.seq_out + (uintptr_t)base_keep;
    if (tid == 0) s.n_runs = 0;
    // summary of the dropped-byte mask (one ballot per 32 words) so that the end of a long run is found in a few steps
    for (int w = tid; w < F0_WORDS; w += F0_THREADS) {
        unsigned any = __ballot_sync(FULL, (s.nl[w] | s.hdr[w]) != 0);
        if (lane == 0) s.drop_sum[w >> 5] = any;
    }
    __syncthreads();
    {
        int dst = ex_keep;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int      w = wlo + i;
            const unsigned k = keep[i];
            const unsigned prevbit = w ? ((~(s.nl[w - 1] | s.hdr[w - 1])) >> 31) : 0u; // was the byte before this word kept?
            unsigned       starts = k & ~((k << 1) | prevbit);
            while (starts) {
                int bit = __ffs(starts) - 1; starts &= starts - 1;
                int p = (w << 5) + bit;
                // run end: first dropped byte at or after p
                int      ww = w;
                unsigned d = (s.nl[ww] | s.hdr[ww]) & (0xFFFFFFFFu << bit); // dropped bits at or after p in this word
                if (!d) { // next mask word with a dropped byte, through the summary
                    int      sw = (w + 1) >> 5;
                    unsigned sm = sw < F0_WORDS / 32 ? (s.drop_sum[sw] & (0xFFFFFFFFu << ((w + 1) & 31))) : 0u;
                    if (((w + 1) & 31) == 0 && sw < F0_WORDS / 32) sm = s.drop_sum[sw];
                    while (!sm && ++sw < F0_WORDS / 32) sm = s.drop_sum[sw];
                    if (sm) { ww = (sw << 5) + __ffs(sm) - 1; d = s.nl[ww] | s.hdr[ww]; }
                }
                int e = d ? (ww << 5) + __ffs(d) - 1 : F0_TILE;
                int len = e - p, dd = dst + __popc(k & ((1u << bit) - 1u));
                for (int o = 0; o < len; o += F0_PIECE) {
                    int slot = atomicAdd(&s.n_runs, 1);
                    if (slot < F0_RUNCAP) { s.run_src[slot] = (uint16_t)(p + o); s.run_dst[slot] = (uint16_t)(dd + o); s.run_len[slot] = (uint16_t)min(F0_PIECE, len - o); }
                }
            }
            dst += __popc(k);
        }
    }
    __syncthreads();
    uint8_t* const gout = reinterpret_cast<uint8_t*>(gdst0); // arena position of the tile's first kept byte
    if (s.n_runs <= F0_RUNCAP) {
        // every warp copies pieces of <= 1 KiB straight from the staged text to the arena: bytes up to the destination's
        // first 16-byte boundary, then aligned 128-bit streaming stores fed by realigning shared-memory loads, then bytes
        const int nr = s.n_runs, warp = warp_id();
        for (int r = warp; r < nr; r += F0_THREADS / 32) {
            const uint8_t* src = s.text + s.run_src[r];
            uint8_t*       o = gout + s.run_dst[r];
            const int      n = s.run_len[r];
            int head = (int)((16 - ((uintptr_t)o & 15)) & 15);
            if (head > n) head = n;
            if (lane < head) o[lane] = src[lane];
            const int nbody = (n - head) >> 4;
            uint4*    o16 = reinterpret_cast<uint4*>(o + head);
            for (int c = lane; c < nbody; c += 32) stg_stream(o16 + c, lds_unaligned16(src + head + (c << 4)));
            const int done = head + (nbody << 4);
            if (done + lane < n) o[done + lane] = src[done + lane];
        }
    } else { // pathological tile (thousands of tiny lines): every thread copies the kept bytes of its own words
        uint8_t* o = gout + ex_keep;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            unsigned       m = keep[i];
            const uint8_t* src = s.text + ((wlo + i) << 5);
            while (m) { int bit = __ffs(m) - 1; m &= m - 1; *o++ = src[bit]; }
        }
    }
}

int         fasta_tokenize_tiles(int64_t nbytes) { return (int)((nbytes + F0_TILE - 1) / F0_TILE); }
cudaError_t launch_fasta_tokenize(const FastaTokArgs& a, cudaStream_t st)
{
    if (a.n_tiles <= 0) return cudaSuccess;
    auto go = [&](auto kern) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(F0Smem));
        if (e != cudaSuccess) return e;
        kern<<<a.n_tiles, F0_THREADS, sizeof(F0Smem), st>>>(a);
        return cudaGetLastError();
    };
    return a.fastq ? go(k_fasta_tokenize<true>) : go(k_fasta_tokenize<false>);
}

// ---------------------------------------------------------------- strict FASTQ: quality lines
// kseq reads quality lines until it has as many bytes as bases (kseq.h:290-296).  With exactly one quality line of the
// same length, which ends right before the next record's '@' (or the end of the file), the four-lines-per-record reading
// of k_fasta_tokenize<true> is what kseq does; anything else (wrapped or truncated qualities) is left to the host reader.
__global__ void __launch_bounds__(256) k_fastq_verify(int64_t n, const int64_t* __restrict__ seq_off, const int64_t* __restrict__ rec_gpos,
                                                      const int64_t* __restrict__ qual_gpos, int64_t file_end, int* flags)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t bases = seq_off[i + 1] - seq_off[i];
    const int64_t next = i + 1 < n ? rec_gpos[i + 1] : file_end; // one past the quality line's newline (a missing final newline is counted by the caller)
    if (next - qual_gpos[i] - 1 != bases) atomicOr(flags, 16);
}
void launch_fastq_verify(int64_t n, const int64_t* seq_off, const int64_t* rec_gpos, const int64_t* qual_gpos, int64_t file_end, int* flags,
                         cudaStream_t st)
{
    if (n > 0) k_fastq_verify<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, seq_off, rec_gpos, qual_gpos, file_end, flags);
}

// ---------------------------------------------------------------- names: header up to the first isspace byte (kseq.h:254)
__global__ void __launch_bounds__(256) k_fasta_name_len(const uint8_t* __restrict__ text, int64_t nbytes, const int64_t* __restrict__ rec_pos,
                                                         int64_t n, int32_t* name_len)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t p = rec_pos[i] + 1, q = p;
    while (q < nbytes) { uint8_t c = text[q]; if (c == ' ' || (c >= 9 && c <= 13)) break; q++; }
    name_len[i] = (int32_t)(q - p);
}
__global__ void __launch_bounds__(256) k_fasta_name_copy(const uint8_t* __restrict__ text, const int64_t* __restrict__ rec_pos,
                                                          const int64_t* __restrict__ name_off, int64_t n, uint8_t* names)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* src = text + rec_pos[i] + 1;
    uint8_t*       dst = names + name_off[i];
    int64_t        len = name_off[i + 1] - name_off[i];
    for (int64_t k = 0; k < len; k++) dst[k] = src[k];
}
void launch_fasta_name_len(const uint8_t* text, int64_t nbytes, const int64_t* rec_pos, int64_t n, int32_t* name_len, cudaStream_t st)
{
    if (n > 0) k_fasta_name_len<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(text, nbytes, rec_pos, n, name_len);
}
void launch_fasta_name_copy(const uint8_t* text, const int64_t* rec_pos, const int64_t* name_off, int64_t n, uint8_t* names, cudaStream_t st)
{
    if (n > 0) k_fasta_name_copy<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(text, rec_pos, name_off, n, names);
}

__global__ void __launch_bounds__(256) k_add_offset_i64(const int64_t* __restrict__ src, int64_t n, int64_t add, int64_t* dst)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] + add;
}
void launch_add_offset_i64(const int64_t* src, int64_t n, int64_t add, int64_t* dst, cudaStream_t st)
{
    if (n > 0) k_add_offset_i64<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, n, add, dst);
}

} // namespace raftk
