// host_io.cpp — host side of the boundary: FASTA/FASTQ(+gz) reader and the file-level drop-in for
// break_long_reads (chop.hpp:331-373).  No compute here: everything numeric happens on the device
// through the C ABI in include/raft_b200.h.
#include <cuda_runtime.h>
#include <zlib.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/raft_b200.h"

namespace {

// Streaming FASTA/FASTQ record reader with the record grammar of kseq_read (kseq.h:240-298) as used
// by loadFASTA (chop.hpp:88-131): name = header up to the first isspace byte, the rest of the header
// line is a comment, sequence = following lines joined (blank lines skipped, a trailing '\r' of a line
// dropped once more than one byte has been collected), a '+' line switches to FASTQ quality which
// must reach the sequence length.  Lengths and names stop at an embedded NUL (strlen / std::string).
class SeqReader {
public:
    std::vector<uint8_t> seq, names;
    std::vector<int64_t> seq_off{0}, name_off{0};
    bool                 fastq_truncated = false;

    void feed(const uint8_t* p, size_t n)
    {
        size_t i = 0;
        while (i < n && !stopped_) {
            switch (st_) {
            case SEEK:
                while (i < n && p[i] != '>' && p[i] != '@') i++;
                if (i < n) { i++; begin_record(); }
                break;
            case NAME:
                any_after_marker_ = true;
                while (i < n && !is_space(p[i])) name_.push_back((char)p[i++]);
                if (i < n) { st_ = (p[i] == '\n') ? SEQ_BOL : COMMENT; i++; }
                break;
            case COMMENT:
                while (i < n && p[i] != '\n') i++;
                if (i < n) { i++; st_ = SEQ_BOL; }
                break;
            case SEQ_BOL: {
                uint8_t c = p[i++];
                if (c == '>' || c == '@') { end_record(); begin_record(); }
                else if (c == '+') st_ = PLUS;
                else if (c == '\n') {}
                else { cur_.push_back(c); st_ = SEQ_LINE; }
                break;
            }
            case SEQ_LINE: {
                const uint8_t* e = (const uint8_t*)memchr(p + i, '\n', n - i);
                size_t         k = e ? (size_t)(e - p) : n;
                cur_.insert(cur_.end(), p + i, p + k);
                i = k;
                if (e) { i++; strip_cr(cur_); st_ = SEQ_BOL; }
                break;
            }
            case PLUS:
                while (i < n && p[i] != '\n') i++;
                if (i < n) { i++; st_ = QUAL; qual_len_ = 0; qual_last_ = 0; qual_lines_ = 0; }
                break;
            case QUAL: {
                // kseq.h:292: read lines while the quality is shorter than the sequence
                if (qual_lines_ > 0 && qual_len_ >= cur_.size()) { end_fastq(); break; }
                const uint8_t* e = (const uint8_t*)memchr(p + i, '\n', n - i);
                size_t         k = e ? (size_t)(e - p) : n;
                if (k > i) { qual_len_ += k - i; qual_last_ = p[k - 1]; }
                i = k;
                if (e) {
                    i++;
                    if (qual_len_ > 1 && qual_last_ == '\r') { qual_len_--; qual_last_ = 0; }
                    qual_lines_++;
                    if (qual_len_ >= cur_.size()) end_fastq();
                }
                break;
            }
            }
        }
    }

    void finish()
    {
        if (stopped_) return;
        switch (st_) {
        case SEEK: break;
        case NAME: if (any_after_marker_) end_record(); break; // EOF right after the marker: no record (kseq.h:254-255)
        case COMMENT: case SEQ_BOL: end_record(); break;
        case SEQ_LINE: strip_cr(cur_); end_record(); break;
        case PLUS: fastq_truncated = true; break;               // kseq.h:288-289
        case QUAL:
            if (qual_len_ > 1 && qual_last_ == '\r') qual_len_--;
            end_fastq();
            break;
        }
        st_ = SEEK;
    }

private:
    enum St { SEEK, NAME, COMMENT, SEQ_BOL, SEQ_LINE, PLUS, QUAL } st_ = SEEK;
    std::string          name_;
    std::vector<uint8_t> cur_;
    size_t               qual_len_ = 0, qual_lines_ = 0;
    uint8_t              qual_last_ = 0;
    bool                 any_after_marker_ = false, stopped_ = false;

    static bool is_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }
    static void strip_cr(std::vector<uint8_t>& s) { if (s.size() > 1 && s.back() == '\r') s.pop_back(); }
    void begin_record() { name_.clear(); cur_.clear(); any_after_marker_ = false; st_ = NAME; }
    void end_record()
    {
        size_t sl = strnlen((const char*)cur_.data(), cur_.size());
        size_t nl = strnlen(name_.data(), name_.size());
        seq.insert(seq.end(), cur_.begin(), cur_.begin() + sl);
        names.insert(names.end(), name_.begin(), name_.begin() + nl);
        seq_off.push_back((int64_t)seq.size());
        name_off.push_back((int64_t)names.size());
        st_ = SEEK;
    }
    void end_fastq()
    {
        if (qual_len_ != cur_.size()) { fastq_truncated = true; stopped_ = true; return; } // kseq.h:297-298: -2 ends loadFASTA's loop
        end_record(); // last_char = 0: the next record is searched from the following byte (kseq.h:296)
    }
};

template <typename T>
T* steal(const std::vector<T>& v)
{
    T* p = (T*)malloc(sizeof(T) * (v.size() ? v.size() : 1));
    if (p && !v.empty()) memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}

// Sequential byte source over several files (gzip or plain), read by a background thread into two pinned buffers.
class FilePipeline {
public:
    FilePipeline(const char* const* paths, int n) : paths_(paths, paths + n)
    {
        for (int k = 0; k < 2; k++) {
            void* p = nullptr;
            if (cudaHostAlloc(&p, CAP, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); p = malloc(CAP); pinned_[k] = false; }
            buf_[k] = (uint8_t*)p;
        }
        th_ = std::thread([this] { run(); });
    }
    ~FilePipeline()
    {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
        cv_.notify_all();
        if (th_.joinable()) th_.join();
        for (int k = 0; k < 2; k++) if (buf_[k]) { if (pinned_[k]) cudaFreeHost(buf_[k]); else free(buf_[k]); }
    }
    // next filled buffer; false on a read error
    bool next(const uint8_t** data, size_t* fill, bool* last)
    {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [&] { return state_[cons_] == FULL || error_; });
        if (error_) return false;
        *data = buf_[cons_]; *fill = fill_[cons_]; *last = last_[cons_];
        return true;
    }
    void release()
    {
        { std::lock_guard<std::mutex> g(mu_); state_[cons_] = EMPTY; cons_ ^= 1; }
        cv_.notify_all();
    }

private:
    static constexpr size_t CAP = 256u << 20;
    enum { EMPTY, FULL };
    void run()
    {
        size_t file = 0;
        gzFile gz = nullptr;
        int    k = 0;
        bool   done = false;
        while (!done) {
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return state_[k] == EMPTY || stop_; });
                if (stop_) break;
            }
            size_t fill = 0;
            while (fill < CAP && !done) {
                if (!gz) {
                    if (file == paths_.size()) { done = true; break; }
                    gz = gzopen(paths_[file++], "r"); // transparent for plain files, like the reference (paf.hpp:29)
                    if (!gz) { fail(); return; }
                    gzbuffer(gz, 4 << 20);
                }
                int got = gzread(gz, buf_[k] + fill, (unsigned)std::min<size_t>(CAP - fill, 1u << 30));
                if (got < 0) { gzclose(gz); fail(); return; }
                if (got == 0) { gzclose(gz); gz = nullptr; continue; }
                fill += (size_t)got;
            }
            {
                std::lock_guard<std::mutex> g(mu_);
                fill_[k] = fill; last_[k] = done; state_[k] = FULL;
            }
            cv_.notify_all();
            k ^= 1;
        }
        if (gz) gzclose(gz);
    }
    void fail()
    {
        { std::lock_guard<std::mutex> g(mu_); error_ = true; }
        cv_.notify_all();
    }
    std::vector<const char*> paths_;
    uint8_t*                 buf_[2] = {nullptr, nullptr};
    bool                     pinned_[2] = {true, true};
    size_t                   fill_[2] = {0, 0};
    bool                     last_[2] = {false, false};
    int                      state_[2] = {EMPTY, EMPTY};
    int                      cons_ = 0;
    bool                     stop_ = false, error_ = false;
    std::mutex               mu_;
    std::condition_variable  cv_;
    std::thread              th_;
};

bool file_missing_or_empty(const char* fn)
{ // chop.hpp:326-349
    std::ifstream f(fn);
    return !f || f.peek() == std::ifstream::traits_type::eof();
}

} // namespace

extern "C" void raftgpu_free_host(void* p) { free(p); }

extern "C" int raftgpu_load_fasta(const char* path, int64_t* n, int64_t** seq_off, uint8_t** seq, int64_t** name_off, uint8_t** names)
{
    if (!path || !n || !seq_off || !seq || !name_off || !names) return RAFTGPU_E_ARG;
    gzFile fp = gzopen(path, "r"); // transparent for plain files, like the reference (chop.hpp:93)
    if (!fp) return RAFTGPU_E_IO;
    gzbuffer(fp, 1 << 20);
    SeqReader            rd;
    std::vector<uint8_t> buf(8 << 20);
    for (;;) {
        int got = gzread(fp, buf.data(), (unsigned)buf.size());
        if (got <= 0) break;
        rd.feed(buf.data(), (size_t)got);
    }
    gzclose(fp);
    rd.finish();
    *n = (int64_t)rd.seq_off.size() - 1;
    *seq_off = steal(rd.seq_off); *name_off = steal(rd.name_off);
    *seq = steal(rd.seq); *names = steal(rd.names);
    if (!*seq_off || !*name_off || !*seq || !*names) return RAFTGPU_E_NOMEM;
    return RAFTGPU_OK;
}

// Writes one output stream to `path`: this thread fetches 256 MiB windows into two pinned buffers (D2H through the
// C ABI) while a writer thread drains the other buffer to the file.
static bool write_stream(raftgpu_ctx* ctx, int which, const std::string& path, std::string& err)
{
    uint64_t total = 0;
    if (raftgpu_output_size(ctx, which, &total)) { err = raftgpu_last_error(ctx); return false; }
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { err = "cannot open " + path; return false; }
    const size_t W = (size_t)std::min<uint64_t>(256u << 20, total ? total : 1);
    uint8_t*     buf[2] = {nullptr, nullptr};
    bool         pinned[2] = {true, true};
    for (int k = 0; k < 2; k++) {
        void* p = nullptr;
        if (cudaHostAlloc(&p, W, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); p = malloc(W); pinned[k] = false; }
        buf[k] = (uint8_t*)p;
    }
    std::mutex              mu;
    std::condition_variable cv;
    size_t                  len[2] = {0, 0};
    bool                    full[2] = {false, false}, done = false, werr = false;
    std::thread writer([&] {
        for (int k = 0;; k ^= 1) {
            std::unique_lock<std::mutex> g(mu);
            cv.wait(g, [&] { return full[k] || done; });
            if (!full[k]) return;
            g.unlock();
            if (!werr && fwrite(buf[k], 1, len[k], f) != len[k]) werr = true;
            g.lock();
            full[k] = false;
            cv.notify_all();
        }
    });
    bool ok = true;
    int  k = 0;
    for (uint64_t off = 0; off < total && ok; off += W, k ^= 1) {
        {
            std::unique_lock<std::mutex> g(mu);
            cv.wait(g, [&] { return !full[k]; });
        }
        const size_t n = (size_t)std::min<uint64_t>(W, total - off);
        const int    st = raftgpu_fetch(ctx, which, off, buf[k], n);
        if (st) { err = std::string(raftgpu_strerror(st)) + ": " + raftgpu_last_error(ctx); ok = false; break; }
        {
            std::lock_guard<std::mutex> g(mu);
            len[k] = n; full[k] = true;
        }
        cv.notify_all();
    }
    {
        std::unique_lock<std::mutex> g(mu);
        cv.wait(g, [&] { return !full[0] && !full[1]; });
        done = true;
    }
    cv.notify_all();
    writer.join();
    for (int q = 0; q < 2; q++) if (buf[q]) { if (pinned[q]) cudaFreeHost(buf[q]); else free(buf[q]); }
    if (fclose(f) != 0) werr = true;
    if (werr && ok) { err = "short write to " + path; ok = false; }
    return ok;
}

extern "C" int raftgpu_break_long_reads(const char* readfilename, const char* paffilename, const raftgpu_params* p, const char* prefix,
                                        int device, raftgpu_stats* stats_out)
{
    return raftgpu_break_long_reads_multi(readfilename, 1, &paffilename, p, prefix, device, stats_out);
}

extern "C" int raftgpu_break_long_reads_multi(const char* readfilename, int n_paf, const char* const* paffilenames, const raftgpu_params* p,
                                              const char* prefix, int device, raftgpu_stats* stats_out)
{
    if (!readfilename || n_paf < 1 || !paffilenames || !p || !prefix) return RAFTGPU_E_ARG;
    for (int k = 0; k < n_paf; k++) if (!paffilenames[k]) return RAFTGPU_E_ARG;
    const std::string pre(prefix);
    { std::ofstream touch(pre + ".reads.fasta"); } // chop.hpp:333: created before the inputs are validated
    if (file_missing_or_empty(readfilename)) {
        printf("ERROR, break_long_reads(), %s input file either does not exist or is empty\n", readfilename);
        return RAFTGPU_E_IO;
    }
    for (int k = 0; k < n_paf; k++)
        if (file_missing_or_empty(paffilenames[k])) {
            printf("ERROR, break_long_reads(), %s input file either does not exist or is empty\n", paffilenames[k]);
            return RAFTGPU_E_IO;
        }
    raftgpu_ctx* ctx = nullptr;
    int          st = raftgpu_create(p, device, &ctx);
    if (st) { fprintf(stderr, "raft_b200: %s\n", raftgpu_strerror(st)); return st; }
    auto fail = [&](int code) {
        fprintf(stderr, "raft_b200: %s: %s\n", raftgpu_strerror(code), raftgpu_last_error(ctx));
        raftgpu_destroy(ctx);
        return code;
    };

    // reads: plain FASTA text is tokenised on the device; gzip, FASTQ and CR LF text go through the host reader
    bool on_device = false;
    {
        FILE* f = fopen(readfilename, "rb");
        if (!f) return fail(RAFTGPU_E_IO);
        unsigned char magic[2] = {0, 0};
        size_t        got = fread(magic, 1, 2, f);
        if (!(got == 2 && magic[0] == 0x1f && magic[1] == 0x8b)) {
            fseeko(f, 0, SEEK_END);
            const uint64_t fsize = (uint64_t)ftello(f);
            FilePipeline   pipe(&readfilename, 1); // reader thread + two pinned buffers, like the PAF below
            st = RAFTGPU_OK;
            while (st == RAFTGPU_OK) {
                const uint8_t* data = nullptr;
                size_t         fill = 0;
                bool           last = false;
                if (!pipe.next(&data, &fill, &last)) { fclose(f); return fail(RAFTGPU_E_IO); }
                st = raftgpu_ingest_fasta(ctx, data, fill, last ? 1 : 0, fsize);
                pipe.release();
                if (last) break;
            }
            if (st == RAFTGPU_OK) on_device = true;
            else if (st != RAFTGPU_E_UNSUPPORTED) { fclose(f); return fail(st); }
        }
        fclose(f);
    }
    raftgpu_stats s{};
    if (!on_device) {
        int64_t  n = 0;
        int64_t *seq_off = nullptr, *name_off = nullptr;
        uint8_t *seq = nullptr, *names = nullptr;
        if ((st = raftgpu_load_fasta(readfilename, &n, &seq_off, &seq, &name_off, &names))) return fail(st);
        st = raftgpu_set_reads(ctx, n, seq_off, seq, name_off, names);
        free(seq_off); free(seq); free(name_off); free(names);
        if (st) return fail(st);
    }

    // PAF: one reader thread inflates (gz, paf.hpp:24-38) or reads (plain) the files back to back into two pinned
    // buffers while this thread hands the other buffer to the tokenizer, so disk/zlib time overlaps H2D + parse and
    // only the decoded records (28 B each) stay on the device: the text itself may exceed HBM.  Several files behave
    // like `cat a b | raft ...` (README.md:32-38: hifiasm writes two *.ovlp.paf files).
    {
        FilePipeline pipe(paffilenames, n_paf);
        for (;;) {
            const uint8_t* data = nullptr;
            size_t         fill = 0;
            bool           last = false;
            if (!pipe.next(&data, &fill, &last)) return fail(RAFTGPU_E_IO);
            st = raftgpu_ingest_paf(ctx, data, fill, last ? 1 : 0);
            pipe.release();
            if (st) return fail(st);
            if (last) break;
        }
    }
    if ((st = raftgpu_run(ctx, &s))) {
        // the reference prints these before it would have crashed; keep stdout comparable up to the failure
        return fail(st);
    }
    printf("Real Reads %d \n", s.real_reads);                              // chop.hpp:105
    printf("INFO, Symmetric overlaps %d \n", s.symmetric);                 // chop.hpp:189
    printf("INFO, length of alignments  %d()\n", (int)s.n_records);        // chop.hpp:190
    printf("high_cov %d\n", s.high_cov);                                   // repeat.hpp:91
    double coverage_per_window = (double)s.total_cov / s.total_windows;    // repeat.hpp:173-178
    double fraction_of_repeat_length = (double)s.total_repeat_len / s.total_read_len;
    printf("coverage per window is %f \n", coverage_per_window);
    printf("coverage per window/average coverage is %f \n", coverage_per_window / p->est_cov);
    printf("fraction_of_repeat_length %f \n", fraction_of_repeat_length);

    std::string err;
    const char* suf[4] = {".coverage.txt", ".long_repeats.txt", ".long_repeats.bed", ".reads.fasta"};
    for (int w = 0; w < 4; w++)
        if (!write_stream(ctx, w, pre + suf[w], err)) { fprintf(stderr, "raft_b200: %s\n", err.c_str()); raftgpu_destroy(ctx); return RAFTGPU_E_IO; }
    if (stats_out) *stats_out = s;
    raftgpu_destroy(ctx);
    return RAFTGPU_OK;
}
