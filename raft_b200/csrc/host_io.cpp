// host_io.cpp — host side of the boundary: FASTA/FASTQ(+gz) reader and the file-level drop-in for
// break_long_reads (chop.hpp:331-373).  No compute here: everything numeric happens on the device
// through the C ABI in include/raft_b200.h.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/statvfs.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
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
#include "file_io.h"

namespace {
using raftio::SliceWriter;
using raftio::map_threads;
using raftio::parallel_pread;

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

// Two pinned host buffers shared by every file phase of a run (reads in, PAF in, outputs out): pinning is the expensive
// part of a small run, so it is done once and sized by the largest file at hand.
struct PinnedPair {
    uint8_t* buf[2] = {nullptr, nullptr};
    bool     pinned[2] = {false, false};
    size_t   cap = 0;
    explicit PinnedPair(size_t want)
    {
        cap = std::min<size_t>(64u << 20, std::max<size_t>(want, 1u << 20)); // 64 MiB chunks keep every pipe busy; pinning costs ~0.4 ms per MiB
        cap = (cap + 4095) & ~(size_t)4095;
        for (int k = 0; k < 2; k++) {
            void* p = nullptr;
            if (cudaHostAlloc(&p, cap, cudaHostAllocDefault) == cudaSuccess) pinned[k] = true;
            else { cudaGetLastError(); p = malloc(cap); }
            buf[k] = (uint8_t*)p;
        }
    }
    ~PinnedPair()
    {
        for (int k = 0; k < 2; k++) if (buf[k]) { if (pinned[k]) cudaFreeHost(buf[k]); else free(buf[k]); }
    }
    bool ok() const { return buf[0] && buf[1]; }
};

// One piece of input: bytes [off, off+len) of a plain file (len < 0: to the end), or a whole gzip file.
struct Segment { std::string path; int64_t off = 0, len = -1; bool gz = false; };

bool is_gzip_file(const char* path)
{
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    unsigned char m[2] = {0, 0};
    const size_t  got = fread(m, 1, 2, f);
    fclose(f);
    return got == 2 && m[0] == 0x1f && m[1] == 0x8b;
}
int64_t file_size(const char* path)
{
    struct stat sb;
    return stat(path, &sb) == 0 ? (int64_t)sb.st_size : -1;
}

// Sequential byte source over a list of segments, read by a background thread into the two pinned buffers: plain files
// with pread straight into pinned memory, gzip files through zlib (paf.hpp:24-38, chop.hpp:93).  With `join_lines` a
// newline is inserted between two segments when the first does not end with one, so that the last line of one file is
// never glued to the first line of the next.
class ByteSource {
public:
    ByteSource(PinnedPair& pp, std::vector<Segment> segs, bool join_lines) : pp_(pp), segs_(std::move(segs)), join_(join_lines)
    {
        th_ = std::thread([this] { run(); });
    }
    ~ByteSource()
    {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
        cv_.notify_all();
        if (th_.joinable()) th_.join();
    }
    // next filled buffer; false on a read error
    bool next(const uint8_t** data, size_t* fill, bool* last)
    {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [&] { return state_[cons_] == FULL || error_; });
        if (error_) return false;
        *data = pp_.buf[cons_]; *fill = fill_[cons_]; *last = last_[cons_];
        return true;
    }
    void release()
    {
        { std::lock_guard<std::mutex> g(mu_); state_[cons_] = EMPTY; cons_ ^= 1; }
        cv_.notify_all();
    }

private:
    enum { EMPTY, FULL };
    void run()
    {
        size_t  seg = 0;
        gzFile  gz = nullptr;
        int     fd = -1;
        int64_t pos = 0, end = 0;
        int     k = 0;
        bool    done = false, open = false;
        uint8_t last_byte = '\n';
        const size_t CAP = pp_.cap;
        while (!done) {
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return state_[k] == EMPTY || stop_; });
                if (stop_) break;
            }
            size_t fill = 0;
            while (fill < CAP && !done) {
                if (!open) {
                    if (seg == segs_.size()) { done = true; break; }
                    if (join_ && seg > 0 && last_byte != '\n') { pp_.buf[k][fill++] = '\n'; last_byte = '\n'; continue; }
                    const Segment& sg = segs_[seg++];
                    if (sg.gz) {
                        gz = gzopen(sg.path.c_str(), "r");
                        if (!gz) { fail(); return; }
                        gzbuffer(gz, 4 << 20);
                    } else {
                        fd = ::open(sg.path.c_str(), O_RDONLY);
                        if (fd < 0) { fail(); return; }
                        pos = sg.off;
                        end = sg.len < 0 ? INT64_MAX : sg.off + sg.len;
                    }
                    open = true;
                    continue;
                }
                long got;
                if (gz) {
                    got = gzread(gz, pp_.buf[k] + fill, (unsigned)std::min<size_t>(CAP - fill, 1u << 30));
                } else {
                    const size_t want = (size_t)std::min<int64_t>((int64_t)(CAP - fill), end - pos);
                    got = want ? parallel_pread(fd, pp_.buf[k] + fill, want, pos) : 0;
                    if (got > 0) pos += got;
                }
                if (got < 0) { close_cur(gz, fd); fail(); return; }
                if (got == 0) { close_cur(gz, fd); open = false; continue; }
                fill += (size_t)got;
                last_byte = pp_.buf[k][fill - 1];
            }
            {
                std::lock_guard<std::mutex> g(mu_);
                fill_[k] = fill; last_[k] = done; state_[k] = FULL;
            }
            cv_.notify_all();
            k ^= 1;
        }
        close_cur(gz, fd);
    }
    static void close_cur(gzFile& gz, int& fd)
    {
        if (gz) { gzclose(gz); gz = nullptr; }
        if (fd >= 0) { ::close(fd); fd = -1; }
    }
    void fail()
    {
        { std::lock_guard<std::mutex> g(mu_); error_ = true; }
        cv_.notify_all();
    }
    PinnedPair&          pp_;
    std::vector<Segment> segs_;
    bool                 join_;
    size_t               fill_[2] = {0, 0};
    bool                 last_[2] = {false, false};
    int                  state_[2] = {EMPTY, EMPTY};
    int                  cons_ = 0;
    bool                 stop_ = false, error_ = false;
    std::mutex           mu_;
    std::condition_variable cv_;
    std::thread          th_;
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

// Writes the context's slice of one output stream at `file_base` of the open file (`file_total` bytes once every rank has
// written): this thread fetches windows into the two pinned buffers (D2H through the C ABI) while a writer thread drains the
// other buffer into the file.
static bool write_stream(raftgpu_ctx* ctx, int which, int fd, uint64_t file_base, uint64_t file_total, int ranks, PinnedPair& pp, std::string& err)
{
    uint64_t total = 0;
    if (raftgpu_output_size(ctx, which, &total)) { err = raftgpu_last_error(ctx); return false; }
    if (!total) return true;
    SliceWriter             out(fd, file_base, total, std::max(file_total, file_base + total), map_threads(ranks));
    const size_t            W = pp.cap;
    std::mutex              mu;
    std::condition_variable cv;
    size_t                  len[2] = {0, 0};
    uint64_t                at[2] = {0, 0};
    bool                    full[2] = {false, false}, done = false, werr = false;
    std::thread writer([&] {
        for (int k = 0;; k ^= 1) {
            std::unique_lock<std::mutex> g(mu);
            cv.wait(g, [&] { return full[k] || done; });
            if (!full[k]) return;
            g.unlock();
            if (!werr && !out.put(pp.buf[k], len[k], at[k])) werr = true;
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
        const int    st = raftgpu_fetch(ctx, which, off, pp.buf[k], n);
        if (st) { err = std::string(raftgpu_strerror(st)) + ": " + raftgpu_last_error(ctx); ok = false; break; }
        {
            std::lock_guard<std::mutex> g(mu);
            len[k] = n; at[k] = file_base + off; full[k] = true;
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
    if (werr && ok) { err = "short write"; ok = false; }
    return ok;
}

static const char* const OUT_SUFFIX[4] = {".coverage.txt", ".long_repeats.txt", ".long_repeats.bed", ".reads.fasta"};

static void print_run_lines(const raftgpu_params* p, int real_reads, int symmetric, long long n_records, int high_cov, long long total_cov,
                            long long n_bins, long long total_repeat_len, long long total_read_len)
{
    const int total_windows = (int)(uint32_t)(uint64_t)n_bins;           // an int in the reference (repeat.hpp:95,117): wraps past 2^31
    printf("Real Reads %d \n", real_reads);                              // chop.hpp:105
    printf("INFO, Symmetric overlaps %d \n", symmetric);                 // chop.hpp:189
    printf("INFO, length of alignments  %d()\n", (int)n_records);        // chop.hpp:190
    printf("high_cov %d\n", high_cov);                                   // repeat.hpp:91
    double coverage_per_window = (double)total_cov / total_windows;      // repeat.hpp:173-178
    double fraction_of_repeat_length = (double)total_repeat_len / total_read_len;
    printf("coverage per window is %f \n", coverage_per_window);
    printf("coverage per window/average coverage is %f \n", coverage_per_window / p->est_cov);
    printf("fraction_of_repeat_length %f \n", fraction_of_repeat_length);
}

struct Timer {
    bool on = getenv("RAFT_B200_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char* what, int rank = -1)
    {
        if (!on) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[raft_b200 timing] %s%s%s %.3f s\n", what, rank >= 0 ? " rank " : "", rank >= 0 ? std::to_string(rank).c_str() : "",
                std::chrono::duration<double>(now - t).count());
        t = now;
    }
};

// plain FASTA text on the device (K0); RAFTGPU_E_UNSUPPORTED when the text needs the host reader
static int ingest_fasta_segments(raftgpu_ctx* ctx, PinnedPair& pp, const std::vector<Segment>& segs, uint64_t total_hint)
{
    ByteSource src(pp, segs, false);
    for (;;) {
        const uint8_t* data = nullptr;
        size_t         fill = 0;
        bool           last = false;
        if (!src.next(&data, &fill, &last)) return RAFTGPU_E_IO;
        const int st = raftgpu_ingest_fasta(ctx, data, fill, last ? 1 : 0, total_hint);
        src.release();
        if (st) return st;
        if (last) return RAFTGPU_OK;
    }
}

extern "C" int raftgpu_break_long_reads(const char* readfilename, const char* paffilename, const raftgpu_params* p, const char* prefix,
                                        int device, raftgpu_stats* stats_out)
{
    return raftgpu_break_long_reads_multi(readfilename, 1, &paffilename, p, prefix, device, stats_out);
}

static int validate_inputs(const char* readfilename, int n_paf, const char* const* paffilenames, const std::string& pre)
{
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
    return RAFTGPU_OK;
}

extern "C" int raftgpu_break_long_reads_multi(const char* readfilename, int n_paf, const char* const* paffilenames, const raftgpu_params* p,
                                              const char* prefix, int device, raftgpu_stats* stats_out)
{
    if (!readfilename || n_paf < 1 || !paffilenames || !p || !prefix) return RAFTGPU_E_ARG;
    for (int k = 0; k < n_paf; k++) if (!paffilenames[k]) return RAFTGPU_E_ARG;
    const std::string pre(prefix);
    int st = validate_inputs(readfilename, n_paf, paffilenames, pre);
    if (st) return st;
    Timer tm;
    raftgpu_ctx* ctx = nullptr;
    st = raftgpu_create(p, device, &ctx);
    if (st) { fprintf(stderr, "raft_b200: %s\n", raftgpu_strerror(st)); return st; }
    auto fail = [&](int code) {
        fprintf(stderr, "raft_b200: %s: %s\n", raftgpu_strerror(code), raftgpu_last_error(ctx));
        raftgpu_destroy(ctx);
        return code;
    };
    int64_t biggest = file_size(readfilename);
    for (int k = 0; k < n_paf; k++) biggest = std::max(biggest, file_size(paffilenames[k]));
    PinnedPair pp((size_t)std::max<int64_t>(biggest, 0));
    if (!pp.ok()) return fail(RAFTGPU_E_NOMEM);
    tm.lap("context + pinned buffers");

    // reads: plain FASTA text is tokenised on the device; gzip, wrapped FASTQ and CR LF text go through the host reader
    bool on_device = false;
    if (!is_gzip_file(readfilename)) {
        st = ingest_fasta_segments(ctx, pp, {Segment{readfilename, 0, -1, false}}, (uint64_t)std::max<int64_t>(file_size(readfilename), 0));
        if (st == RAFTGPU_OK) on_device = true;
        else if (st != RAFTGPU_E_UNSUPPORTED) return fail(st);
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
    tm.lap("reads");

    // PAF: one reader thread inflates (gz, paf.hpp:24-38) or reads (plain) the files back to back into the two pinned
    // buffers while this thread hands the other buffer to the tokenizer, so disk/zlib time overlaps H2D + parse and
    // only the decoded records (25 B each) stay on the device: the text itself may exceed HBM.  Several files behave
    // like `cat a b | raft ...` (README.md:32-38: hifiasm writes two *.ovlp.paf files).
    {
        std::vector<Segment> segs;
        for (int k = 0; k < n_paf; k++) segs.push_back(Segment{paffilenames[k], 0, -1, is_gzip_file(paffilenames[k])});
        ByteSource src(pp, segs, true);
        for (;;) {
            const uint8_t* data = nullptr;
            size_t         fill = 0;
            bool           last = false;
            if (!src.next(&data, &fill, &last)) return fail(RAFTGPU_E_IO);
            st = raftgpu_ingest_paf(ctx, data, fill, last ? 1 : 0);
            src.release();
            if (st) return fail(st);
            if (last) break;
        }
    }
    tm.lap("paf");
    if ((st = raftgpu_run(ctx, &s))) return fail(st);
    tm.lap("run");
    print_run_lines(p, s.real_reads, s.symmetric, s.n_records, s.high_cov, s.total_cov, s.n_bins, s.total_repeat_len, s.total_read_len);

    std::string err;
    for (int w = 0; w < 4; w++) {
        const std::string path = pre + OUT_SUFFIX[w];
        int fd = ::open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
        bool ok = fd >= 0 && write_stream(ctx, w, fd, 0, 0, 1, pp, err);
        if (fd >= 0 && ::close(fd) != 0) ok = false;
        if (!ok) { fprintf(stderr, "raft_b200: cannot write %s: %s\n", path.c_str(), err.c_str()); raftgpu_destroy(ctx); return RAFTGPU_E_IO; }
        tm.lap(OUT_SUFFIX[w]);
    }
    if (stats_out) *stats_out = s;
    // A process that exits right after the call (the `raft` CLI) can leave the teardown of tens of GB of device buffers
    // to the driver's process cleanup instead of paying for it buffer by buffer (the reference frees nothing either).
    if (!getenv("RAFT_B200_NO_TEARDOWN")) raftgpu_destroy(ctx);
    tm.lap("destroy");
    return RAFTGPU_OK;
}

// ------------------------------------------------------------------------------------------------ several GPUs of one box
namespace {

struct Barrier { // reusable; wait() returns the OR of the flags passed by the threads of this round
    explicit Barrier(int n) : n_(n) {}
    bool wait(bool flag = false)
    {
        std::unique_lock<std::mutex> g(mu_);
        acc_ = acc_ || flag;
        const int gen = gen_;
        if (++count_ == n_) { result_ = acc_; acc_ = false; count_ = 0; gen_++; cv_.notify_all(); return result_; }
        cv_.wait(g, [&] { return gen_ != gen; });
        return result_;
    }
    int n_, count_ = 0, gen_ = 0;
    bool acc_ = false, result_ = false;
    std::mutex mu_;
    std::condition_variable cv_;
};

// first byte after the first '\n' at or after `pos` whose next byte satisfies `starts_record` (any byte when null); file size when none
int64_t next_line_start(int fd, int64_t pos, int64_t size, bool want_gt)
{
    std::vector<uint8_t> buf(1 << 20);
    if (pos <= 0) return 0;
    int64_t at = pos - 1; // a record may start exactly at pos: look at the byte before it
    bool    prev_nl = false;
    while (at < size) {
        const ssize_t got = pread(fd, buf.data(), buf.size(), (off_t)at);
        if (got <= 0) break;
        for (ssize_t i = 0; i < got; i++) {
            if (prev_nl && (!want_gt || buf[i] == '>')) return at + i;
            prev_nl = buf[i] == '\n';
        }
        at += got;
    }
    return size;
}

struct MgpuShared {
    int                  P = 0;
    std::vector<int>     devices;
    std::vector<raftgpu_ctx*> ctx;
    uint8_t              comm_id[RAFTGPU_COMM_ID_BYTES];
    std::vector<int>     status;
    // reads
    bool                 device_reads = false;
    std::vector<Segment> fasta_seg;
    std::vector<int64_t> n_local, name_bytes;
    std::vector<int64_t> bounds, name_base;
    std::vector<int64_t> lengths_all, name_off_all;
    std::vector<uint8_t> names_all;
    int64_t              hn = 0;
    int64_t *            hseq_off = nullptr, *hname_off = nullptr;
    uint8_t *            hseq = nullptr, *hnames = nullptr;
    // PAF
    std::vector<std::vector<Segment>> paf_seg;
    // results
    std::vector<raftgpu_shard_info> info;
    std::vector<raftgpu_stats>      stats;
    int                  out_fd[4] = {-1, -1, -1, -1};
    std::string          err;
    std::mutex           mu;
};

} // namespace

extern "C" int raftgpu_break_long_reads_mgpu(const char* readfilename, int n_paf, const char* const* paffilenames, const raftgpu_params* p,
                                             const char* prefix, int ndev, const int* devices, raftgpu_stats* stats_out)
{
    if (!readfilename || n_paf < 1 || !paffilenames || !p || !prefix || ndev < 1 || ndev > 64 || !devices) return RAFTGPU_E_ARG;
    if (ndev == 1) return raftgpu_break_long_reads_multi(readfilename, n_paf, paffilenames, p, prefix, devices[0], stats_out);
    for (int k = 0; k < n_paf; k++) if (!paffilenames[k]) return RAFTGPU_E_ARG;
    const std::string pre(prefix);
    int st = validate_inputs(readfilename, n_paf, paffilenames, pre);
    if (st) return st;
    Timer      tm;
    MgpuShared sh;
    const int  P = ndev;
    sh.P = P; sh.devices.assign(devices, devices + P);
    sh.ctx.assign(P, nullptr); sh.status.assign(P, 0); sh.n_local.assign(P, 0); sh.name_bytes.assign(P, 0);
    sh.info.assign(P, raftgpu_shard_info{}); sh.stats.assign(P, raftgpu_stats{}); sh.paf_seg.assign(P, {});
    auto destroy_all = [&] { for (auto& c : sh.ctx) if (c) { raftgpu_destroy(c); c = nullptr; } };
    {   // one CUDA context per device: created side by side (each costs some tenths of a second)
        std::vector<std::thread> mk;
        for (int r = 0; r < P; r++) mk.emplace_back([&, r] { sh.status[r] = raftgpu_create(p, devices[r], &sh.ctx[r]); });
        for (auto& t : mk) t.join();
        for (int r = 0; r < P; r++)
            if ((st = sh.status[r])) { fprintf(stderr, "raft_b200: device %d: %s\n", devices[r], raftgpu_strerror(st)); destroy_all(); return st; }
    }
    if ((st = raftgpu_comm_unique_id(sh.comm_id))) { destroy_all(); return st; }

    // ---- how the inputs are cut.  Reads: plain FASTA is cut into P byte ranges at record starts ('>' opening a line), each
    // tokenised on its own GPU, and that cut IS the read partition (bytes ~ bases ~ coverage slots); anything else (gzip,
    // FASTQ, CR LF) is read once by the host reader and partitioned by coverage slots.  PAF: every plain file is cut
    // into P byte ranges at line ends; a gzip file cannot be cut and goes to one rank as a whole.
    const int64_t fsize = file_size(readfilename);
    {
        int fd = ::open(readfilename, O_RDONLY);
        uint8_t first = 0;
        if (fd >= 0 && !is_gzip_file(readfilename) && pread(fd, &first, 1, 0) == 1 && first == '>') {
            sh.device_reads = true;
            std::vector<int64_t> cut(P + 1, fsize);
            cut[0] = 0;
            for (int r = 1; r < P; r++) cut[r] = std::max(cut[r - 1], next_line_start(fd, fsize / P * r, fsize, true));
            for (int r = 0; r < P; r++) sh.fasta_seg.push_back(Segment{readfilename, cut[r], cut[r + 1] - cut[r], false});
        }
        if (fd >= 0) ::close(fd);
    }
    int64_t biggest = sh.device_reads ? fsize / P + 1 : 0;
    for (int k = 0; k < n_paf; k++) {
        const int64_t sz = file_size(paffilenames[k]);
        if (is_gzip_file(paffilenames[k])) { sh.paf_seg[k % P].push_back(Segment{paffilenames[k], 0, -1, true}); biggest = std::max<int64_t>(biggest, 64 << 20); continue; }
        int fd = ::open(paffilenames[k], O_RDONLY);
        if (fd < 0) { destroy_all(); return RAFTGPU_E_IO; }
        std::vector<int64_t> cut(P + 1, sz);
        cut[0] = 0;
        for (int r = 1; r < P; r++) cut[r] = std::max(cut[r - 1], next_line_start(fd, sz / P * r, sz, false));
        ::close(fd);
        for (int r = 0; r < P; r++) if (cut[r + 1] > cut[r]) sh.paf_seg[r].push_back(Segment{paffilenames[k], cut[r], cut[r + 1] - cut[r], false});
        biggest = std::max(biggest, sz / P + 1);
    }
    for (int w = 0; w < 4; w++) {
        sh.out_fd[w] = ::open((pre + OUT_SUFFIX[w]).c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
        if (sh.out_fd[w] < 0) { fprintf(stderr, "raft_b200: cannot open %s%s\n", pre.c_str(), OUT_SUFFIX[w]); destroy_all(); return RAFTGPU_E_IO; }
    }
    tm.lap("contexts + input cuts");

    // The communicator is built first, by one short-lived thread per rank, with stdout pointed at stderr meanwhile: NCCL
    // prints its version banner on stdout at the first initialisation (NCCL_DEBUG=VERSION), and the reference's stdout
    // lines are part of the drop-in contract.
    {
        fflush(stdout);
        const int saved = dup(1);
        if (saved >= 0) dup2(2, 1);
        std::vector<std::thread> ci;
        for (int r = 0; r < P; r++) ci.emplace_back([&, r] { sh.status[r] = raftgpu_comm_init(sh.ctx[r], P, r, sh.comm_id); });
        for (auto& t : ci) t.join();
        fflush(stdout);
        if (saved >= 0) { dup2(saved, 1); close(saved); }
        for (int r = 0; r < P; r++)
            if (sh.status[r]) {
                fprintf(stderr, "raft_b200: device %d: %s: %s\n", devices[r], raftgpu_strerror(sh.status[r]), raftgpu_last_error(sh.ctx[r]));
                st = sh.status[r];
                for (int w = 0; w < 4; w++) ::close(sh.out_fd[w]);
                destroy_all();
                return st;
            }
    }
    tm.lap("communicator");

    Barrier bar(P);
    auto    body = [&](int r) {
        raftgpu_ctx* ctx = sh.ctx[r];
        int&         rc = sh.status[r];
        Timer        tr;
        PinnedPair   pp((size_t)std::max<int64_t>(biggest, 0));
        rc = pp.ok() ? RAFTGPU_OK : RAFTGPU_E_NOMEM;
        if (bar.wait(rc != 0)) return;
        // ---- reads
        if (sh.device_reads) {
            rc = sh.fasta_seg[r].len > 0 ? ingest_fasta_segments(ctx, pp, {sh.fasta_seg[r]}, (uint64_t)sh.fasta_seg[r].len)
                                         : raftgpu_ingest_fasta(ctx, nullptr, 0, 1, 0);
            if (rc == RAFTGPU_OK) rc = raftgpu_reads_info(ctx, &sh.n_local[r], &sh.name_bytes[r], nullptr);
        }
        bool bad = sh.device_reads && rc != RAFTGPU_OK && rc != RAFTGPU_E_UNSUPPORTED;
        if (bar.wait(bad)) return;
        const bool fallback = bar.wait(sh.device_reads && rc == RAFTGPU_E_UNSUPPORTED) || !sh.device_reads;
        rc = RAFTGPU_OK;
        if (!fallback) {
            if (r == 0) { // global read numbering = file order: prefix sums over the ranks' byte ranges
                sh.bounds.assign(P + 1, 0); sh.name_base.assign(P + 1, 0);
                for (int q = 0; q < P; q++) { sh.bounds[q + 1] = sh.bounds[q] + sh.n_local[q]; sh.name_base[q + 1] = sh.name_base[q] + sh.name_bytes[q]; }
                sh.lengths_all.resize((size_t)sh.bounds[P] + 1); sh.name_off_all.resize((size_t)sh.bounds[P] + 1); sh.names_all.resize((size_t)sh.name_base[P] + 1);
                sh.hn = sh.bounds[P];
            }
            bar.wait();
            std::vector<int64_t> no((size_t)sh.n_local[r] + 1);
            rc = raftgpu_reads_copy(ctx, no.data(), sh.names_all.data() + sh.name_base[r], sh.lengths_all.data() + sh.bounds[r]);
            for (int64_t i = 0; i < sh.n_local[r]; i++) sh.name_off_all[(size_t)(sh.bounds[r] + i)] = sh.name_base[r] + no[(size_t)i];
            if (r == P - 1) sh.name_off_all[(size_t)sh.bounds[P]] = sh.name_base[P];
            if (bar.wait(rc != 0)) return;
            const int64_t* d_seq_off = nullptr;
            const uint8_t* d_seq = nullptr;
            rc = raftgpu_reads_device(ctx, &d_seq_off, &d_seq);
            if (rc == RAFTGPU_OK)
                rc = raftgpu_set_reads_sharded(ctx, sh.hn, sh.lengths_all.data(), sh.name_off_all.data(), sh.names_all.data(), sh.bounds[r],
                                               sh.n_local[r], d_seq_off, d_seq);
        } else {
            if (r == 0) {
                rc = raftgpu_load_fasta(readfilename, &sh.hn, &sh.hseq_off, &sh.hseq, &sh.hname_off, &sh.hnames);
                if (rc == RAFTGPU_OK) {
                    sh.lengths_all.resize((size_t)sh.hn + 1);
                    for (int64_t i = 0; i < sh.hn; i++) sh.lengths_all[(size_t)i] = sh.hseq_off[i + 1] - sh.hseq_off[i];
                    sh.bounds.assign(P + 1, 0);
                    rc = raftgpu_partition_reads(sh.lengths_all.data(), sh.hn, p->reso, P, sh.bounds.data());
                }
            }
            if (bar.wait(rc != 0)) return;
            const int64_t b0 = sh.bounds[r], b1 = sh.bounds[r + 1];
            std::vector<int64_t> own_off((size_t)(b1 - b0) + 1);
            for (int64_t i = b0; i <= b1; i++) own_off[(size_t)(i - b0)] = sh.hseq_off[i] - sh.hseq_off[b0];
            rc = raftgpu_set_reads_sharded(ctx, sh.hn, sh.lengths_all.data(), sh.hname_off, sh.hnames, b0, b1 - b0, own_off.data(), sh.hseq + sh.hseq_off[b0]);
        }
        if (bar.wait(rc != 0)) return;
        tr.lap("reads", r);
        // ---- PAF: the rank's segments stream through the pinned buffers; the first chunk carries the record-0 protocol
        {
            ByteSource src(pp, sh.paf_seg[r], true);
            bool       begun = false, skip = false;
            for (;;) {
                const uint8_t* data = nullptr;
                size_t         fill = 0;
                bool           last = false;
                if (!src.next(&data, &fill, &last)) { rc = RAFTGPU_E_IO; break; }
                if (!begun) { // record 0 of the file is looked for in the first megabytes of the first chunk
                    const size_t head = std::min<size_t>(fill, 4u << 20);
                    rc = raftgpu_sharded_begin(ctx, sh.bounds.data(), data, head, last && head == fill ? 1 : 0);
                    begun = true;
                    if (rc) { src.release(); break; }
                }
                if (!skip) {
                    const int st2 = raftgpu_ingest_paf(ctx, data, fill, last ? 1 : 0);
                    if (st2 == RAFTGPU_E_UNKNOWN_NAME || st2 == RAFTGPU_E_RANGE) skip = true; // carried into the finish below: the ranks must still meet
                    else if (st2) rc = st2;
                }
                src.release();
                if (rc || last) break;
            }
        }
        // a rank that failed outside the data domain cannot take part in the collectives: stop everybody before them
        if (bar.wait(rc != 0)) return;
        tr.lap("paf", r);
        rc = raftgpu_sharded_finish(ctx, &sh.stats[r], &sh.info[r]);
        if (bar.wait(rc != 0)) return;
        tr.lap("exchange + run", r);
        // ---- outputs: this rank's slice of every file, at its file offset
        std::string err;
        for (int w = 0; w < 4 && rc == RAFTGPU_OK; w++)
            if (!write_stream(ctx, w, sh.out_fd[w], sh.info[r].stream_base[w], sh.info[r].stream_total[w], P, pp, err)) {
                std::lock_guard<std::mutex> g(sh.mu);
                sh.err = err; rc = RAFTGPU_E_IO;
            }
        tr.lap("outputs", r);
    };
    std::vector<std::thread> th;
    for (int r = 0; r < P; r++) th.emplace_back(body, r);
    for (auto& t : th) t.join();
    for (int w = 0; w < 4; w++) if (::close(sh.out_fd[w]) != 0 && !sh.status[0]) sh.status[0] = RAFTGPU_E_IO;
    free(sh.hseq_off); free(sh.hseq); free(sh.hname_off); free(sh.hnames);
    // the rank with a real error speaks; RAFTGPU_E_PEER only if nobody else does
    int rc = RAFTGPU_OK, who = -1;
    for (int r = 0; r < P; r++) if (sh.status[r] && (rc == RAFTGPU_OK || rc == RAFTGPU_E_PEER) ) { rc = sh.status[r]; who = r; }
    if (rc) {
        fprintf(stderr, "raft_b200: %s: %s%s\n", raftgpu_strerror(rc), who >= 0 && sh.ctx[who] ? raftgpu_last_error(sh.ctx[who]) : "", sh.err.c_str());
        destroy_all();
        return rc;
    }
    const raftgpu_shard_info& i0 = sh.info[0];
    print_run_lines(p, sh.stats[0].real_reads, i0.symmetric, i0.n_records_total, sh.stats[0].high_cov, i0.total_cov, i0.n_bins_total,
                    i0.total_repeat_len, i0.total_read_len);
    if (stats_out) {
        raftgpu_stats s = sh.stats[0];
        s.n_records = i0.n_records_total; s.symmetric = i0.symmetric; s.total_cov = i0.total_cov; s.total_repeat_len = i0.total_repeat_len;
        s.total_read_len = i0.total_read_len; s.n_bins = i0.n_bins_total; s.total_windows = (int32_t)(uint32_t)(uint64_t)i0.n_bins_total;
        s.n_fragments = i0.n_fragments_total; s.n_repeats = 0;
        for (int r = 0; r < P; r++) s.n_repeats += sh.stats[r].n_repeats;
        for (int w = 0; w < 4; w++) s.out_bytes[w] = i0.stream_total[w];
        *stats_out = s;
    }
    if (!getenv("RAFT_B200_NO_TEARDOWN")) destroy_all();
    tm.lap("all ranks");
    return RAFTGPU_OK;
}
