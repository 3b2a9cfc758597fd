// file_io.h — plain-file helpers of the host side (no CUDA): sliced pread / pwrite and the mapped writer of an output
// slice.  Header-only so that tests/test_host_io.py can compile them into a CPU harness.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/statvfs.h>
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace raftio {

// Page-cache / tmpfs copies run at a few GB/s per thread: large transfers are cut into slices moved by a few threads.
constexpr size_t IO_SLICE = 16u << 20;
constexpr int    IO_THREADS = 4;
// pread of [pos, pos+want) into dst; returns the bytes read (short only at the end of the file), -1 on error
inline long parallel_pread(int fd, uint8_t* dst, size_t want, int64_t pos)
{
    if (want < 2 * IO_SLICE) {
        size_t got = 0;
        while (got < want) {
            ssize_t r = pread(fd, dst + got, want - got, (off_t)(pos + (int64_t)got));
            if (r < 0) return -1;
            if (r == 0) break;
            got += (size_t)r;
        }
        return (long)got;
    }
    const int    T = (int)std::min<size_t>(IO_THREADS, want / IO_SLICE);
    const size_t per = ((want + T - 1) / T + 4095) & ~(size_t)4095;
    std::vector<long>        res(T, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            const size_t lo = std::min(want, per * t), hi = std::min(want, per * (t + 1));
            size_t       got = 0;
            while (lo + got < hi) {
                ssize_t r = pread(fd, dst + lo + got, hi - lo - got, (off_t)(pos + (int64_t)(lo + got)));
                if (r < 0) { res[t] = -1; return; }
                if (r == 0) break;
                got += (size_t)r;
            }
            res[t] = (long)got;
        });
    for (auto& x : th) x.join();
    long total = 0;
    for (int t = 0; t < T; t++) {
        if (res[t] < 0) return -1;
        total += res[t];
        if ((size_t)res[t] < std::min(want, per * (t + 1)) - std::min(want, per * t)) break; // end of file inside this slice
    }
    return total;
}
inline bool parallel_pwrite(int fd, const uint8_t* src, size_t n, uint64_t at)
{
    auto put = [&](size_t lo, size_t hi) {
        while (lo < hi) {
            ssize_t w = pwrite(fd, src + lo, hi - lo, (off_t)(at + lo));
            if (w <= 0) return false;
            lo += (size_t)w;
        }
        return true;
    };
    if (n < 2 * IO_SLICE) return put(0, n);
    const int    T = (int)std::min<size_t>(IO_THREADS, n / IO_SLICE);
    const size_t per = ((n + T - 1) / T + 4095) & ~(size_t)4095;
    std::vector<char>        ok(T, 1);
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back([&, t] { ok[t] = put(std::min(n, per * t), std::min(n, per * (t + 1))) ? 1 : 0; });
    for (auto& x : th) x.join();
    for (int t = 0; t < T; t++) if (!ok[t]) return false;
    return true;
}

// threads that fill a mapped output slice: the cores of the box shared out among the ranks, 2..8 each
inline int map_threads(int ranks)
{
    if (const char* e = getenv("RAFT_B200_IO_THREADS")) return std::max(1, std::min(64, atoi(e)));
    const int hw = (int)std::thread::hardware_concurrency();
    return std::max(2, std::min(8, (hw > 0 ? hw : 8) / std::max(ranks, 1)));
}

// Destination of one rank's slice of an output file: bytes [base, base + total) of a file that will hold `file_total`.
// The page-cache copy of pwrite runs under the file's inode lock, so several threads writing one file take turns
// (≈2 GB/s whatever their number); page faults on a shared mapping of the same file do not, so the slice is mapped and
// filled with memcpy by a few threads.  pwrite stays as the fallback: when the file cannot be mapped, and when the volume is
// short of space (a fault on a full volume is a SIGBUS, a failed pwrite is an error code the caller can report).
class SliceWriter {
public:
    SliceWriter(int fd, uint64_t base, uint64_t total, uint64_t file_total, int threads) : fd_(fd), threads_(threads)
    {
        if (!total || getenv("RAFT_B200_NO_MMAP")) return;
        struct stat sb;
        struct statvfs vfs;
        if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) return;
        if (fstatvfs(fd, &vfs) != 0 || (uint64_t)vfs.f_bavail * vfs.f_frsize < total + (256ull << 20)) return;
        {   // grow only: the ranks of a run reach this point in any order
            static std::mutex           grow;
            std::lock_guard<std::mutex> g(grow);
            if (fstat(fd, &sb) != 0) return;
            if ((uint64_t)sb.st_size < file_total && ftruncate(fd, (off_t)file_total) != 0) return;
        }
        const long page = sysconf(_SC_PAGESIZE);
        lo_ = base & ~(uint64_t)(page - 1);
        len_ = (size_t)(base + total - lo_);
        void* m = mmap(nullptr, len_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, (off_t)lo_);
        if (m != MAP_FAILED) map_ = (uint8_t*)m;
    }
    ~SliceWriter() { if (map_) munmap(map_, len_); }
    SliceWriter(const SliceWriter&) = delete;
    SliceWriter& operator=(const SliceWriter&) = delete;
    bool mapped() const { return map_ != nullptr; }
    // n bytes at file offset `at` (inside the slice)
    bool put(const uint8_t* src, size_t n, uint64_t at)
    {
        if (!map_) return parallel_pwrite(fd_, src, n, at);
        uint8_t*     dst = map_ + (at - lo_);
        const size_t MIN_SLICE = 4u << 20;
        const int    T = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads_, n / MIN_SLICE));
        if (T == 1) { memcpy(dst, src, n); return true; }
        // cuts on destination page boundaries, so that no two threads fault the same page
        const size_t head = (size_t)(-(intptr_t)dst & 4095);
        const size_t per = (((n - head) + T - 1) / T + 4095) & ~(size_t)4095;
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) {
            const size_t a = t ? std::min(n, head + per * t) : 0, b = std::min(n, head + per * (t + 1));
            if (b > a) th.emplace_back([=] { memcpy(dst + a, src + a, b - a); });
        }
        for (auto& x : th) x.join();
        return true;
    }

private:
    int      fd_;
    int      threads_;
    uint8_t* map_ = nullptr;
    uint64_t lo_ = 0;
    size_t   len_ = 0;
};

} // namespace raftio
