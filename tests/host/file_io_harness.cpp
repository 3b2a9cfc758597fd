// CPU harness for raft_b200/csrc/file_io.h (compiled and run by tests/test_host_io.py):
//   file_io_harness <scratch dir>
// prints one "name ok|FAIL detail" line per case; exit code 0 iff every case passed.
#include <cstdio>
#include <string>

#include "../../raft_b200/csrc/file_io.h"

using namespace raftio;

static uint8_t byte_at(uint64_t i) { return (uint8_t)((i * 2654435761u) >> 13); }
static int     failures = 0;
static void    report(const char* name, bool ok, const std::string& detail = "")
{
    printf("%s %s %s\n", name, ok ? "ok" : "FAIL", detail.c_str());
    if (!ok) failures++;
}

static bool file_matches(const std::string& path, uint64_t total)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    std::vector<uint8_t> all(total + 16);
    const size_t         got = fread(all.data(), 1, all.size(), f);
    fclose(f);
    if (got != total) return false;
    for (uint64_t i = 0; i < total; i++) if (all[i] != byte_at(i)) return false;
    return true;
}

// three ranks write ragged slices of one file concurrently, in windows that are not multiples of anything
static void slices(const std::string& path, bool want_map, const char* name)
{
    const uint64_t sizes[3] = {50000123, 1, (32u << 20) + 4097};
    uint64_t       base[4] = {0, 0, 0, 0};
    for (int r = 0; r < 3; r++) base[r + 1] = base[r] + sizes[r];
    const uint64_t total = base[3];
    if (want_map) unsetenv("RAFT_B200_NO_MMAP"); else setenv("RAFT_B200_NO_MMAP", "1", 1);
    int  fd = open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
    bool mapped[3] = {false, false, false}, put_ok[3] = {true, true, true};
    std::vector<std::thread> th;
    for (int r = 2; r >= 0; r--)
        th.emplace_back([&, r] {
            SliceWriter          w(fd, base[r], sizes[r], total, 3);
            mapped[r] = w.mapped();
            const size_t         W = (20u << 20) + 5;
            std::vector<uint8_t> buf(W);
            for (uint64_t off = 0; off < sizes[r]; off += W) {
                const size_t n = (size_t)std::min<uint64_t>(W, sizes[r] - off);
                for (size_t i = 0; i < n; i++) buf[i] = byte_at(base[r] + off + i);
                if (!w.put(buf.data(), n, base[r] + off)) put_ok[r] = false;
            }
        });
    for (auto& t : th) t.join();
    close(fd);
    const bool modes = mapped[0] == want_map && mapped[1] == want_map && mapped[2] == want_map;
    report(name, modes && put_ok[0] && put_ok[1] && put_ok[2] && file_matches(path, total),
           std::string("mapped=") + (mapped[0] ? "1" : "0") + (mapped[1] ? "1" : "0") + (mapped[2] ? "1" : "0"));
    unlink(path.c_str());
    unsetenv("RAFT_B200_NO_MMAP");
}

// the file never shrinks: a rank that arrives late with a smaller `file_total` must not cut what another rank mapped
static void grow_only(const std::string& path)
{
    int fd = open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
    std::vector<uint8_t> a(9000), b(5000);
    for (size_t i = 0; i < a.size(); i++) a[i] = byte_at(5000 + i);
    for (size_t i = 0; i < b.size(); i++) b[i] = byte_at(i);
    bool ok = true;
    {
        SliceWriter hi(fd, 5000, 9000, 14000, 2);
        SliceWriter lo(fd, 0, 5000, 5000, 2); // wrong (short) total
        ok = hi.put(a.data(), a.size(), 5000) && lo.put(b.data(), b.size(), 0);
    }
    close(fd);
    report("grow_only", ok && file_matches(path, 14000));
    unlink(path.c_str());
}

static void not_a_regular_file()
{
    int fd = open("/dev/null", O_RDWR);
    SliceWriter w(fd, 0, 1 << 20, 1 << 20, 4);
    std::vector<uint8_t> buf(1 << 20, 7);
    report("dev_null", !w.mapped() && w.put(buf.data(), buf.size(), 0));
    close(fd);
}

static void empty_slice(const std::string& path)
{
    int fd = open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
    { SliceWriter w(fd, 0, 0, 0, 4); report("empty_slice", !w.mapped()); }
    struct stat sb;
    fstat(fd, &sb);
    report("empty_slice_size", sb.st_size == 0);
    close(fd);
    unlink(path.c_str());
}

// sliced pread: whole windows, a window that ends past the end of the file, a window starting at the end
static void preads(const std::string& path)
{
    const uint64_t total = (70u << 20) + 123;
    {
        FILE* f = fopen(path.c_str(), "wb");
        std::vector<uint8_t> buf(1 << 20);
        for (uint64_t off = 0; off < total; off += buf.size()) {
            const size_t n = (size_t)std::min<uint64_t>(buf.size(), total - off);
            for (size_t i = 0; i < n; i++) buf[i] = byte_at(off + i);
            fwrite(buf.data(), 1, n, f);
        }
        fclose(f);
    }
    int fd = open(path.c_str(), O_RDONLY);
    std::vector<uint8_t> dst(64u << 20);
    auto check = [&](int64_t pos, size_t want, long expect) {
        const long got = parallel_pread(fd, dst.data(), want, pos);
        if (got != expect) return false;
        for (long i = 0; i < got; i++) if (dst[(size_t)i] != byte_at((uint64_t)pos + (uint64_t)i)) return false;
        return true;
    };
    report("pread_whole_window", check(0, 64u << 20, 64l << 20));
    report("pread_unaligned", check(12345, (40u << 20) + 7, (40l << 20) + 7));
    report("pread_short_at_eof", check(64ll << 20, 64u << 20, (6l << 20) + 123));
    report("pread_eof_in_first_slice", check((70ll << 20) + 100, 64u << 20, 23));
    report("pread_at_eof", check((int64_t)total, 64u << 20, 0));
    report("pread_small", check(5, 1000, 1000));
    close(fd);
    unlink(path.c_str());
}

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    const std::string dir = argv[1];
    slices(dir + "/slices_map.bin", true, "slices_mapped");
    slices(dir + "/slices_pw.bin", false, "slices_pwrite");
    grow_only(dir + "/grow.bin");
    not_a_regular_file();
    empty_slice(dir + "/empty.bin");
    preads(dir + "/pread.bin");
    report("map_threads_range", map_threads(1) >= 2 && map_threads(1) <= 8 && map_threads(64) == 2);
    setenv("RAFT_B200_IO_THREADS", "5", 1);
    report("map_threads_env", map_threads(3) == 5);
    return failures ? 1 : 0;
}
