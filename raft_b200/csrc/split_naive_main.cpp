// split_naive_main.cpp — the reference's `split_naive` tool (split_naive.cpp:46-62) on top of libraft_b200.so.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/raft_b200.h"

static void print_help()
{ // split_naive.cpp:46-52
    std::cout << "Purpose: Split input reads naively into non-overlapping subreads. The output format is FASTA\n";
    std::cout << "Usage: split_naive <inputfilename> <outputfilename> SPLITLEN\n";
    std::cout << "Example: split_naive input.fastq output.fragmented.fasta 20000\n";
    exit(1);
}

int main(int argc, char** argv)
{
    if (argc < 4) print_help();
    const int      sublen = std::stoi(argv[3]);
    raftgpu_params p;
    raftgpu_default_params(&p);
    p.est_cov = 1;
    raftgpu_ctx* ctx = nullptr;
    int          st = raftgpu_create(&p, getenv("RAFT_B200_DEVICE") ? atoi(getenv("RAFT_B200_DEVICE")) : 0, &ctx);
    if (st) { fprintf(stderr, "split_naive: %s\n", raftgpu_strerror(st)); return 2; }
    int64_t  n = 0;
    int64_t *seq_off = nullptr, *name_off = nullptr;
    uint8_t *seq = nullptr, *names = nullptr;
    FILE*    out = fopen(argv[2], "wb"); // the reference opens the output before reading (split_naive.cpp:20)
    if ((st = raftgpu_load_fasta(argv[1], &n, &seq_off, &seq, &name_off, &names)) == RAFTGPU_OK) st = raftgpu_set_reads(ctx, n, seq_off, seq, name_off, names);
    if (st == RAFTGPU_OK) st = raftgpu_split_naive(ctx, sublen);
    uint64_t total = 0;
    if (st == RAFTGPU_OK) st = raftgpu_output_size(ctx, RAFTGPU_OUT_SPLIT_NAIVE, &total);
    std::vector<uint8_t> buf((size_t)std::min<uint64_t>(total ? total : 1, 256u << 20));
    for (uint64_t off = 0; st == RAFTGPU_OK && out && off < total; off += buf.size()) {
        size_t len = (size_t)std::min<uint64_t>(buf.size(), total - off);
        st = raftgpu_fetch(ctx, RAFTGPU_OUT_SPLIT_NAIVE, off, buf.data(), len);
        if (st == RAFTGPU_OK && fwrite(buf.data(), 1, len, out) != len) st = RAFTGPU_E_IO;
    }
    if (out) fclose(out);
    if (st) fprintf(stderr, "split_naive: %s: %s\n", raftgpu_strerror(st), raftgpu_last_error(ctx));
    raftgpu_destroy(ctx);
    return st ? 2 : 0;
}
