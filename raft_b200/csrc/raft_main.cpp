// raft_main.cpp — the `raft` command line on top of libraft_b200.so.
// Same flags, defaults, quirks, log lines and exit codes as the reference's main.cpp:7-87; the work
// itself is raftgpu_break_long_reads (the drop-in for break_long_reads, chop.hpp:331).
#include <getopt.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/raft_b200.h"

static void print_help(const raftgpu_params& p, const std::string& prefix)
{ // main.cpp:7-19 (exit code 1 included)
    std::cout << "Usage: raft [options] <input-reads.fa> <in.paf>\n";
    std::cout << "  -r NUM     resolution of coverage " << p.reso << "\n";
    std::cout << "  -e NUM     estimated coverage " << "\n";
    std::cout << "  -m NUM     coverage multiplier " << p.cov_mul << "\n";
    std::cout << "  -l NUM     read_length " << p.read_length << "\n";
    std::cout << "  -v NUM     overlap_length " << p.overlap_length << "\n";
    std::cout << "  -p NUM     repeat_length " << p.repeat_length << "\n";
    std::cout << "  -f NUM     flanking_length " << p.flanking_length << "\n";
    std::cout << "  -o FILE    prefix of output files " << prefix << "\n";
    exit(1);
}

int main(int argc, char* argv[])
{
    raftgpu_params p;
    raftgpu_default_params(&p);
    std::string prefix = "raft"; // param.hpp:28
    int         option;
    while ((option = getopt(argc, argv, "r:e:m:l:i:p:f:v:o:")) != -1) {
        switch (option) {
        case 'r': p.reso = atoi(optarg); break;
        case 'e': p.est_cov = atoi(optarg); break;
        case 'm': p.cov_mul = std::stod(optarg); break;
        case 'l': p.read_length = atoi(optarg); break;
        case 'p': p.repeat_length = atoi(optarg); p.interval_length = atoi(optarg); break; // main.cpp:44-47
        case 'f': p.flanking_length = atoi(optarg); break;
        case 'v': p.overlap_length = atoi(optarg); // main.cpp:51-55: no break, -v also sets the prefix
        /* fall through */
        case 'o': prefix = optarg; break;
        default: print_help(p, prefix); // includes -i, accepted by the optstring but unhandled (main.cpp:28,56-57)
        }
    }
    if (argc < optind + 2) print_help(p, prefix);
    if (p.est_cov <= 0) {
        std::cout << "ERROR, main(), estimated coverage must be set properly\n";
        print_help(p, prefix);
    }
    // param.hpp:33-43
    std::cout << "INFO, printParams(), reso = " << p.reso << "\n";
    std::cout << "INFO, printParams(), est_cov = " << p.est_cov << "\n";
    std::cout << "INFO, printParams(), cov_mul = " << p.cov_mul << "\n";
    std::cout << "INFO, printParams(), repeat_length = " << p.repeat_length << "\n";
    std::cout << "INFO, printParams(), interval_length = " << p.interval_length << "\n";
    std::cout << "INFO, printParams(), read_length = " << p.read_length << "\n";
    std::cout << "INFO, printParams(), overlap_length = " << p.overlap_length << "\n";
    std::cout << "INFO, printParams(), flanking_length = " << p.flanking_length << "\n";

    auto t0 = std::chrono::system_clock::now();
    std::cout << "INFO, main(), started timer\n";
    std::cout.flush();

    const char* dev_env = getenv("RAFT_B200_DEVICE");
    // RAFT_B200_DEVICES=0,1,...,7: the run is sharded over these GPUs (reads by id range, PAF by byte range, NCCL exchange)
    std::vector<int> devs;
    if (const char* e = getenv("RAFT_B200_DEVICES")) {
        for (const char* q = e; *q;) {
            char* end = nullptr;
            long  v = strtol(q, &end, 10);
            if (end == q) break;
            devs.push_back((int)v);
            q = *end == ',' ? end + 1 : end;
        }
    }
    // main.cpp:75 passes argv[optind+2] on and break_long_reads ignores it (chop.hpp:331): so do we.  Opt-in extension:
    // with RAFT_B200_MULTI_PAF=1 every further positional argument is one more PAF file, ingested back to back as `cat`
    // would join them (README.md:35-36 merges hifiasm's two *.ovlp.paf files before calling raft).
    const char* multi = getenv("RAFT_B200_MULTI_PAF");
    const int   n_paf = (multi && atoi(multi) != 0) ? argc - optind - 1 : 1;
    // NCCL (sharded runs) writes its debug lines to stdout by default: the reference's stdout lines are part of the drop-in
    // contract, so they go to stderr unless the user chose a file
    if (devs.size() > 1) setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
    // CUDA initialises every visible GPU of the box: expose only the ones this run uses (about 0.2 s per spared device)
    if (!getenv("CUDA_VISIBLE_DEVICES")) {
        if (devs.empty()) devs.push_back(dev_env ? atoi(dev_env) : 0);
        std::string vis;
        for (size_t k = 0; k < devs.size(); k++) { vis += (k ? "," : "") + std::to_string(devs[k]); devs[k] = (int)k; }
        setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 1);
    }
    setenv("RAFT_B200_NO_TEARDOWN", "1", 0); // this process ends right after the call: see host_io.cpp
    int         st = devs.size() > 1 ? raftgpu_break_long_reads_mgpu(argv[optind], n_paf, argv + optind + 1, &p, prefix.c_str(), (int)devs.size(),
                                                                     devs.data(), nullptr)
                                     : raftgpu_break_long_reads_multi(argv[optind], n_paf, argv + optind + 1, &p, prefix.c_str(),
                                                                      devs.size() == 1 ? devs[0] : (dev_env ? atoi(dev_env) : 0), nullptr);
    fflush(stdout);
    if (st == RAFTGPU_E_IO) return 1; // chop.hpp:339-348 exit(1)
    if (st != RAFTGPU_OK) return 2;   // inputs on which the reference crashes (segfault / SIGFPE / throw) or CUDA failure

    std::chrono::duration<double> wct = std::chrono::system_clock::now() - t0;
    std::cout << "INFO, main(), program completed after " << wct.count() << " seconds\n";
    fprintf(stdout, "INFO, %s(), CMD:", __func__);
    for (int i = 0; i < argc; ++i) fprintf(stdout, " %s", argv[i]);
    std::cout << "\n";
    std::cout.flush();
    fflush(stdout);
    _exit(0); // skip the CUDA runtime's exit-time teardown: the driver reclaims the context with the process
}
