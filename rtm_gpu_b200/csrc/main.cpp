// rtm_b200: drop-in replacement of the reference's RTM executable for the hot path.
// Run in the directory that holds 2D_Real_RVSP_RTM.txt (kernel.cu:542), like the reference.
//   rtm_b200 [run-file] [--gpus N] [--batch B] [--quiet] [--timing]
#include "../../include/rtm_b200.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

int main(int argc, char** argv)
{
    const char* run = "2D_Real_RVSP_RTM.txt";
    int gpus = 0, batch = 0, verbose = 1;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--batch") && i + 1 < argc) batch = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--quiet")) verbose &= ~1;
        else if (!std::strcmp(argv[i], "--timing")) verbose |= 2;
        else run = argv[i];
    }
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = rtm_run_driver(run, gpus, batch, verbose);
    if (rc) {
        std::fprintf(stderr, "rtm_b200: error %d: %s\n", rc, rtm_last_error());
        return 1;
    }
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("%f seconds\n", s);  // kernel.cu:1211-1213
    return 0;
}
