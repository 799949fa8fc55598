// monortm_main.cpp -- `monortm_b200`: the stand-in for the reference executable (PROGRAM MONORTM,
// src/monortm.f90) for layer input.  Run it in a directory holding MONORTM.IN, MONORTM_PROF.IN and
// TAPE3 like the reference (run/run_monortm_examples:18-123); it writes MONORTM.OUT / MONORTM.LOG.
//   monortm_b200 [-C workdir] [-d device] [--nwnmx N] [-q]
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../../include/monortm_b200.h"

int main(int argc, char** argv)
{
    const char* dir = "";
    int device = 0, verbose = 1;
    long long nwnmx = 0;
    for (int i = 1; i < argc; i++) {
        if (!std::strcmp(argv[i], "-C") && i + 1 < argc) dir = argv[++i];
        else if (!std::strcmp(argv[i], "-d") && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--nwnmx") && i + 1 < argc) nwnmx = std::atoll(argv[++i]);
        else if (!std::strcmp(argv[i], "-q")) verbose = 0;
        else {
            std::fprintf(stderr, "usage: %s [-C workdir] [-d device] [--nwnmx N] [-q]\n", argv[0]);
            return 2;
        }
    }
    if (verbose) std::printf(" **********************************\n *        M O N O R T M           *\n *  %s\n **********************************\n", mrtm_version());
    const int rc = mrtm_host_run_monortm(dir, device, nwnmx, verbose);
    if (rc) {
        std::fprintf(stderr, "STOP %s (%s)\n", mrtm_host_last_error(), mrtm_strerror(rc));
        return 1;
    }
    return 0;
}
