/* qxrun -- run a QXTools simulation file triple on a B200 through libqxb200's C ABI.
 *
 * Same flags as the reference's runner (/root/reference/bin/qxrun.jl:12-59):
 *   -d/--dsl FILE  -p/--parameter-file FILE  -i/--input-file FILE  -o/--output-file FILE
 *   -a/--number-amplitudes N  -n/--number-slices N  -t/--timings  -g/--gpu (always on)
 *   -b/--blas-threads N (accepted, ignored: no BLAS on the path)
 *   -m/--mpi: every visible GPU of this node, one compiled replica each, driven by this one process (qxb_multi:
 *   the bitstrings are split across the devices); -s/--sub-comm-size K: groups of K devices share a bitstring range
 *   and split the slices, their partial amplitudes are summed (bin/qxrun.jl:40-46, docs/src/users_guide.md:11-20)
 * Extra: --dtype c32|c64 (default c32), --replan N (re-plan the contraction tree with N candidates),
 *        --device K.
 * Plain C on purpose: it is also the proof that include/qxb200.h is consumable without C++.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qxb200.h"

static int is_opt(const char* a, const char* s, const char* l) { return strcmp(a, s) == 0 || strcmp(a, l) == 0; }

int main(int argc, char** argv) {
    const char *dsl = NULL, *param = NULL, *input = NULL, *output = NULL;
    long long n_amp = -1, n_slices = -1;
    int timings = 0, dtype = QXB_C32, replan = 0, device = 0, multi = 0, sub_comm = 0;
    for (int i = 1; i < argc; ++i) {
        const char* a = argv[i];
        const char* v = i + 1 < argc ? argv[i + 1] : NULL;
        if (is_opt(a, "-d", "--dsl") && v) { dsl = v; ++i; }
        else if (is_opt(a, "-p", "--parameter-file") && v) { param = v; ++i; }
        else if (is_opt(a, "-i", "--input-file") && v) { input = v; ++i; }
        else if (is_opt(a, "-o", "--output-file") && v) { output = v; ++i; }
        else if (is_opt(a, "-a", "--number-amplitudes") && v) { n_amp = atoll(v); ++i; }
        else if (is_opt(a, "-n", "--number-slices") && v) { n_slices = atoll(v); ++i; }
        else if (is_opt(a, "-b", "--blas-threads") && v) { ++i; }
        else if (is_opt(a, "-t", "--timings")) timings = 1;
        else if (is_opt(a, "-g", "--gpu")) {}
        else if (strcmp(a, "--dtype") == 0 && v) { dtype = strcmp(v, "c64") == 0 ? QXB_C64 : QXB_C32; ++i; }
        else if (strcmp(a, "--replan") == 0 && v) { replan = atoi(v); ++i; }
        else if (strcmp(a, "--device") == 0 && v) { device = atoi(v); ++i; }
        else if (is_opt(a, "-m", "--mpi")) multi = 1;
        else if (is_opt(a, "-s", "--sub-comm-size") && v) { sub_comm = atoi(v); ++i; }
        else {
            fprintf(stderr, "usage: qxrun -d FILE.qx -o OUT.jld2 [-p FILE.yml] [-i FILE.jld2] [-a N] [-n N] [-t] [-m [-s K]] "
                            "[--dtype c32|c64] [--replan N] [--device K]\n");
            return 2;
        }
    }
    if (!dsl || !output) { fprintf(stderr, "qxrun: --dsl and --output-file are required\n"); return 2; }
    if (qxb_init(device) != QXB_OK) { fprintf(stderr, "qxrun: %s\n", qxb_last_error()); return 1; }
    /* with -t the reference runs twice so that the second table excludes warm-up (qxrun.jl:89-96) */
    for (int pass = 0; pass < 1 + timings; ++pass) {
        int64_t n = 0;
        double sec[4];
        /* -m: all visible devices (n_devices 0); otherwise the one device selected above */
        if (qxb_execute_files_multi(dsl, input, param, output, dtype, n_amp, n_slices, replan, multi ? 0 : 1, multi ? sub_comm : 1,
                                    &n, sec) != QXB_OK) {
            fprintf(stderr, "qxrun: %s\n", qxb_last_error());
            return 1;
        }
        if (timings && pass == 1) {
            static const char* names[4] = {"Parse input files", "Create Context", "Simulation", "Write results"};
            for (int k = 0; k < 4; ++k) printf("  %-20s %10.3f ms\n", names[k], sec[k] * 1e3);
            printf("  %lld amplitudes, %.4g amplitudes/s\n", (long long)n, sec[2] > 0 ? n / sec[2] : 0.0);
        }
    }
    qxb_shutdown();
    return 0;
}
