// mg_main.cpp -- MG_GPU command line, same usage as the reference programs
// (MG_solver_CPU.cpp:41 / MG_solver_GPU.cu:54-97):
//     ./MG_GPU (N_THREADS_OMP) (cycle_filename.txt)
// N_THREADS_OMP is accepted for compatibility (no host loop is threaded here).
// Prints the reference's log and writes Sol_GPU_<cycle file> as CSV (MG_solver_GPU.cu:487-490).
// Environment: MG_DEVICE (default 0), MG_UNFUSED=1 (one ABI operator per reference call),
// MG_NO_CSV=1 (skip the dump, for large N).
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/mg_abi.h"

int main(int argc, char *argv[])
{
    if (argc != 3) {
        printf("[ ERROR ]: Wrong input numbers of parameter.\n");
        return 1;
    }
    printf("OpenMP threads = %d\n", atoi(argv[1]));
    printf("Cycle structure file name = %s\n", argv[2]);

    const char *dev = getenv("MG_DEVICE");
    if (mgInit(dev ? atoi(dev) : 0) != 0) {
        printf("[ ERROR ]: %s\n", mgLastError());
        return 1;
    }
    const char *unfused = getenv("MG_UNFUSED");
    const int flags = (unfused && atoi(unfused)) ? MG_RUN_UNFUSED : MG_RUN_FUSED;

    // size of the top grid: third line of the file
    FILE *fp = fopen(argv[2], "r");
    if (!fp) {
        printf("[ ERROR ]: Cannot open file %s\n", argv[2]);
        return 1;
    }
    double L, mx, my;
    int cs, cn, N_max = 0, N_min = 0;
    const int got = fscanf(fp, "%lf %lf %lf %d %d %d %d", &L, &mx, &my, &cs, &cn, &N_max, &N_min);
    fclose(fp);
    if (got != 7 || N_max < 3) {
        printf("[ ERROR ]: Cannot parse the header of %s\n", argv[2]);
        return 1;
    }

    const bool want_csv = !(getenv("MG_NO_CSV") && atoi(getenv("MG_NO_CSV")));
    std::vector<double> U;
    if (want_csv) U.resize((size_t)N_max * N_max);
    mgCycleResult res;
    const int rc = mgRunCycleFileHost(argv[2], flags, nullptr, want_csv ? U.data() : nullptr, nullptr, 0, &res);
    if (rc != 0) {
        printf("[ ERROR ]: cycle failed (code %d) %s\n", rc, mgLastError());
        return 1;
    }
    if (want_csv) {
        const std::string name = std::string("Sol_GPU_") + argv[2];
        if (mgPrint2File(N_max, U.data(), name.c_str()) != 0) {
            printf("[ ERROR ]: Cannot write %s\n", name.c_str());
            return 1;
        }
        printf("Output file name = %s\n", name.c_str());
    }
    mgShutdown();
    return 0;
}
