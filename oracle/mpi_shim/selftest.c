/* Self-test of the multi-process mode of oracle/mpi_shim/mpi.h (TEST INFRASTRUCTURE ONLY): every collective the
 * reference uses, with rank-dependent patterns, repeated so that a missing barrier would show.  Prints "shim ok <n>" on
 * rank 0; any mismatch aborts with exit code 1.   gcc -O2 -Ioracle/mpi_shim selftest.c -o selftest */
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "rank %d: check failed at line %d: %s\n", rank, __LINE__, #c); exit(1); } } while (0)

int main(int argc, char **argv) {
    MPI_Init(&argc, &argv);
    int rank, n;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &n);
    for (int rep = 0; rep < 200; rep++) {
        double x = rank == 0 ? 3.5 + rep : -1;
        MPI_Bcast(&x, 1, MPI_DOUBLE, 0, MPI_COMM_WORLD);
        CHECK(x == 3.5 + rep);
        double g[64];
        g[rank] = 10.0 * rank + rep;
        MPI_Allgather(MPI_IN_PLACE, 0, MPI_DOUBLE, g, 1, MPI_DOUBLE, MPI_COMM_WORLD);
        for (int p = 0; p < n; p++) CHECK(g[p] == 10.0 * p + rep);
        int mine = rank * 7 + rep, all[64];
        MPI_Allgather(&mine, 1, MPI_INT, all, 1, MPI_INT, MPI_COMM_WORLD);
        for (int p = 0; p < n; p++) CHECK(all[p] == p * 7 + rep);
        /* allgatherv in place: rank p contributes p + 1 entries */
        int cnt[64], dsp[64], tot = 0, buf[64 * 65];
        for (int p = 0; p < n; p++) { cnt[p] = p + 1; dsp[p] = tot; tot += p + 1; }
        for (int k = 0; k < cnt[rank]; k++) buf[dsp[rank] + k] = 1000 * rank + k + rep;
        MPI_Allgatherv(MPI_IN_PLACE, 0, MPI_DATATYPE_NULL, buf, cnt, dsp, MPI_INT, MPI_COMM_WORLD);
        for (int p = 0; p < n; p++)
            for (int k = 0; k < cnt[p]; k++) CHECK(buf[dsp[p] + k] == 1000 * p + k + rep);
        int gath[64];
        MPI_Gather(&mine, 1, MPI_INT, gath, 1, MPI_INT, 0, MPI_COMM_WORLD);
        if (rank == 0) for (int p = 0; p < n; p++) CHECK(gath[p] == p * 7 + rep);
        unsigned sc[64], got = 0;
        if (rank == 0) for (int p = 0; p < n; p++) sc[p] = 5u * p + rep;
        MPI_Scatter(sc, 1, MPI_UINT32_T, &got, 1, MPI_UINT32_T, 0, MPI_COMM_WORLD);
        CHECK(got == 5u * rank + rep);
        int s2[64], r2[64];
        for (int p = 0; p < n; p++) s2[p] = 100 * rank + p + rep;
        MPI_Alltoall(s2, 1, MPI_INT, r2, 1, MPI_INT, MPI_COMM_WORLD);
        for (int p = 0; p < n; p++) CHECK(r2[p] == 100 * p + rank + rep);
        /* alltoallv as the Adder uses it: fixed displacements p * cap, rank r sends (r + p + rep) % 5 doubles to p */
        enum { cap = 8 };
        double sv[64 * cap], rv[64 * cap];
        int scnt[64], rcnt[64], disp[64];
        for (int p = 0; p < n; p++) {
            scnt[p] = (rank + p + rep) % 5;
            disp[p] = p * cap;
            for (int k = 0; k < scnt[p]; k++) sv[disp[p] + k] = rank * 1e4 + p * 1e2 + k + rep * 1e-3;
        }
        MPI_Alltoall(scnt, 1, MPI_INT, rcnt, 1, MPI_INT, MPI_COMM_WORLD);
        MPI_Alltoallv(sv, scnt, disp, MPI_DOUBLE, rv, rcnt, disp, MPI_DOUBLE, MPI_COMM_WORLD);
        for (int p = 0; p < n; p++) {
            CHECK(rcnt[p] == (p + rank + rep) % 5);
            for (int k = 0; k < rcnt[p]; k++) CHECK(rv[disp[p] + k] == p * 1e4 + rank * 1e2 + k + rep * 1e-3);
        }
    }
    if (rank == 0) printf("shim ok %d\n", n);
    MPI_Finalize();
    return 0;
}
