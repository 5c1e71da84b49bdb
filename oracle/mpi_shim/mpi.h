/* Single-rank stand-in for <mpi.h>.
 *
 * TEST INFRASTRUCTURE ONLY.  The reference (sgreene8/FRIES) is an MPI code and this image has no
 * MPI.  This header lets the reference's own sources compile and run as ONE rank so that they can
 * serve as the parity oracle (oracle/_ref) and as the timed CPU baseline.  It implements exactly
 * the MPI symbols the reference's library, tests and the three in-scope drivers use
 * (SURVEY.md section 2c).  Nothing in the product (fries_b200/) includes this file.
 */
#ifndef FRIES_B200_ORACLE_MPI_SHIM_H
#define FRIES_B200_ORACLE_MPI_SHIM_H

#include <stddef.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype; /* value = size of the type in bytes */

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_IN_PLACE ((void *)1)
#define MPI_DATATYPE_NULL 0
#define MPI_DOUBLE 8
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_UINT8_T 1
#define MPI_UINT16_T 2
#define MPI_UINT32_T 4
#define MPI_UINT64_T 8
#define MPI_LONG_LONG 8
#define MPI_UNSIGNED_LONG 8
#define MPI_CHAR 1
#define MPI_BYTE 1

static inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = 0; return MPI_SUCCESS; }
/* The communicator has one rank.  For ONE purpose a test may pretend otherwise: piv_budget (compress_utils.cpp:564-608)
 * does all of its arithmetic on rank 0 and only scatters the result, so with fries_shim_world_size = n rank 0 computes
 * the budgets of n ranks from caller-supplied norms and MPI_Scatter logs what it would have sent. */
__attribute__((weak)) int fries_shim_world_size = 1;
__attribute__((weak)) unsigned char fries_shim_scatter_log[1024];
static inline int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = fries_shim_world_size; return MPI_SUCCESS; }

static inline int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm c) {
    (void)buf; (void)count; (void)t; (void)root; (void)c;
    return MPI_SUCCESS;
}

static inline void fries_shim_copy_(const void *src, void *dst, size_t bytes) {
    if (src != MPI_IN_PLACE && src != dst && bytes) {
        memmove(dst, src, bytes);
    }
}

static inline int MPI_Allgather(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount,
                                MPI_Datatype rt, MPI_Comm c) {
    (void)rcount; (void)rt; (void)c;
    fries_shim_copy_(sbuf, rbuf, (size_t)scount * (size_t)st);
    return MPI_SUCCESS;
}

static inline int MPI_Allgatherv(const void *sbuf, int scount, MPI_Datatype st, void *rbuf,
                                 const int *rcounts, const int *displs, MPI_Datatype rt, MPI_Comm c) {
    (void)rcounts; (void)c;
    if (sbuf != MPI_IN_PLACE) {
        fries_shim_copy_(sbuf, (char *)rbuf + (size_t)displs[0] * (size_t)rt, (size_t)scount * (size_t)st);
    }
    return MPI_SUCCESS;
}

static inline int MPI_Gather(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount,
                             MPI_Datatype rt, int root, MPI_Comm c) {
    (void)rcount; (void)rt; (void)root; (void)c;
    fries_shim_copy_(sbuf, rbuf, (size_t)scount * (size_t)st);
    return MPI_SUCCESS;
}

static inline int MPI_Scatter(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount,
                              MPI_Datatype rt, int root, MPI_Comm c) {
    (void)root; (void)c;
    if (fries_shim_world_size > 1 && (size_t)fries_shim_world_size * (size_t)scount * (size_t)st <= sizeof(fries_shim_scatter_log)) {
        memcpy(fries_shim_scatter_log, sbuf, (size_t)fries_shim_world_size * (size_t)scount * (size_t)st);
    }
    if (rbuf != MPI_IN_PLACE) {
        fries_shim_copy_(sbuf, rbuf, (size_t)rcount * (size_t)rt);
    }
    return MPI_SUCCESS;
}

static inline int MPI_Alltoall(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount,
                               MPI_Datatype rt, MPI_Comm c) {
    (void)rcount; (void)rt; (void)c;
    fries_shim_copy_(sbuf, rbuf, (size_t)scount * (size_t)st);
    return MPI_SUCCESS;
}

static inline int MPI_Alltoallv(const void *sbuf, const int *scounts, const int *sdispls, MPI_Datatype st,
                                void *rbuf, const int *rcounts, const int *rdispls, MPI_Datatype rt,
                                MPI_Comm c) {
    (void)rcounts; (void)c;
    if (sbuf != MPI_IN_PLACE) {
        memmove((char *)rbuf + (size_t)rdispls[0] * (size_t)rt,
                (const char *)sbuf + (size_t)sdispls[0] * (size_t)st, (size_t)scounts[0] * (size_t)st);
    }
    return MPI_SUCCESS;
}

#ifdef __cplusplus
}
#endif

#endif
