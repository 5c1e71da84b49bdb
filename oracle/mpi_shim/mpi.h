/* Stand-in for <mpi.h>: one rank, or N processes over a shared-memory file (see "multi-process mode" below).
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
#include <fcntl.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <unistd.h>

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

/* ---- optional multi-process mode ---------------------------------------------------------------------------
 * oracle/mpi_shim/shimrun.py -n N <driver> ... starts N copies of a reference driver with FRIES_SHIM_SHM (a file that
 * all of them map), FRIES_SHIM_RANK, FRIES_SHIM_SIZE and FRIES_SHIM_SLOT (bytes per rank) in the environment.  Every
 * collective the reference uses (all blocking, all on MPI_COMM_WORLD) is then "publish my part in my slot, barrier,
 * copy what I need from the others' slots, barrier".  Without FRIES_SHIM_SHM the communicator has one rank, as before.
 * This is how the CPU baseline uses the GPU box's host cores without an MPI installation. */

typedef struct {
    int arrive;
    int sense;
    int n_ranks;
    int pad;
    size_t slot_bytes;
} fries_shim_shared;
#define FRIES_SHIM_HDR 4096

/* Number of ranks.  Without the multi-process mode it is 1; for ONE purpose a test may pretend otherwise: piv_budget
 * (compress_utils.cpp:564-608) does all of its arithmetic on rank 0 and only scatters the result, so with
 * fries_shim_world_size = n (and no shared segment) rank 0 computes the budgets of n ranks from caller-supplied norms and
 * MPI_Scatter logs what it would have sent. */
__attribute__((weak)) int fries_shim_world_size = 1;
__attribute__((weak)) unsigned char fries_shim_scatter_log[1024];
static inline int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = fries_shim_world_size; return MPI_SUCCESS; }

__attribute__((weak)) fries_shim_shared *fries_shim_sh = 0;
__attribute__((weak)) int fries_shim_rank_ = 0;
__attribute__((weak)) int fries_shim_sense_ = 0;

static inline int MPI_Init(int *argc, char ***argv) {
    (void)argc; (void)argv;
    const char *path = getenv("FRIES_SHIM_SHM");
    if (path && !fries_shim_sh) {
        int size = atoi(getenv("FRIES_SHIM_SIZE")), rank = atoi(getenv("FRIES_SHIM_RANK"));
        size_t slot = (size_t)atoll(getenv("FRIES_SHIM_SLOT"));
        int fd = open(path, O_RDWR);
        if (fd < 0) { perror("mpi shim: open"); exit(3); }
        void *m = mmap(0, FRIES_SHIM_HDR + (size_t)size * slot, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        if (m == MAP_FAILED) { perror("mpi shim: mmap"); exit(3); }
        close(fd);
        fries_shim_sh = (fries_shim_shared *)m;   /* the launcher wrote n_ranks and slot_bytes, counters are zero */
        fries_shim_rank_ = rank;
        fries_shim_world_size = size;
    }
    return MPI_SUCCESS;
}
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = fries_shim_rank_; return MPI_SUCCESS; }

static inline void fries_shim_barrier_(void) {
    fries_shim_shared *s = fries_shim_sh;
    int my = fries_shim_sense_ ^= 1;
    if (__atomic_add_fetch(&s->arrive, 1, __ATOMIC_ACQ_REL) == s->n_ranks) {
        __atomic_store_n(&s->arrive, 0, __ATOMIC_RELAXED);
        __atomic_store_n(&s->sense, my, __ATOMIC_RELEASE);
    } else {
        unsigned spins = 0;  /* spin, then yield, then sleep: more ranks than cores must not starve the last arrival */
        while (__atomic_load_n(&s->sense, __ATOMIC_ACQUIRE) != my) {
            if (++spins > 20000) usleep(50);
            else if (spins > 2000) sched_yield();
        }
    }
}
static inline char *fries_shim_slot_(int p) { return (char *)fries_shim_sh + FRIES_SHIM_HDR + (size_t)p * fries_shim_sh->slot_bytes; }
static inline void fries_shim_need_(size_t bytes) {
    if (bytes > fries_shim_sh->slot_bytes) {
        fprintf(stderr, "mpi shim: a collective needs %zu bytes per rank, slots hold %zu (raise FRIES_SHIM_SLOT)\n", bytes,
                fries_shim_sh->slot_bytes);
        exit(4);
    }
}

static inline int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm c) {
    (void)c;
    if (!fries_shim_sh) return MPI_SUCCESS;
    size_t n = (size_t)count * (size_t)t;
    fries_shim_need_(n);
    if (fries_shim_rank_ == root) memcpy(fries_shim_slot_(root), buf, n);
    fries_shim_barrier_();
    if (fries_shim_rank_ != root) memcpy(buf, fries_shim_slot_(root), n);
    fries_shim_barrier_();
    return MPI_SUCCESS;
}

static inline void fries_shim_copy_(const void *src, void *dst, size_t bytes) {
    if (src != MPI_IN_PLACE && src != dst && bytes) {
        memmove(dst, src, bytes);
    }
}

static inline int MPI_Allgather(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount,
                                MPI_Datatype rt, MPI_Comm c) {
    (void)c;
    if (!fries_shim_sh) {
        fries_shim_copy_(sbuf, rbuf, (size_t)scount * (size_t)st);
        return MPI_SUCCESS;
    }
    size_t blk = (size_t)rcount * (size_t)rt;
    int me = fries_shim_rank_, n = fries_shim_sh->n_ranks;
    fries_shim_need_(blk);
    memcpy(fries_shim_slot_(me), sbuf == MPI_IN_PLACE ? (const char *)rbuf + (size_t)me * blk : (const char *)sbuf, blk);
    fries_shim_barrier_();
    for (int p = 0; p < n; p++) memcpy((char *)rbuf + (size_t)p * blk, fries_shim_slot_(p), blk);
    fries_shim_barrier_();
    return MPI_SUCCESS;
}

static inline int MPI_Allgatherv(const void *sbuf, int scount, MPI_Datatype st, void *rbuf,
                                 const int *rcounts, const int *displs, MPI_Datatype rt, MPI_Comm c) {
    (void)c;
    if (!fries_shim_sh) {
        if (sbuf != MPI_IN_PLACE) {
            fries_shim_copy_(sbuf, (char *)rbuf + (size_t)displs[0] * (size_t)rt, (size_t)scount * (size_t)st);
        }
        return MPI_SUCCESS;
    }
    int me = fries_shim_rank_, n = fries_shim_sh->n_ranks;
    size_t mine = (size_t)rcounts[me] * (size_t)rt;
    fries_shim_need_(mine);
    memcpy(fries_shim_slot_(me), sbuf == MPI_IN_PLACE ? (const char *)rbuf + (size_t)displs[me] * (size_t)rt : (const char *)sbuf, mine);
    fries_shim_barrier_();
    for (int p = 0; p < n; p++) memcpy((char *)rbuf + (size_t)displs[p] * (size_t)rt, fries_shim_slot_(p), (size_t)rcounts[p] * (size_t)rt);
    fries_shim_barrier_();
    return MPI_SUCCESS;
}

static inline int MPI_Gather(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount,
                             MPI_Datatype rt, int root, MPI_Comm c) {
    (void)c;
    if (!fries_shim_sh) {
        fries_shim_copy_(sbuf, rbuf, (size_t)scount * (size_t)st);
        return MPI_SUCCESS;
    }
    int me = fries_shim_rank_, n = fries_shim_sh->n_ranks;
    size_t blk = (size_t)scount * (size_t)st;
    if (sbuf == MPI_IN_PLACE) blk = (size_t)rcount * (size_t)rt;
    fries_shim_need_(blk);
    memcpy(fries_shim_slot_(me), sbuf == MPI_IN_PLACE ? (const char *)rbuf + (size_t)me * blk : (const char *)sbuf, blk);
    fries_shim_barrier_();
    if (me == root)
        for (int p = 0; p < n; p++) memcpy((char *)rbuf + (size_t)p * blk, fries_shim_slot_(p), blk);
    fries_shim_barrier_();
    return MPI_SUCCESS;
}

static inline int MPI_Scatter(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount,
                              MPI_Datatype rt, int root, MPI_Comm c) {
    (void)c;
    if (!fries_shim_sh) {
        if (fries_shim_world_size > 1 && (size_t)fries_shim_world_size * (size_t)scount * (size_t)st <= sizeof(fries_shim_scatter_log)) {
            memcpy(fries_shim_scatter_log, sbuf, (size_t)fries_shim_world_size * (size_t)scount * (size_t)st);
        }
        if (rbuf != MPI_IN_PLACE) {
            fries_shim_copy_(sbuf, rbuf, (size_t)rcount * (size_t)rt);
        }
        return MPI_SUCCESS;
    }
    int me = fries_shim_rank_, n = fries_shim_sh->n_ranks;
    size_t blk = (size_t)rcount * (size_t)rt;
    fries_shim_need_((size_t)n * blk);
    if (me == root) memcpy(fries_shim_slot_(root), sbuf, (size_t)n * (size_t)scount * (size_t)st);
    fries_shim_barrier_();
    if (rbuf != MPI_IN_PLACE) memcpy(rbuf, fries_shim_slot_(root) + (size_t)me * blk, blk);
    fries_shim_barrier_();
    return MPI_SUCCESS;
}

static inline int MPI_Alltoall(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount,
                               MPI_Datatype rt, MPI_Comm c) {
    (void)c;
    if (!fries_shim_sh) {
        fries_shim_copy_(sbuf, rbuf, (size_t)scount * (size_t)st);
        return MPI_SUCCESS;
    }
    int me = fries_shim_rank_, n = fries_shim_sh->n_ranks;
    size_t blk = (size_t)rcount * (size_t)rt;
    fries_shim_need_((size_t)n * blk);
    memcpy(fries_shim_slot_(me), sbuf == MPI_IN_PLACE ? rbuf : sbuf, (size_t)n * blk);
    fries_shim_barrier_();
    for (int p = 0; p < n; p++) memcpy((char *)rbuf + (size_t)p * blk, fries_shim_slot_(p) + (size_t)me * blk, blk);
    fries_shim_barrier_();
    return MPI_SUCCESS;
}

static inline int MPI_Alltoallv(const void *sbuf, const int *scounts, const int *sdispls, MPI_Datatype st,
                                void *rbuf, const int *rcounts, const int *rdispls, MPI_Datatype rt,
                                MPI_Comm c) {
    (void)c;
    if (!fries_shim_sh) {
        (void)rcounts;
        if (sbuf != MPI_IN_PLACE) {
            memmove((char *)rbuf + (size_t)rdispls[0] * (size_t)rt,
                    (const char *)sbuf + (size_t)sdispls[0] * (size_t)st, (size_t)scounts[0] * (size_t)st);
        }
        return MPI_SUCCESS;
    }
    /* slot: int counts[n], int displs[n], then (64-byte aligned) the send buffer's used segments at their own offsets */
    int me = fries_shim_rank_, n = fries_shim_sh->n_ranks;
    size_t hdr = (((size_t)2 * n * sizeof(int)) + 63) & ~(size_t)63, extent = 0;
    for (int p = 0; p < n; p++) {
        size_t e = ((size_t)sdispls[p] + (size_t)scounts[p]) * (size_t)st;
        if (scounts[p] && e > extent) extent = e;
    }
    fries_shim_need_(hdr + extent);
    char *mine = fries_shim_slot_(me);
    memcpy(mine, scounts, (size_t)n * sizeof(int));
    memcpy(mine + (size_t)n * sizeof(int), sdispls, (size_t)n * sizeof(int));
    for (int p = 0; p < n; p++)
        memcpy(mine + hdr + (size_t)sdispls[p] * (size_t)st, (const char *)sbuf + (size_t)sdispls[p] * (size_t)st, (size_t)scounts[p] * (size_t)st);
    fries_shim_barrier_();
    for (int p = 0; p < n; p++) {
        const char *theirs = fries_shim_slot_(p);
        int cnt = ((const int *)theirs)[me], dsp = ((const int *)theirs)[n + me];
        if (cnt != rcounts[p]) {
            fprintf(stderr, "mpi shim: Alltoallv count mismatch (rank %d expects %d from %d, which sends %d)\n", me, rcounts[p], p, cnt);
            exit(5);
        }
        memcpy((char *)rbuf + (size_t)rdispls[p] * (size_t)rt, theirs + hdr + (size_t)dsp * (size_t)st, (size_t)cnt * (size_t)st);
    }
    fries_shim_barrier_();
    return MPI_SUCCESS;
}

#ifdef __cplusplus
}
#endif

#endif
