/* Single-rank MPI stand-in used ONLY to compile the unmodified reference CPU sources
 * into oracle/_ref (test infrastructure, never linked into the product library).
 *
 * With one rank no point-to-point call is ever reached by the reference; collectives
 * degenerate to a memcpy from the send to the receive buffer. Datatype handles encode
 * the element size in bytes so the collectives know how much to copy.
 */
#ifndef SPHX_ORACLE_MPI_SHIM_H
#define SPHX_ORACLE_MPI_SHIM_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Fint;

typedef struct MPI_Status
{
    int MPI_SOURCE;
    int MPI_TAG;
    int MPI_ERROR;
    int count_bytes;
} MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_IN_PLACE ((void*)1)
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)

/* datatype handle == sizeof(element) */
#define MPI_CHAR 1
#define MPI_UNSIGNED_CHAR 1
#define MPI_BYTE 1
#define MPI_SHORT 2
#define MPI_UNSIGNED_SHORT 2
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_FLOAT 4
#define MPI_LONG 8
#define MPI_UNSIGNED_LONG 8
#define MPI_UNSIGNED_LONG_LONG 8
#define MPI_LONG_LONG 8
#define MPI_DOUBLE 8

#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3

static inline int MPI_Init(int* a, char*** b)
{
    (void)a;
    (void)b;
    return 0;
}
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_rank(MPI_Comm c, int* r)
{
    (void)c;
    *r = 0;
    return 0;
}
static inline int MPI_Comm_size(MPI_Comm c, int* s)
{
    (void)c;
    *s = 1;
    return 0;
}
static inline int MPI_Get_version(int* v, int* sv)
{
    *v  = 3;
    *sv = 1;
    return 0;
}
static inline int MPI_Barrier(MPI_Comm c)
{
    (void)c;
    return 0;
}
static inline int MPI_Abort(MPI_Comm c, int code)
{
    (void)c;
    fprintf(stderr, "MPI_Abort(%d) in single-rank shim\n", code);
    exit(code);
}
static inline MPI_Fint MPI_Comm_c2f(MPI_Comm c) { return c; }

static inline int shim_copy_(const void* s, void* r, int count, MPI_Datatype t)
{
    if (s != MPI_IN_PLACE && s != r) { memcpy(r, s, (size_t)count * (size_t)t); }
    return 0;
}
static inline int MPI_Allreduce(const void* s, void* r, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
    (void)op;
    (void)c;
    return shim_copy_(s, r, count, t);
}
static inline int MPI_Reduce(const void* s, void* r, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{
    (void)op;
    (void)c;
    (void)root;
    return shim_copy_(s, r, count, t);
}
static inline int MPI_Allgather(const void* s, int sc, MPI_Datatype st, void* r, int rc, MPI_Datatype rt, MPI_Comm c)
{
    (void)rc;
    (void)rt;
    (void)c;
    return shim_copy_(s, r, sc, st);
}
static inline int MPI_Allgatherv(const void* s, int sc, MPI_Datatype st, void* r, const int* rc, const int* displs,
                                 MPI_Datatype rt, MPI_Comm c)
{
    (void)rc;
    (void)c;
    if (s != MPI_IN_PLACE) { memcpy((char*)r + (size_t)displs[0] * (size_t)rt, s, (size_t)sc * (size_t)st); }
    return 0;
}

static inline int shim_p2p_unreachable_(const char* what)
{
    fprintf(stderr, "%s reached in single-rank MPI shim\n", what);
    abort();
    return 1;
}
static inline int MPI_Isend(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request* rq)
{
    (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; (void)rq;
    return shim_p2p_unreachable_("MPI_Isend");
}
static inline int MPI_Irecv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* rq)
{
    (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)rq;
    return shim_p2p_unreachable_("MPI_Irecv");
}
static inline int MPI_Send(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c)
{
    (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c;
    return shim_p2p_unreachable_("MPI_Send");
}
static inline int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* s)
{
    (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)s;
    return shim_p2p_unreachable_("MPI_Recv");
}
static inline int MPI_Probe(int src, int tag, MPI_Comm c, MPI_Status* s)
{
    (void)src; (void)tag; (void)c; (void)s;
    return shim_p2p_unreachable_("MPI_Probe");
}
static inline int MPI_Get_count(const MPI_Status* s, MPI_Datatype t, int* count)
{
    *count = s ? s->count_bytes / t : 0;
    return 0;
}
static inline int MPI_Waitall(int n, MPI_Request* rq, MPI_Status* st)
{
    (void)n; (void)rq; (void)st;
    return 0;
}
static inline int MPI_Wait(MPI_Request* rq, MPI_Status* st)
{
    (void)rq; (void)st;
    return 0;
}

#endif
