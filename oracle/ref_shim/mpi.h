/* Single-rank MPI shim for the serial oracle build of the reference (test infrastructure only).
 * Only the symbols the reference tree uses; semantics of one rank: reductions/alltoall = memcpy. */
#ifndef QB200_ORACLE_MPI_SHIM_H
#define QB200_ORACLE_MPI_SHIM_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int MPI_Comm;
typedef int MPI_Group;
typedef int MPI_Datatype;   /* value = size in bytes */
typedef int MPI_Op;
typedef int MPI_Info;
typedef long long MPI_Offset;
typedef FILE* MPI_File;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL (-1)
#define MPI_COMM_SELF 0
#define MPI_INFO_NULL 0
#define MPI_SUCCESS 0
#define MPI_CHAR 1
#define MPI_BYTE 1
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_FLOAT 4
#define MPI_LONG 8
#define MPI_UNSIGNED_LONG 8
#define MPI_LONG_LONG 8
#define MPI_LONG_LONG_INT 8
#define MPI_DOUBLE 8
#define MPI_DOUBLE_COMPLEX 16
#define MPI_SUM 0
#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_MAX_PROCESSOR_NAME 64
#define MPI_MODE_WRONLY 1
#define MPI_MODE_CREATE 2
#define MPI_MODE_RDONLY 4
#define MPI_SEEK_SET 0
static inline int MPI_Init(int* a, char*** b) { (void)a; (void)b; return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Abort(MPI_Comm c, int e) { (void)c; exit(e); return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int* n) { (void)c; *n = 1; return 0; }
static inline int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return 0; }
static inline int MPI_Comm_group(MPI_Comm c, MPI_Group* g) { (void)c; *g = 0; return 0; }
static inline int MPI_Group_incl(MPI_Group g, int n, const int* r, MPI_Group* o) { (void)g; (void)n; (void)r; *o = 0; return 0; }
static inline int MPI_Group_free(MPI_Group* g) { (void)g; return 0; }
static inline int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm* o) { (void)c; (void)g; *o = 0; return 0; }
static inline int MPI_Comm_dup(MPI_Comm c, MPI_Comm* o) { *o = c; return 0; }
static inline int MPI_Comm_split(MPI_Comm c, int col, int key, MPI_Comm* o) { (void)col; (void)key; *o = c; return 0; }
static inline int MPI_Comm_free(MPI_Comm* c) { (void)c; return 0; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
static inline int MPI_Pcontrol(int l, ...) { (void)l; return 0; }
static inline int MPI_Bcast(void* b, int n, MPI_Datatype t, int r, MPI_Comm c) { (void)b; (void)n; (void)t; (void)r; (void)c; return 0; }
static inline int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c)
{ (void)o; (void)c; if (s != r) memcpy(r, s, (size_t)n * (size_t)t); return 0; }
static inline int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op o, int root, MPI_Comm c)
{ (void)o; (void)c; (void)root; if (s != r) memcpy(r, s, (size_t)n * (size_t)t); return 0; }
static inline int MPI_Scan(const void* s, void* r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c)
{ (void)o; (void)c; if (s != r) memcpy(r, s, (size_t)n * (size_t)t); return 0; }
static inline int MPI_Alltoall(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c)
{ (void)rn; (void)rt; (void)c; memcpy(r, s, (size_t)sn * (size_t)st); return 0; }
static inline int MPI_Alltoallv(const void* s, const int* sc, const int* sd, MPI_Datatype st,
                                void* r, const int* rc, const int* rd, MPI_Datatype rt, MPI_Comm c)
{ (void)rc; (void)rt; (void)c;
  memcpy((char*)r + (size_t)rd[0] * (size_t)rt, (const char*)s + (size_t)sd[0] * (size_t)st, (size_t)sc[0] * (size_t)st); return 0; }
static inline int MPI_Send(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; fprintf(stderr, "mpi shim: MPI_Send on one rank\n"); abort(); return 0; }
static inline int MPI_Recv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status* st)
{ (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)st; fprintf(stderr, "mpi shim: MPI_Recv on one rank\n"); abort(); return 0; }
static inline double MPI_Wtime(void) { struct timeval tv; gettimeofday(&tv, 0); return tv.tv_sec + 1e-6 * tv.tv_usec; }
static inline int MPI_Get_processor_name(char* n, int* l) { strcpy(n, "oracle"); *l = 6; return 0; }
static inline int PMPI_Get_processor_name(char* n, int* l) { return MPI_Get_processor_name(n, l); }
static inline int MPI_File_open(MPI_Comm c, const char* fn, int mode, MPI_Info i, MPI_File* fh)
{ (void)c; (void)i; *fh = fopen(fn, (mode & MPI_MODE_RDONLY) ? "rb" : "wb"); return *fh ? 0 : 1; }
static inline int MPI_File_close(MPI_File* fh) { if (*fh) fclose(*fh); *fh = 0; return 0; }
static inline int MPI_File_write(MPI_File fh, const void* b, int n, MPI_Datatype t, MPI_Status* s)
{ (void)s; fwrite(b, (size_t)t, (size_t)n, fh); return 0; }
static inline int MPI_File_write_at(MPI_File fh, MPI_Offset off, const void* b, int n, MPI_Datatype t, MPI_Status* s)
{ (void)s; fseek(fh, (long)off, SEEK_SET); fwrite(b, (size_t)t, (size_t)n, fh); return 0; }
static inline int MPI_File_write_at_all(MPI_File fh, MPI_Offset off, const void* b, int n, MPI_Datatype t, MPI_Status* s)
{ return MPI_File_write_at(fh, off, b, n, t, s); }
static inline int MPI_File_seek(MPI_File fh, MPI_Offset off, int whence) { (void)whence; fseek(fh, (long)off, SEEK_SET); return 0; }
static inline int MPI_File_get_position(MPI_File fh, MPI_Offset* off) { *off = ftell(fh); return 0; }
static inline int MPI_File_sync(MPI_File fh) { fflush(fh); return 0; }
#ifdef __cplusplus
}
#endif
#endif
