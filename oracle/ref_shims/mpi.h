// TEST INFRASTRUCTURE ONLY — single-rank MPI stand-in used to compile the unmodified
// reference (/root/reference/src/*.cpp) into oracle/_ref/. This image has no MPI.
// Only the calls the reference makes are provided; files use POSIX pread/pwrite so
// that sparse writes leave the same NUL holes MPI-IO leaves (SURVEY.md §8 a-io).
#pragma once
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstddef>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef long long MPI_Offset;
struct MPI_Status { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; int count_bytes; };
struct vo_mpi_file { int fd; MPI_Offset disp; };
typedef vo_mpi_file* MPI_File;

enum { MPI_SUCCESS = 0, MPI_ERR_OTHER = 15 };
enum { MPI_COMM_WORLD = 1, MPI_COMM_SELF = 2 };
enum { MPI_INFO_NULL = 0 };
enum { MPI_THREAD_SINGLE = 0, MPI_THREAD_FUNNELED, MPI_THREAD_SERIALIZED, MPI_THREAD_MULTIPLE };
enum { MPI_CHAR = 1, MPI_INT = 4, MPI_DOUBLE = 8, MPI_UNSIGNED_LONG_LONG = 9 };
enum { MPI_SUM = 1, MPI_MAX = 2 };
enum { MPI_MODE_RDONLY = 1, MPI_MODE_WRONLY = 2, MPI_MODE_RDWR = 4, MPI_MODE_CREATE = 8, MPI_MODE_EXCL = 16 };

static inline int vo_mpi_dtsize(MPI_Datatype dt) {
    switch (dt) { case MPI_CHAR: return 1; case MPI_INT: return 4; case MPI_DOUBLE: return 8;
                  case MPI_UNSIGNED_LONG_LONG: return 8; default: return 0; }
}
static inline int MPI_Init_thread(int*, char***, int required, int* provided) { *provided = required; return MPI_SUCCESS; }
static inline int MPI_Finalize() { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm, int* s) { *s = 1; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm, int code) { fflush(stdout); fflush(stderr); _exit(code ? code : 1); return 0; }
static inline double MPI_Wtime() { struct timeval tv; gettimeofday(&tv, nullptr); return tv.tv_sec + 1e-6 * tv.tv_usec; }
static inline int MPI_Type_size(MPI_Datatype dt, int* sz) { *sz = vo_mpi_dtsize(dt); return MPI_SUCCESS; }
static inline int MPI_Allreduce(const void* s, void* r, int count, MPI_Datatype dt, MPI_Op, MPI_Comm) {
    memcpy(r, s, (size_t)count * vo_mpi_dtsize(dt)); return MPI_SUCCESS;
}
static inline int MPI_Get_count(const MPI_Status* st, MPI_Datatype dt, int* count) {
    *count = st->count_bytes / vo_mpi_dtsize(dt); return MPI_SUCCESS;
}
static inline int MPI_File_open(MPI_Comm, const char* path, int amode, MPI_Info, MPI_File* fh) {
    int flags = 0;
    if (amode & MPI_MODE_RDWR) flags |= O_RDWR; else if (amode & MPI_MODE_WRONLY) flags |= O_WRONLY; else flags |= O_RDONLY;
    if (amode & MPI_MODE_CREATE) flags |= O_CREAT;
    if (amode & MPI_MODE_EXCL) flags |= O_EXCL;
    int fd = open(path, flags, 0644);
    if (fd < 0) { *fh = nullptr; return MPI_ERR_OTHER; }
    *fh = new vo_mpi_file{fd, 0};
    return MPI_SUCCESS;
}
static inline int MPI_File_close(MPI_File* fh) {
    if (fh && *fh) { close((*fh)->fd); delete *fh; *fh = nullptr; }
    return MPI_SUCCESS;
}
static inline int MPI_File_delete(const char* path, MPI_Info) { return unlink(path) == 0 ? MPI_SUCCESS : MPI_ERR_OTHER; }
static inline int MPI_File_set_view(MPI_File fh, MPI_Offset disp, MPI_Datatype, MPI_Datatype, const char*, MPI_Info) {
    if (!fh) return MPI_ERR_OTHER;   // the reference does not check opens in mpi_store/read_vec (utilities.cpp:245,256)
    fh->disp = disp; return MPI_SUCCESS;
}
static inline int vo_mpi_rw(MPI_File fh, MPI_Offset off, void* buf, int count, MPI_Datatype dt, MPI_Status* st, bool wr) {
    if (!fh) return MPI_ERR_OTHER;
    size_t es = vo_mpi_dtsize(dt), total = (size_t)count * es, done = 0;
    MPI_Offset pos = fh->disp + off * (MPI_Offset)1;   // etype is bytes unless a view was set; the reference's views use
                                                       // MPI_DOUBLE etype with offset 0 only, so byte maths is exact here
    while (done < total) {
        ssize_t k = wr ? pwrite(fh->fd, (const char*)buf + done, total - done, pos + done)
                       : pread(fh->fd, (char*)buf + done, total - done, pos + done);
        if (k < 0) return MPI_ERR_OTHER;
        if (k == 0) break;
        done += (size_t)k;
    }
    if (st) { st->MPI_SOURCE = 0; st->MPI_TAG = 0; st->MPI_ERROR = 0; st->count_bytes = (int)done; }
    return MPI_SUCCESS;
}
static inline int MPI_File_read_at(MPI_File fh, MPI_Offset off, void* buf, int count, MPI_Datatype dt, MPI_Status* st) {
    return vo_mpi_rw(fh, off, buf, count, dt, st, false);
}
static inline int MPI_File_write_at(MPI_File fh, MPI_Offset off, const void* buf, int count, MPI_Datatype dt, MPI_Status* st) {
    return vo_mpi_rw(fh, off, const_cast<void*>(buf), count, dt, st, true);
}
static inline int MPI_File_write_at_all(MPI_File fh, MPI_Offset off, const void* buf, int count, MPI_Datatype dt, MPI_Status* st) {
    return vo_mpi_rw(fh, off, const_cast<void*>(buf), count, dt, st, true);
}
