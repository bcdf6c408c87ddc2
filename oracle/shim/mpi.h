/* oracle/shim/mpi.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A header-only stand-in for <mpi.h> so that the reference's own CPU code
 * (/root/reference/include/metaLBM, compiled where it lies by oracle/refbuild.py)
 * builds in an image that has no MPI.  It implements exactly the calls the
 * reference makes (MPIInitializer.h:29-62, Communication.h:61-88,145-178,
 * CUDAInitializer.h:16-29) and nothing else.
 *
 * Ranks are forked processes: MPI_Init_thread() forks NPROCS-1 children that
 * share one anonymous MAP_SHARED segment holding per-rank mailboxes, a barrier
 * and a reduction scratch area.  NPROCS is the same -DNPROCS=... macro the
 * reference reads (src/Input_*.in: "constexpr int numProcs = NPROCS").
 *
 * Semantics relied upon by the reference's halo exchange: for every population
 * a rank posts Irecv(tag), Isend(tag), Waitall -- one message per (peer,tag) in
 * flight, so a mailbox per (receiver, tag-slot) with a full/empty flag suffices.
 */
#pragma once

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <sched.h>
#include <unistd.h>

#ifndef NPROCS
#define NPROCS 1
#endif

typedef int MPI_Comm;
typedef int MPI_Info;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int source; int tag; } MPI_Status;
typedef struct {
  int kind; /* 0 = none, 1 = recv, 2 = send (already delivered) */
  void* buffer;
  size_t bytes;
  int peer;
  int tag;
} MPI_Request;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_DOUBLE 8
#define MPI_SUM 1
#define MPI_IN_PLACE ((void*)-1)
#define MPI_THREAD_FUNNELED 1
#define MPI_MAX_PROCESSOR_NAME 64
#define MPI_COMM_TYPE_SHARED 1
#define MPI_INFO_NULL 0

#define MLBM_SHIM_NSLOT 4
#define MLBM_SHIM_MAXRED 4096

typedef struct {
  volatile int full;
  size_t bytes;
} mlbm_shim_slot_header;

typedef struct {
  volatile int barrierCount;
  volatile int barrierSense;
  double reduction[NPROCS][MLBM_SHIM_MAXRED];
} mlbm_shim_control;

static struct {
  int rank;
  int size;
  size_t slotBytes;   /* payload capacity of one mailbox */
  size_t slotStride;  /* header + payload, 64-byte aligned */
  char* mailboxes;    /* [size][NSLOT] */
  mlbm_shim_control* control;
  int localSense;
  pid_t children[NPROCS > 1 ? NPROCS : 1];
} mlbm_shim = {0, 1, 0, 0, NULL, NULL, 0, {0}};

static inline int mlbm_shim_slot_of_tag(int tag) {
  /* the reference uses tag 17 (to the right) and tag 23 (to the left) */
  return tag == 17 ? 0 : (tag == 23 ? 1 : 2 + (tag & 1));
}

static inline mlbm_shim_slot_header* mlbm_shim_mailbox(int rank, int slot) {
  return (mlbm_shim_slot_header*)(mlbm_shim.mailboxes +
                                  ((size_t)rank * MLBM_SHIM_NSLOT + slot) *
                                      mlbm_shim.slotStride);
}

static inline void mlbm_shim_relax(void) { sched_yield(); }

static inline int MPI_Barrier(MPI_Comm comm) {
  (void)comm;
  if (mlbm_shim.size == 1) return MPI_SUCCESS;
  mlbm_shim_control* c = mlbm_shim.control;
  int sense = !mlbm_shim.localSense;
  mlbm_shim.localSense = sense;
  if (__atomic_add_fetch(&c->barrierCount, 1, __ATOMIC_ACQ_REL) == mlbm_shim.size) {
    __atomic_store_n(&c->barrierCount, 0, __ATOMIC_RELAXED);
    __atomic_store_n(&c->barrierSense, sense, __ATOMIC_RELEASE);
  } else {
    while (__atomic_load_n(&c->barrierSense, __ATOMIC_ACQUIRE) != sense) mlbm_shim_relax();
  }
  return MPI_SUCCESS;
}

static inline int MPI_Init_thread(int* argc, char*** argv, int required, int* provided) {
  (void)argc; (void)argv;
  if (provided) *provided = required;
  mlbm_shim.size = NPROCS;
  mlbm_shim.rank = 0;

  size_t slotBytes = 0;
  const char* env = getenv("MLBM_SHIM_SLOT_BYTES");
  if (env) slotBytes = (size_t)strtoull(env, NULL, 10);
#if defined(GLOBAL_LENGTH_Y) && defined(GLOBAL_LENGTH_Z)
  if (!slotBytes) slotBytes = (size_t)3 * (GLOBAL_LENGTH_Y + 6) * (GLOBAL_LENGTH_Z + 6) * sizeof(double);
#endif
  if (!slotBytes) slotBytes = (size_t)1 << 24;
  mlbm_shim.slotBytes = slotBytes;
  mlbm_shim.slotStride = (sizeof(mlbm_shim_slot_header) + slotBytes + 63) & ~(size_t)63;

  size_t mailBytes = (size_t)NPROCS * MLBM_SHIM_NSLOT * mlbm_shim.slotStride;
  size_t total = mailBytes + sizeof(mlbm_shim_control);
  void* segment = mmap(NULL, total, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (segment == MAP_FAILED) { perror("mlbm mpi shim: mmap"); exit(1); }
  memset(segment, 0, total);
  mlbm_shim.mailboxes = (char*)segment;
  mlbm_shim.control = (mlbm_shim_control*)((char*)segment + mailBytes);

  fflush(NULL);
  for (int r = 1; r < NPROCS; ++r) {
    pid_t pid = fork();
    if (pid < 0) { perror("mlbm mpi shim: fork"); exit(1); }
    if (pid == 0) { mlbm_shim.rank = r; break; }
    mlbm_shim.children[r] = pid;
  }
  return MPI_SUCCESS;
}

static inline int MPI_Finalize(void) {
  MPI_Barrier(MPI_COMM_WORLD);
  fflush(NULL);
  if (mlbm_shim.rank == 0) {
    for (int r = 1; r < mlbm_shim.size; ++r) {
      int status = 0;
      waitpid(mlbm_shim.children[r], &status, 0);
    }
  }
  return MPI_SUCCESS;
}

static inline int MPI_Get_processor_name(char* name, int* length) {
  strcpy(name, "shim");
  *length = 4;
  return MPI_SUCCESS;
}
static inline int MPI_Comm_size(MPI_Comm c, int* size) { (void)c; *size = mlbm_shim.size; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int* rank) { (void)c; *rank = mlbm_shim.rank; return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; fflush(NULL); _exit(code ? code : 1); return 0; }
static inline int MPI_Info_create(MPI_Info* info) { *info = 0; return MPI_SUCCESS; }
static inline int MPI_Info_free(MPI_Info* info) { (void)info; return MPI_SUCCESS; }
static inline int MPI_Comm_split_type(MPI_Comm c, int type, int key, MPI_Info info, MPI_Comm* out) {
  (void)c; (void)type; (void)key; (void)info; *out = 0; return MPI_SUCCESS;
}
static inline int MPI_Comm_free(MPI_Comm* c) { (void)c; return MPI_SUCCESS; }

static inline int MPI_Irecv(void* buffer, int count, MPI_Datatype type, int source, int tag,
                            MPI_Comm comm, MPI_Request* request) {
  (void)comm;
  request->kind = 1;
  request->buffer = buffer;
  request->bytes = (size_t)count * (size_t)type;
  request->peer = source;
  request->tag = tag;
  return MPI_SUCCESS;
}

static inline int MPI_Isend(const void* buffer, int count, MPI_Datatype type, int destination,
                            int tag, MPI_Comm comm, MPI_Request* request) {
  (void)comm;
  size_t bytes = (size_t)count * (size_t)type;
  if (bytes > mlbm_shim.slotBytes) {
    fprintf(stderr, "mlbm mpi shim: message of %zu bytes exceeds mailbox (%zu); set MLBM_SHIM_SLOT_BYTES\n",
            bytes, mlbm_shim.slotBytes);
    MPI_Abort(comm, 2);
  }
  mlbm_shim_slot_header* box = mlbm_shim_mailbox(destination, mlbm_shim_slot_of_tag(tag));
  while (__atomic_load_n(&box->full, __ATOMIC_ACQUIRE)) mlbm_shim_relax();
  box->bytes = bytes;
  memcpy((char*)(box + 1), buffer, bytes);
  __atomic_store_n(&box->full, 1, __ATOMIC_RELEASE);
  request->kind = 2;
  request->buffer = NULL;
  request->bytes = bytes;
  request->peer = destination;
  request->tag = tag;
  return MPI_SUCCESS;
}

static inline int MPI_Waitall(int count, MPI_Request* requests, MPI_Status* statuses) {
  (void)statuses;
  for (int i = 0; i < count; ++i) {
    MPI_Request* request = &requests[i];
    if (request->kind != 1) { request->kind = 0; continue; }
    mlbm_shim_slot_header* box = mlbm_shim_mailbox(mlbm_shim.rank, mlbm_shim_slot_of_tag(request->tag));
    while (!__atomic_load_n(&box->full, __ATOMIC_ACQUIRE)) mlbm_shim_relax();
    if (box->bytes != request->bytes) {
      fprintf(stderr, "mlbm mpi shim: size mismatch (%zu sent, %zu expected)\n", box->bytes, request->bytes);
      MPI_Abort(0, 3);
    }
    memcpy(request->buffer, (char*)(box + 1), request->bytes);
    __atomic_store_n(&box->full, 0, __ATOMIC_RELEASE);
    request->kind = 0;
  }
  return MPI_SUCCESS;
}

static inline int MPI_Reduce(const void* sendBuffer, void* receiveBuffer, int count, MPI_Datatype type,
                             MPI_Op op, int root, MPI_Comm comm) {
  (void)type; (void)op;
  if (count > MLBM_SHIM_MAXRED) { fprintf(stderr, "mlbm mpi shim: reduce too large\n"); MPI_Abort(comm, 4); }
  const double* mine = (const double*)(sendBuffer == MPI_IN_PLACE ? receiveBuffer : sendBuffer);
  if (mlbm_shim.size == 1) {
    if (sendBuffer != MPI_IN_PLACE && sendBuffer != receiveBuffer)
      memcpy(receiveBuffer, sendBuffer, (size_t)count * sizeof(double));
    return MPI_SUCCESS;
  }
  mlbm_shim_control* c = mlbm_shim.control;
  for (int i = 0; i < count; ++i) c->reduction[mlbm_shim.rank][i] = mine[i];
  MPI_Barrier(comm);
  if (mlbm_shim.rank == root) {
    double* out = (double*)receiveBuffer;
    for (int i = 0; i < count; ++i) {
      double sum = c->reduction[0][i];
      for (int r = 1; r < mlbm_shim.size; ++r) sum += c->reduction[r][i];
      out[i] = sum;
    }
  }
  MPI_Barrier(comm);
  return MPI_SUCCESS;
}

static inline int MPI_Scatter(const void* sendBuffer, int sendCount, MPI_Datatype sendType,
                              void* receiveBuffer, int receiveCount, MPI_Datatype receiveType,
                              int root, MPI_Comm comm) {
  (void)receiveCount; (void)receiveType; (void)root;
  if (mlbm_shim.size != 1) { fprintf(stderr, "mlbm mpi shim: MPI_Scatter is single-rank only\n"); MPI_Abort(comm, 5); }
  if (sendBuffer != receiveBuffer) memcpy(receiveBuffer, sendBuffer, (size_t)sendCount * (size_t)sendType);
  return MPI_SUCCESS;
}

static inline int MPI_Gather(const void* sendBuffer, int sendCount, MPI_Datatype sendType,
                             void* receiveBuffer, int receiveCount, MPI_Datatype receiveType,
                             int root, MPI_Comm comm) {
  (void)receiveCount; (void)receiveType; (void)root;
  if (mlbm_shim.size != 1) { fprintf(stderr, "mlbm mpi shim: MPI_Gather is single-rank only\n"); MPI_Abort(comm, 5); }
  if (sendBuffer != receiveBuffer) memcpy(receiveBuffer, sendBuffer, (size_t)sendCount * (size_t)sendType);
  return MPI_SUCCESS;
}
