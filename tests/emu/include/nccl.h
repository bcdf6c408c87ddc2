// tests/emu/include/nccl.h -- TEST INFRASTRUCTURE ONLY: NCCL for the emulated library (tests/emu/build_context.py).
//
// One PROCESS per rank, as on the box; the "network" is a POSIX shared-memory segment named by the unique id.  Streams
// are synchronous in the emulated runtime, so an operation completes inside the call that issues it (at ncclGroupEnd for
// grouped operations).  Every rank must issue the same sequence of groups -- exactly NCCL's own rule; a rank that waits
// for more than a minute aborts, so a mismatch shows up as a failure, not as a hang.
//   group protocol: (1) sends are written to the sender's outbox, (2) barrier, (3) receives are matched in order against the
//   sources' outboxes, (4) barrier, (5) all-reduces one after the other: contribution to the outbox, barrier, reduce over the
//   ranks in rank order (deterministic), barrier.
#pragma once
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef int ncclResult_t;
enum { ncclSuccess = 0, ncclSystemError = 2, ncclInvalidArgument = 4 };
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclFloat = 7, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclMax = 2 } ncclRedOp_t;

namespace nccl_emu {
constexpr int kMaxRanks = 8;
constexpr size_t kOutboxBytes = 16u << 20;   // per rank; the emulated suites move kilobytes
constexpr int kMaxMessages = 4096;

struct Message { int destination; size_t bytes, offset; };
struct Outbox {
  std::atomic<int> messages;
  Message table[kMaxMessages];
  size_t used;
  alignas(64) unsigned char data[kOutboxBytes];
};
struct Segment {
  std::atomic<int> arrived, generation, attached;
  Outbox outbox[kMaxRanks];
};
struct Operation { int kind; const void* send; void* receive; size_t count; ncclDataType_t type; int peer; ncclRedOp_t op; };
struct Comm {
  Segment* segment;
  int nranks, rank;
  char name[128];
  int readCursor[kMaxRanks];
};
inline int& groupDepth() { static int depth = 0; return depth; }
inline std::vector<std::pair<Comm*, Operation>>& pending() { static auto* p = new std::vector<std::pair<Comm*, Operation>>(); return *p; }
inline size_t elementSize(ncclDataType_t type) { return type == ncclDouble ? 8 : 4; }

inline void barrier(Comm* c) {
  Segment* s = c->segment;
  const int generation = s->generation.load();
  if (s->arrived.fetch_add(1) + 1 == c->nranks) {
    s->arrived.store(0);
    s->generation.fetch_add(1);
    return;
  }
  const auto start = std::chrono::steady_clock::now();
  while (s->generation.load() == generation) {
    sched_yield();
    if (std::chrono::steady_clock::now() - start > std::chrono::seconds(90)) {
      std::fprintf(stderr, "nccl_emu: rank %d waited 90 s for the other ranks (mismatched sequence of collectives?)\n", c->rank);
      std::abort();
    }
  }
}

inline void execute(Comm* c, std::vector<Operation>& operations) {
  Outbox& mine = c->segment->outbox[c->rank];
  mine.messages.store(0);
  mine.used = 0;
  for (int r = 0; r < c->nranks; ++r) c->readCursor[r] = 0;
  bool pointToPoint = false;
  for (const Operation& o : operations) {
    if (o.kind != 0) { pointToPoint = pointToPoint || o.kind == 1; continue; }
    pointToPoint = true;
    const size_t bytes = o.count * elementSize(o.type);
    const int index = mine.messages.load();
    if (index >= kMaxMessages || mine.used + bytes > kOutboxBytes) { std::fprintf(stderr, "nccl_emu: outbox full\n"); std::abort(); }
    std::memcpy(mine.data + mine.used, o.send, bytes);
    mine.table[index] = Message{o.peer, bytes, mine.used};
    mine.used += (bytes + 63) / 64 * 64;
    mine.messages.store(index + 1);
  }
  if (pointToPoint) {
    barrier(c);
    for (const Operation& o : operations) {
      if (o.kind != 1) continue;
      Outbox& source = c->segment->outbox[o.peer];
      int& cursor = c->readCursor[o.peer];
      while (cursor < source.messages.load() && source.table[cursor].destination != c->rank) ++cursor;
      const size_t bytes = o.count * elementSize(o.type);
      if (cursor >= source.messages.load() || source.table[cursor].bytes != bytes) {
        std::fprintf(stderr, "nccl_emu: rank %d has no matching send from rank %d for a receive of %zu bytes\n", c->rank, o.peer, bytes);
        std::abort();
      }
      std::memcpy(o.receive, source.data + source.table[cursor].offset, bytes);
      ++cursor;
    }
    barrier(c);
  }
  for (const Operation& o : operations) {
    if (o.kind != 2) continue;
    const size_t bytes = o.count * elementSize(o.type);
    std::memcpy(mine.data, o.send, bytes);
    barrier(c);
    std::vector<unsigned char> result(bytes);
    for (size_t i = 0; i < o.count; ++i) {
      double total = 0.0;
      for (int r = 0; r < c->nranks; ++r) {
        const unsigned char* data = c->segment->outbox[r].data;
        const double v = o.type == ncclDouble ? reinterpret_cast<const double*>(data)[i] : (double)reinterpret_cast<const float*>(data)[i];
        total = r == 0 ? v : (o.op == ncclSum ? total + v : (v > total ? v : total));
      }
      if (o.type == ncclDouble) reinterpret_cast<double*>(result.data())[i] = total;
      else reinterpret_cast<float*>(result.data())[i] = (float)total;
    }
    barrier(c);   // everybody has read the contributions: the outboxes may be overwritten
    std::memcpy(o.receive, result.data(), bytes);
  }
}

inline ncclResult_t issue(Comm* c, const Operation& operation) {
  pending().push_back({c, operation});
  if (groupDepth() == 0) {
    std::vector<Operation> operations{operation};
    pending().clear();
    execute(c, operations);
  }
  return ncclSuccess;
}
}  // namespace nccl_emu

typedef nccl_emu::Comm* ncclComm_t;
typedef struct ncclConfig_v21700 { int unused; } ncclConfig_t;   // only named in the optional ncclCommSplit entry (never resolved here)

inline ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
  static int counter = 0;
  std::memset(id, 0, sizeof(*id));
  std::snprintf(id->internal, sizeof(id->internal), "/mlbm_emu_%d_%d", (int)getpid(), counter++);
  const int fd = shm_open(id->internal, O_CREAT | O_RDWR | O_EXCL, 0600);
  if (fd < 0 || ftruncate(fd, sizeof(nccl_emu::Segment)) != 0) return ncclSystemError;   // zero-filled: counters start at 0
  close(fd);
  return ncclSuccess;
}
inline ncclResult_t ncclCommInitRank(ncclComm_t* comm, int nranks, ncclUniqueId id, int rank) {
  if (nranks > nccl_emu::kMaxRanks) return ncclInvalidArgument;
  const int fd = shm_open(id.internal, O_RDWR, 0600);
  if (fd < 0) return ncclSystemError;
  void* base = mmap(nullptr, sizeof(nccl_emu::Segment), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (base == MAP_FAILED) return ncclSystemError;
  nccl_emu::Comm* c = new nccl_emu::Comm();
  c->segment = static_cast<nccl_emu::Segment*>(base);
  c->nranks = nranks;
  c->rank = rank;
  std::memcpy(c->name, id.internal, sizeof(c->name));
  nccl_emu::barrier(c);
  *comm = c;
  return ncclSuccess;
}
inline ncclResult_t ncclCommDestroy(ncclComm_t comm) {
  nccl_emu::barrier(comm);
  if (comm->rank == 0) shm_unlink(comm->name);
  munmap(comm->segment, sizeof(nccl_emu::Segment));
  delete comm;
  return ncclSuccess;
}
inline ncclResult_t ncclGroupStart() { ++nccl_emu::groupDepth(); return ncclSuccess; }
inline ncclResult_t ncclGroupEnd() {
  if (--nccl_emu::groupDepth() > 0 || nccl_emu::pending().empty()) return ncclSuccess;
  nccl_emu::Comm* c = nccl_emu::pending().front().first;
  std::vector<nccl_emu::Operation> operations;
  for (auto& entry : nccl_emu::pending()) operations.push_back(entry.second);
  nccl_emu::pending().clear();
  nccl_emu::execute(c, operations);
  return ncclSuccess;
}
inline ncclResult_t ncclSend(const void* data, size_t count, ncclDataType_t type, int peer, ncclComm_t comm, cudaStream_t) {
  return nccl_emu::issue(comm, nccl_emu::Operation{0, data, nullptr, count, type, peer, ncclSum});
}
inline ncclResult_t ncclRecv(void* data, size_t count, ncclDataType_t type, int peer, ncclComm_t comm, cudaStream_t) {
  return nccl_emu::issue(comm, nccl_emu::Operation{1, nullptr, data, count, type, peer, ncclSum});
}
inline ncclResult_t ncclAllReduce(const void* send, void* receive, size_t count, ncclDataType_t type, ncclRedOp_t op, ncclComm_t comm, cudaStream_t) {
  return nccl_emu::issue(comm, nccl_emu::Operation{2, send, receive, count, type, -1, op});
}
inline const char* ncclGetErrorString(ncclResult_t result) { return result == ncclSuccess ? "no error" : "emulated NCCL error"; }
