// metaLBM/MPIInitializer.h (B200 drop-in) -- `MPIInitializer<numProcs>` with the reference's static members
// (MPIInitializer.h:14-77): hostName, size, rank, rankLeft, rankRight.
//
// On one 8xB200 box the ranks are plain processes, one per GPU, started by any launcher that exports a rank and
// a world size (torchrun RANK/WORLD_SIZE, OpenMPI OMPI_COMM_WORLD_*, PMI_*, SLURM_*, or MLBM_RANK/MLBM_NRANKS).
// The only host-side collective the step path needs is shipping 128 bytes (the NCCL id) from rank 0 to the
// others; with real MPI (-DLBM_B200_USE_MPI) that is an MPI_Bcast, otherwise a file under a rendezvous
// directory all ranks of the box can see.
#pragma once

#include <unistd.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>

#ifdef LBM_B200_USE_MPI
#include <mpi.h>
#endif

#include "Commons.h"
#include "MathVector.h"
#include "Options.h"

namespace lbm {

namespace b200 {
inline int envInt(std::initializer_list<const char*> names, int fallback) {
  for (const char* name : names) {
    const char* value = std::getenv(name);
    if (value && *value) return std::atoi(value);
  }
  return fallback;
}
}  // namespace b200

template <int numProcsAtCompileTile>
struct MPIInitializer {
  static std::string hostName;
  static MathVector<int, 3> size;
  static MathVector<int, 3> rank;
  static int rankLeft;
  static int rankRight;
  static int rankTop;
  static int rankBottom;
  static int rankFront;
  static int rankBack;

  MPIInitializer(int argc, char** argv) {
    (void)argc; (void)argv;
    char name[256] = {0};
    gethostname(name, sizeof(name) - 1);
    hostName = name;
#ifdef LBM_B200_USE_MPI
    int provided;
    MPI_Init_thread(&argc, &argv, MPI_THREAD_FUNNELED, &provided);
    MPI_Comm_size(MPI_COMM_WORLD, &size[d::X]);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank[d::X]);
#else
    size[d::X] = b200::envInt({"MLBM_NRANKS", "WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS"}, 1);
    rank[d::X] = b200::envInt({"MLBM_RANK", "RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID"}, 0);
#endif
    if (size[d::X] != numProcsAtCompileTile) {
      std::cout << "Compile-time and runtime number of process don't match\n";
      std::exit(1);
    }
    rankLeft = (rank[d::X] + size[d::X] - 1) % size[d::X];
    rankRight = (rank[d::X] + 1) % size[d::X];
  }

  ~MPIInitializer() {
#ifdef LBM_B200_USE_MPI
    MPI_Finalize();
#endif
  }

  // rank 0's `bytes` -> every rank (used once, for the 128-byte NCCL id)
  static void broadcastFromRoot(void* bytes, size_t count) {
    if (size[d::X] <= 1) return;
#ifdef LBM_B200_USE_MPI
    MPI_Bcast(bytes, (int)count, MPI_BYTE, 0, MPI_COMM_WORLD);
#else
    const char* directory = std::getenv("MLBM_RENDEZVOUS_DIR");
    const char* session = std::getenv("MLBM_SESSION");
    if (!session) session = std::getenv("MASTER_PORT");
    const std::string path = std::string(directory ? directory : "/tmp") + "/metalbm_b200_" + (session ? session : "default") + ".id";
    if (rank[d::X] == 0) {
      const std::string temporary = path + ".tmp";
      { std::ofstream out(temporary, std::ios::binary | std::ios::trunc); out.write(static_cast<const char*>(bytes), (std::streamsize)count); }
      std::rename(temporary.c_str(), path.c_str());
    } else {
      for (int attempt = 0; attempt < 60000; ++attempt) {
        std::ifstream in(path, std::ios::binary);
        if (in && in.read(static_cast<char*>(bytes), (std::streamsize)count)) return;
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
      }
      std::fprintf(stderr, "[%s:%d] rendezvous file %s never appeared\n", __FILE__, __LINE__, path.c_str());
      std::exit(-1);
    }
#endif
  }

  // every rank's `count` bytes -> `all` (size * count bytes) on every rank; used once, for the CUDA IPC handles of the
  // direct peer halos.  `salt` (the NCCL id, unique per run) keeps the rendezvous files of different runs apart.
  static void allGather(const void* mine, void* all, size_t count, const unsigned char* salt, size_t saltCount) {
    if (size[d::X] <= 1) { std::memcpy(all, mine, count); return; }
#ifdef LBM_B200_USE_MPI
    (void)salt; (void)saltCount;
    MPI_Allgather(mine, (int)count, MPI_BYTE, all, (int)count, MPI_BYTE, MPI_COMM_WORLD);
#else
    unsigned long long hash = 1469598103934665603ull;  // FNV-1a
    for (size_t i = 0; i < saltCount; ++i) hash = (hash ^ salt[i]) * 1099511628211ull;
    const char* directory = std::getenv("MLBM_RENDEZVOUS_DIR");
    const std::string base = std::string(directory ? directory : "/tmp") + "/metalbm_b200_" + std::to_string(hash) + ".peer.";
    {
      const std::string path = base + std::to_string(rank[d::X]), temporary = path + ".tmp";
      { std::ofstream out(temporary, std::ios::binary | std::ios::trunc); out.write(static_cast<const char*>(mine), (std::streamsize)count); }
      std::rename(temporary.c_str(), path.c_str());
    }
    for (int other = 0; other < size[d::X]; ++other) {
      char* target = static_cast<char*>(all) + (size_t)other * count;
      bool done = false;
      for (int attempt = 0; attempt < 60000 && !done; ++attempt) {
        std::ifstream in(base + std::to_string(other), std::ios::binary);
        if (in && in.read(target, (std::streamsize)count)) done = true;
        else std::this_thread::sleep_for(std::chrono::milliseconds(1));
      }
      if (!done) {
        std::fprintf(stderr, "[%s:%d] rendezvous file %s%d never appeared\n", __FILE__, __LINE__, base.c_str(), other);
        std::exit(-1);
      }
    }
#endif
  }
};

using MPIInit = MPIInitializer<numProcs>;

template <> std::string MPIInit::hostName = "";
template <> MathVector<int, 3> MPIInit::size = MathVector<int, 3>{{1, 1, 1}};
template <> MathVector<int, 3> MPIInit::rank = MathVector<int, 3>{{0, 0, 0}};
template <> int MPIInit::rankLeft = 0;
template <> int MPIInit::rankRight = 0;
template <> int MPIInit::rankTop = 0;
template <> int MPIInit::rankBottom = 0;
template <> int MPIInit::rankFront = 0;
template <> int MPIInit::rankBack = 0;

}  // namespace lbm
