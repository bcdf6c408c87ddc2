/* metalbm_b200.h -- C-ABI of the B200-native fused collide-and-stream (pull) step.
 *
 * This is the drop-in boundary for ONE hot path of gtauzin/metaLBM: the per-step
 * work of `lbm::Algorithm<T, AlgorithmType::Pull, Architecture::GPU, MemoryLayout::SoA,
 * PartitionningType::OneD, CommunicationType::MPI, Overlapping>` and what it calls.
 * Every entry point names the reference interface it replaces (file:line under the
 * reference tree).  The reference's C++ template spellings are kept by the header-only
 * shim `include/metaLBM_b200/`, which forwards to these functions.
 *
 * Conventions
 *  - plain C types only; one opaque `mlbm_ctx` per rank == per GPU (the reference's
 *    "one MPI rank <-> one GPU", CUDAInitializer.h:23-26);
 *  - every function returns 0 on success or a negative `mlbm_status`; the message of the
 *    last failure on the calling thread is `mlbm_last_error()`.  The reference prints and
 *    exit(-1)s (Commons.h:6-28); the C++ shim re-applies that convention;
 *  - a context is not thread-safe; all calls for one context come from one host thread
 *    (same as the reference);
 *  - there is NO CPU fallback: creating a context without a usable CUDA device fails.
 */
#ifndef METALBM_B200_H
#define METALBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLBM_ABI_VERSION 2

typedef struct mlbm_ctx mlbm_ctx;

typedef enum mlbm_status {
  MLBM_OK = 0,
  MLBM_ERR_INVALID = -1,      /* bad argument / unsupported combination */
  MLBM_ERR_CUDA = -2,         /* CUDA runtime failure (or no device)     */
  MLBM_ERR_COMM = -3,         /* NCCL / peer-access failure              */
  MLBM_ERR_STATE = -4,        /* call made in the wrong state            */
  MLBM_ERR_NOMEM = -5
} mlbm_status;

/* LatticeType (Options.h:13-14, descriptors Lattice.h:80,145,460,535,614).  The multi-speed lattices D2Q13 (:213),
 * D2Q17 (:290), D2Q21 (:372) and D3Q33 (:706) -- jumps of up to 3 nodes, their own sound speeds, TruncationMa3 only --
 * carry dimH halo planes per side and exchange them over NCCL (direct peer halos are for the single-speed lattices).
 * D1Q3 (:22) does not compile in the reference (Force.h:329) and is not rebuilt. */
typedef enum mlbm_lattice {
  MLBM_D2Q5 = 0, MLBM_D2Q9 = 1, MLBM_D3Q15 = 2, MLBM_D3Q19 = 3, MLBM_D3Q27 = 4,
  MLBM_D2Q13 = 5, MLBM_D2Q17 = 6, MLBM_D2Q21 = 7, MLBM_D3Q33 = 8
} mlbm_lattice;

/* CollisionType (Options.h:35-38; Collision.h:103-180 BGK, :182-376 ELBM).
 * In the reference snapshot Approached_ELBM (:378-459), Malaspinas_ELBM (:461-539), Essentially1/2_ELBM (:543-682),
 * ForcedNR_ELBM (:684-724) and ForcedBNR_ELBM (:864-909) only override the PRIVATE, non-virtual calculateAlpha, which the
 * inherited Collision<ELBM>::calculateRelaxationTime (:227-241) never calls: they compile and run bit-identically to
 * ELBM (checked against the compiled reference, tests/test_oracle_vs_reference.py) and share its kernel here.
 * ForcedNR_ELBM_Forcing (:727-857) is a different algorithm (alpha solved on the FORCED populations f + S with the
 * mirror functor EntropicStep.h:65-108, no small-deviation shortcut) with its own kernel variant, pinned by golden
 * vectors of the reference. */
typedef enum mlbm_collision {
  MLBM_BGK = 0, MLBM_ELBM = 1, MLBM_FORCED_NR_ELBM = 2, MLBM_APPROACHED_ELBM = 3, MLBM_MALASPINAS_ELBM = 4,
  MLBM_ESSENTIALLY1_ELBM = 5, MLBM_ESSENTIALLY2_ELBM = 6, MLBM_FORCED_BNR_ELBM = 7, MLBM_FORCED_NR_ELBM_FORCING = 8
} mlbm_collision;

/* EquilibriumType (Options.h:33; Equilibrium.h:14-34 TruncationMa3, :36-126 Exact) */
typedef enum mlbm_equilibrium {
  MLBM_TRUNCATION_MA3 = 0, MLBM_EXACT = 1
} mlbm_equilibrium;

/* ForcingSchemeType (Options.h:40; ForcingScheme.h:41-198) */
typedef enum mlbm_forcing_scheme {
  MLBM_SCHEME_NONE = 0, MLBM_GUO = 1, MLBM_SHAN_CHEN = 2, MLBM_EXACT_DIFFERENCE = 3
} mlbm_forcing_scheme;

/* ForceType (Options.h:41-43; Force.h:104-292).
 * MLBM_FORCE_FIELD is the generic array read Force<T, ForceType::Generic>::setForce (Force.h:39-48): the force of a node
 * is component iD of the force FIELD at the node's local index.  It is what every array-type force of the reference
 * runs through on the step path (ConstantShell, EnergyRemoval, Turbulent2D, Force.h:296-623, fill that array with FFTs
 * outside the step); here the caller supplies the array with mlbm_set_force_field.
 * MLBM_FORCE_CONSTANT_SHELL is Force<double, ForceType::ConstantShell> (Force.h:296-420) for 2-D lattices: the stream
 * function psi^(k) = force_amplitude[0] on the shell force_k_min^2 <= |k|^2 <= force_k_max^2 (integer wave numbers),
 * made solenoidal by MakeIncompressible (Transformer.h:300-384: F^ = (i k_y psi^, -i k_x psi^)), transformed back and
 * divided by the volume; the array is synthesised once on the device when the context is created and then read like
 * MLBM_FORCE_FIELD.  (The reference's 3-D variant corrupts its heap -- Force.h:341-355 writes mirrored indices of a
 * padded local array -- and is not rebuilt; a 3-D shell force goes through MLBM_FORCE_FIELD.)
 * MLBM_FORCE_ENERGY_REMOVAL is Force<double, ForceType::EnergyRemoval> (Force.h:423-561), 2-D: F_d = -force_amplitude[d] x
 * (the momentum rho u_d of the LAST STORED fields, band-passed to the shell force_k_min..force_k_max).  The reference
 * recomputes it with two FFTs in every iterate from fieldList (Force.h:552-558), which only changes on stored steps; here
 * it is recomputed after every step stored with bit 0 of is_stored, as a projection onto the shell's few modes (a
 * reduction + one NCCL all-reduce of 4 doubles per mode) and a synthesis, with no distributed transform.
 * MLBM_FORCE_TURBULENT_2D is Force<double, ForceType::Turbulent2D> (Force.h:564-616): ConstantShell(force_amplitude,
 * force_k_min/max) + EnergyRemoval(removal_amplitude, removal_k_min/max). */
typedef enum mlbm_force {
  MLBM_FORCE_NONE = 0, MLBM_FORCE_CONSTANT = 1, MLBM_FORCE_SINUSOIDAL = 2, MLBM_FORCE_KOLMOGOROV = 3, MLBM_FORCE_FIELD = 4,
  MLBM_FORCE_CONSTANT_SHELL = 5, MLBM_FORCE_ENERGY_REMOVAL = 6, MLBM_FORCE_TURBULENT_2D = 7
} mlbm_force;

/* `dataT` (Input_prod.in:10).  F32 is FP32 storage of populations and fields with the moments,
 * forcing and the entropic solve carried in FP64 registers (the reference never ran FP32). */
typedef enum mlbm_dtype { MLBM_F64 = 0, MLBM_F32 = 1 } mlbm_dtype;

/* Overlapping (Options.h:20): Off = exchange, then one kernel over the slab (Algorithm.h:326-358);
 * On = boundary planes first, exchange overlapped with the bulk kernel (the intent of Algorithm.h:392-447). */
typedef enum mlbm_overlap { MLBM_OVERLAP_OFF = 0, MLBM_OVERLAP_ON = 1 } mlbm_overlap;

/* The compile-time globals of src/Input_*.in that the path reads, as one runtime struct.
 * Kernels stay compile-time specialised; the combination is dispatched once in mlbm_create. */
typedef struct mlbm_config {
  int32_t abi_version;        /* MLBM_ABI_VERSION */
  int32_t lattice;            /* mlbm_lattice        <- latticeT        */
  int32_t collision;          /* mlbm_collision      <- collisionT      */
  int32_t equilibrium;        /* mlbm_equilibrium    <- equilibriumT    */
  int32_t forcing_scheme;     /* mlbm_forcing_scheme <- forcingSchemeT  */
  int32_t force;              /* mlbm_force          <- forceT          */
  int32_t dtype;              /* mlbm_dtype          <- dataT           */
  int32_t overlap;            /* mlbm_overlap        <- overlappingT    */
  int32_t global_length[3];   /* globalLengthX/Y/Z; unused trailing dimensions = 1 */
  int32_t rank;               /* MPIInit::rank[d::X]  (MPIInitializer.h:53)       */
  int32_t nranks;             /* numProcs; must divide global_length[0] (Domain.h:22-24) */
  int32_t device;             /* CUDA ordinal, or -1 = rank % device count (CUDAInitializer.h:23-26) */
  int32_t variant;            /* 0 = default kernel; other values select experimental kernels (bench only) */
  double tau;                 /* relaxationTime  */
  double force_amplitude[3];  /* forceAmplitude  */
  double force_wavelength[3]; /* forceWaveLength */
  int32_t force_k_min;        /* forcekMin (shell forces, Force.h:303, 315) */
  int32_t force_k_max;        /* forcekMax */
  int32_t removal_k_min;      /* removalForcekMin (Turbulent2D, Force.h:575-577) */
  int32_t removal_k_max;      /* removalForcekMax */
  double removal_amplitude[3];/* removalForceAmplitude */
} mlbm_config;

const char* mlbm_last_error(void);
int mlbm_abi_version(void);

/* Algorithm ctor (Algorithm.h:69-95, 225-240, 317-324) + Distribution ctor (Distribution.h:25-28) +
 * FieldList ctor (FieldList.h:44-63): allocates the ping-pong SoA population pair and the
 * density / velocity / alpha / force fields on the device, alpha initialised to 2 (Initialize.h:82-88). */
int mlbm_create(const mlbm_config* config, mlbm_ctx** out);
int mlbm_destroy(mlbm_ctx* ctx);

/* Multi-GPU wiring; replaces MPIInitializer (MPIInitializer.h:27-58) and the Communication ctor.
 * Rank 0 makes an id, the caller ships the 128 bytes to every rank (any transport), every rank attaches.
 * Without it a context with nranks > 1 refuses to step. */
int mlbm_comm_unique_id(void* id128);
int mlbm_comm_init(mlbm_ctx* ctx, const void* id128);

/* The device-independent part of one launch of the fused kernel over local planes x0 + i * plane_step, i < x1 - x0:
 * grid, block, dynamic shared memory and every scalar kernel parameter, exactly as mlbm_step assembles them (pure host
 * logic, no device needed).  The CPU test-suite pins these against the configuration: a launch with, say, beta = 0
 * or without the periodic wrap would still run at full speed and only a GPU parity test could tell. */
typedef struct mlbm_launch_plan {
  int32_t grid[3];
  int32_t block;
  int32_t shared_bytes;
  int32_t x0, plane_step, plane_count, planes_per_block;
  int32_t local_length[3];      /* LX, NM, NR: extents on the kernel axes (slab, middle, unit stride) */
  int32_t wrap_x;               /* 1: single rank, x wraps inside the slab; 0: halo planes hold the neighbours' data */
  int32_t is_stored, hydro_shift;
  int32_t has_force;            /* 0: none, 1: per-axis profiles of the analytic forces, 2: read from the force field */
  uint64_t stride, plane;       /* elements between populations / between x planes */
  double beta;                  /* 1 / (2 tau)                 (Collision.h:122) */
  double guo_factor;            /* (1 - 1/(2 tau)) * inv_cs2   (ForcingScheme.h:115; inv_cs2 of the lattice) */
} mlbm_launch_plan;
int mlbm_launch_plan_for(const mlbm_config* config, int x0, int x1, int is_stored, int plane_step, mlbm_launch_plan* out);

/* Direct peer halos (optional, one process per GPU on one NVLink/NVSwitch box).  Every rank exports a
 * MLBM_PEER_HANDLE_BYTES blob (CUDA IPC handles of its two population buffers and of its handshake words), the
 * caller ships the blobs between ranks (any transport, like the NCCL id) and every rank attaches the blobs of its
 * LEFT and RIGHT ring neighbours (MPIInitializer.h:56-57).  From then on, with overlap == MLBM_OVERLAP_ON, the kernel
 * that computes the two boundary planes stores their outgoing populations straight into the neighbours' halo planes
 * over NVLink -- the whole of Communication::communicateHalos (Communication.h:134-180, 494-500) without a single
 * copy or send/recv -- while the bulk kernel runs; NCCL is only used for the first exchange after an upload, for the
 * observables and as the shutdown barrier.  All ranks must issue the same sequence of calls (as with MPI). */
#define MLBM_PEER_HANDLE_BYTES 256
int mlbm_comm_peer_export(mlbm_ctx* ctx, void* handle);
int mlbm_comm_peer_attach(mlbm_ctx* ctx, const void* left_handle, const void* right_handle);

/* The halo exchange of one step as data (pure host logic, no device needed): which planes of the SoA buffer
 * this rank sends and receives, in issue order.  Mirrors Communication::sendAndReceiveHaloXRight / XLeft
 * (Communication.h:134-180): populations faceQ+1..2*faceQ (c_x > 0) travel right, 1..faceQ (c_x < 0) left.
 * Offsets and counts are in elements of the context dtype, relative to the buffer base. */
typedef struct mlbm_halo_message {
  int32_t population;   /* iQ */
  int32_t peer;         /* rank of the neighbour (MPIInitializer.h:56-57) */
  int32_t is_send;      /* 1 = send, 0 = receive */
  int32_t reserved;
  uint64_t offset;
  uint64_t count;
} mlbm_halo_message;
int mlbm_halo_plan(const mlbm_config* config, mlbm_halo_message* out, int capacity, int* count);

/* Algorithm::unpack (Algorithm.h:141-147, Boundary.h:26-38): host local-padded SoA -> device.
 * Element (iQ, x, y, z) of the host array is host[iQ*component_stride + (x*padded_y + y)*padded_z + z]
 * (lSD::getIndex, Domain.h:88-91; component stride FFTWInit::numberElements), in the context's dtype. */
int mlbm_upload_distribution(mlbm_ctx* ctx, const void* host, size_t component_stride,
                             size_t padded_y, size_t padded_z);
/* Algorithm::pack (Algorithm.h:132-139, Boundary.h:11-24): device -> host, same layout. */
int mlbm_download_distribution(mlbm_ctx* ctx, void* host, size_t component_stride,
                               size_t padded_y, size_t padded_z);
/* Distribution::getHaloDataPrevious() (Distribution.h:19-20, 30-31) as a HOST array: the buffer the next step reads,
 * in the reference's halo space hSD (Domain.h:173-283): element (iQ, iP) at host[iQ * hSD::volume() + hSD::getIndex(iP)],
 * extents local length + 2 dimH in every used dimension.  All halo cells hold what the reference's halo exchange
 * (Communication.h:134-180) and periodic boundaries (Boundary.h:45-102) would have put there before the node update
 * -- the periodic image, or the neighbour rank's plane in x -- for EVERY population, a superset of the cells the
 * reference fills.  This is what the per-node host functions of the template layer read (Moment<T>::calculateDensity /
 * calculateVelocity, Moment.h:14-47; Collision::calculateMoments, Collision.h:72-79).  `capacity` = elements of `host`
 * (>= dimQ * hSD::volume()), in the context's dtype.  An inspection path: one whole-buffer copy plus a host loop. */
int mlbm_download_halo_distribution(mlbm_ctx* ctx, void* host, size_t capacity);

/* Initial state without a host round trip: f = feq(rho, u) (initDistribution, Initialize.h:106-117) from
 * host density / velocity fields laid out like the FieldList arrays (velocity component iD at
 * velocity + iD*component_stride). */
int mlbm_init_equilibrium(mlbm_ctx* ctx, const void* density, const void* velocity,
                          size_t component_stride, size_t padded_y, size_t padded_z);

/* The synthetic initial field of the benchmarks (no reference counterpart; SURVEY.md 8d "Init B") evaluated on the
 * device at global coordinates, f = feq(rho, u) (initDistribution, Initialize.h:106-117), without host field arrays:
 *   3-D: rho = 1 + a sin X cos Y cos Z,  u = b (sin X cos Y cos Z, -cos X sin Y cos Z, cos X cos Y sin Z / 2)
 *   2-D: rho = 1 + a sin X cos Y,        u = b (sin Y, cos X)            with X = 2 pi x / globalLengthX etc. */
int mlbm_init_synthetic(mlbm_ctx* ctx, double density_amplitude, double velocity_amplitude);

/* Synthetic non-equilibrium state for benchmarks and size-independent property tests (no reference counterpart;
 * SURVEY.md 8d "Init B"): multiplies every population by 1 + eps * n, n uniform with unit variance from a
 * counter-based hash of (seed, iQ, global node index) -- the same field whatever the number of ranks. */
int mlbm_perturb_distribution(mlbm_ctx* ctx, double eps, uint64_t seed);

/* The alpha field the entropic solve warm-starts from (Algorithm.h:103-106; initAlpha, Initialize.h:82-88). */
int mlbm_set_alpha(mlbm_ctx* ctx, const void* host, size_t padded_y, size_t padded_z);

/* The force FIELD of a context created with MLBM_FORCE_FIELD (zero until set): [D] components laid out like the
 * FieldList arrays (component iD at host + iD*component_stride, element (x, y, z) at (x*padded_y + y)*padded_z + z), this
 * rank's slab.  Replaces the `forcePtr` that Force<Generic>::setForce reads (Force.h:39-48) and that
 * Force::update / setForceArray fill (Force.h:51-54, 323-331, 452-560); may be called between steps (time-dependent
 * forces).  On stored steps the step writes the same values back into the stored force field (Algorithm.h:186-190). */
int mlbm_set_force_field(mlbm_ctx* ctx, const void* host, size_t component_stride, size_t padded_y, size_t padded_z);

/* Algorithm::iterate (Algorithm.h:326-358 / 392-447): swap, halo exchange, periodic boundaries, fused node
 * update; synchronous like the reference (returns after the device finished).  `is_stored` is
 * Algorithm::isStored (Routine.h:122-124): when non-zero the step also stores density, hydrodynamic
 * velocity, alpha and force (Algorithm::storeFields, Algorithm.h:150-194) and reduces the observables.
 * Bit 0 (value 1, the reference's `true`) = fields + all observables; value 2 = energy / mass / Mach only, no field
 * arrays are allocated or written (what fits when the populations fill the GPU). */
int mlbm_step(mlbm_ctx* ctx, unsigned iteration, int is_stored);

/* `count` calls of iterate for iterations first..first+count-1 with isStored = (iteration % store_every == 0)
 * (store_every == 0: never), enqueued without host synchronisation in between; returns immediately. */
int mlbm_run_async(mlbm_ctx* ctx, unsigned first_iteration, unsigned count, unsigned store_every);
/* The same with the `is_stored` value of the stored steps chosen by the caller (1 or 2, see mlbm_step). */
int mlbm_run_async_stored(mlbm_ctx* ctx, unsigned first_iteration, unsigned count, unsigned store_every, int stored_mode);
int mlbm_sync(mlbm_ctx* ctx);

/* FieldList arrays as of the last stored step (Algorithm::storeFields): any pointer may be NULL.
 * Multi-component fields use component_stride between components. */
int mlbm_download_fields(mlbm_ctx* ctx, void* density, void* velocity, void* alpha, void* force,
                         size_t component_stride, size_t padded_y, size_t padded_z);

/* ScalarAnalysisList::writeAnalyses (AnalysisList.h:55-73) as device reductions, of the last stored step,
 * already summed over ranks (Communication::reduce, Communication.h:76-89) and normalised by the global
 * volume (Analysis.h:30):
 *   out[0] total energy     (Analysis.h:53-61)
 *   out[1] total enstrophy  (Analysis.h:85-93) of the reference's SPECTRAL vorticity (Curl, Transformer.h:118-295,
 *          Routine.h:129-132: integer wave numbers, divided by the volume twice); needs the stored velocity field,
 *          i.e. bit 0 of is_stored -- NaN after a step stored with is_stored == 2
 *   out[2] max Mach number  max |u_hydro| / c_s
 *   out[3] total mass       sum of density (PerformanceAnalysisList mass, Routine.h:117-118) */
int mlbm_observables(mlbm_ctx* ctx, double out[4]);

/* SpectralAnalysisList::writeAnalyses (AnalysisList.h:132-170) of the fields of the last stored step (bit 0 of is_stored):
 * the energy spectrum of the stored velocity and the forcing spectrum of the force array (PowerSpectra, Analysis.h:122-177),
 * summed over ranks (Communication::reduce, AnalysisList.h:184-187).  Bin k collects the stored half-spectrum modes with
 * floor(|k|) == k, each with sum_d |a^_d|^2 of the unnormalised transform, halved where the last wave number is 0; only the
 * energy spectrum is divided by the global volume (AnalysisList.h:189).  *count = gFD::maxWaveNumber()
 * = max(Lx, min(Ly, Lz)) / 2 (FourierDomain.h:77-79 with Helpers.h:37-39); either array may be NULL; `capacity` is the
 * length of the arrays given.  Unlike the reference, the fields are left untouched (its transforms run in place). */
int mlbm_power_spectra(mlbm_ctx* ctx, double* energy_spectrum, double* forcing_spectrum, int capacity, int* count);

/* Checkpoint of the distribution in the reference's data-set layout: DistributionWriter::writeDistribution
 * (Writer.h:400-445) writes dimQ data sets "distribution<iQ>", each the padded global box gSD::pLength() of doubles with rank r
 * at the hyperslab gSD::pOffset(r); DistributionReader::readDistribution (Reader.h:119-157) reads them back.  The container
 * is a flat file (no HDF5 in this build): a 4096-byte JSON header, then the data sets back to back in that order and layout;
 * tools/checkpoint_to_hdf5.py turns it into the reference's .h5 and back.  Every rank calls with the SAME path and moves only
 * its own hyperslab (one contiguous range per data set); a file written on n ranks can be read on m.  FP32 contexts widen /
 * narrow.  `iteration` is stored in / returned from the header (may be NULL when reading).  Host-side I/O: synchronous. */
int mlbm_checkpoint_write(mlbm_ctx* ctx, const char* path, unsigned iteration);
int mlbm_checkpoint_read(mlbm_ctx* ctx, const char* path, unsigned* iteration);

/* What the entropic collision did in the last step, from the alpha field (Algorithm.h:103-106), over all ranks:
 *   out[0] fraction of nodes whose alpha differs from 2, i.e. that left the small-deviation shortcut (Collision.h:357-359)
 *   out[1] smallest alpha, out[2] largest alpha.
 * Characterises a benchmark state (how much of the grid pays for the Newton solve); BGK contexts report {0, 2, 2}. */
int mlbm_alpha_statistics(mlbm_ctx* ctx, double out[3]);

/* Work of the entropic Newton solve (solveAlpha, Collision.h:328-349; NewtonRaphsonSolver, EntropicStep.h:111-140) on THIS
 * rank, counted by the step kernel between the two calls: mode 1 starts counting (zeroes the counters); mode 0 stops and
 * returns out[0] = nodes that took the solve, out[1] = evaluations of (F, F') they needed.  The FP64 side of the entropic
 * roofline (bench.py); costs two atomics per solved node while it counts, nothing otherwise. */
int mlbm_newton_statistics(mlbm_ctx* ctx, int mode, unsigned long long out[2]);

/* Communication::reduce(T* localSumPtr, numberComponents) (Communication.h:76-89): element-wise sum of `count` host
 * doubles over all ranks, result on EVERY rank (the reference leaves it on rank 0 only); a no-op for one rank. */
int mlbm_reduce_sum(mlbm_ctx* ctx, double* values, int count);

/* Self-test hook: out[i] = the kernels' table-driven logarithm of in[i] (host arrays), so that the parity suite can pin
 * the one transcendental of the entropic solve (std::log in EntropicStep.h:31-62) against a high-precision value. */
int mlbm_selftest_log(const double* in, double* out, size_t count);

/* DynamicArray<U, Architecture::CPUPinned> (DynamicArray.cuh:88-127): page-locked host memory, which is where the
 * reference keeps the GPU build's fields (Field.h:62-79) and what makes pack/unpack run at PCIe speed. */
int mlbm_alloc_pinned(size_t bytes, void** out);
int mlbm_free_pinned(void* pointer);

/* Algorithm::getCommunicationTime / getComputationTime (Algorithm.h:128-130), seconds of the last step
 * measured with CUDA events on the device. */
int mlbm_timers(mlbm_ctx* ctx, double* communication_seconds, double* computation_seconds);

/* Device-side access for callers that already hold device memory (zero-copy interop): the SoA buffer the
 * next step will read.  Element (iQ, x, y, z) is at base[iQ*component_stride + (x + halo_x)*plane + y*row + z]. */
typedef struct mlbm_device_layout {
  void* populations;          /* device pointer, context dtype */
  size_t component_stride;    /* elements between populations  */
  size_t plane;               /* elements between x planes     */
  size_t row;                 /* elements between y rows (== 1 in 2-D where y is the unit-stride axis) */
  int32_t halo_x;             /* number of x halo planes on each side */
  int32_t local_length[3];
} mlbm_device_layout;
int mlbm_device_distribution(mlbm_ctx* ctx, mlbm_device_layout* out);

/* Bench support: number of kernels this library launched since the context was created, and the CUDA
 * stream the fused kernel is launched on (for event timing by the caller). */
int mlbm_launch_count(mlbm_ctx* ctx, uint64_t* launches);
int mlbm_stream(mlbm_ctx* ctx, void** cuda_stream);
/* Average device time (ms) of the fused kernel over the launches since the last call, from a CUDA event pair
 * around every launch on that stream.  The first call switches the per-launch events on and returns zeros. */
int mlbm_kernel_time(mlbm_ctx* ctx, double* average_ms, uint64_t* launches);

/* Device-side stopwatch on the compute stream: mlbm_mark records CUDA event `slot` (0..7) after the work
 * enqueued so far; mlbm_elapsed waits for event `to` and returns the milliseconds between two marks. */
int mlbm_mark(mlbm_ctx* ctx, int slot);
int mlbm_elapsed(mlbm_ctx* ctx, int from, int to, double* milliseconds);

#ifdef __cplusplus
}
#endif
#endif /* METALBM_B200_H */
