/* MPIDB200 -- C ABI of the B200-native MPIDForce engine.
 *
 * This is the drop-in boundary below the OpenMM plugin's kernel contract
 *   class CalcMPIDForceKernel  (reference: openmmapi/include/openmm/mpidKernels.h:50-104)
 * Every entry point takes plain pointers and sizes; nothing here knows about OpenMM, torch or
 * Python.  The C++ platform kernel (mpidopenmmplugin_b200/plugin/, B200CalcMPIDForceKernel) and the
 * Python host (mpidopenmmplugin_b200/api.py) both sit on top of exactly these calls.
 *
 * Array orderings are those of MPIDForce::getMultipoleParameters
 *   (reference: openmmapi/include/openmm/MPIDForce.h:262-290, platforms/reference/src/
 *    MPIDReferenceKernels.cpp:84-177):
 *   dipoles[3N]      x y z
 *   quadrupoles[6N]  XX XY YY XZ YZ ZZ
 *   octopoles[10N]   XXX XXY XYY YYY XXZ XYZ YYZ XZZ YZZ ZZZ
 *   alphas[3N]       molecular-frame polarizability diagonal
 * Units: nm, kJ/mol, elementary charges.
 *
 * All functions return 0 on success; on failure they return non-zero and mpidb200_last_error()
 * describes the problem (the C++ plugin layer turns that into an OpenMMException).  There is no CPU
 * fallback: creating a handle without a usable CUDA device fails.
 */
#ifndef MPIDB200_H_
#define MPIDB200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mpidb200_engine* mpidb200_handle;

/* MPIDForce::NonbondedMethod / PolarizationType / precision (MPIDForce.h:58-95) */
enum { MPIDB200_NOCUTOFF = 0, MPIDB200_PME = 1 };
enum { MPIDB200_MUTUAL = 0, MPIDB200_DIRECT = 1, MPIDB200_EXTRAPOLATED = 2 };
enum { MPIDB200_MIXED = 0, MPIDB200_DOUBLE = 1 };
/* induced-dipole solver for MPIDB200_MUTUAL: DIIS reproduces the reference's iteration
 * (MPIDReferenceForce.cpp:1182-1252); CG is the preconditioned conjugate-gradient alternative. */
enum { MPIDB200_SOLVER_DIIS = 0, MPIDB200_SOLVER_CG = 1 };

typedef struct {
    int    num_particles;
    int    nonbonded_method;          /* MPIDForce::getNonbondedMethod            */
    int    polarization_type;         /* MPIDForce::getPolarizationType           */
    double cutoff;                    /* MPIDForce::getCutoffDistance   (PME only) */
    double ewald_alpha;               /* MPIDForce::getPMEParameters; 0 = derive from ewald_tolerance */
    int    grid[3];                   /*   "            "            ; 0 = derive  */
    double ewald_tolerance;           /* MPIDForce::getEwaldErrorTolerance        */
    double default_thole_width;       /* MPIDForce::getDefaultTholeWidth          */
    double scale14;                   /* MPIDForce::get14ScaleFactor              */
    int    max_iterations;            /* MPIDForce::getMutualInducedMaxIterations */
    double target_epsilon;            /* MPIDForce::getMutualInducedTargetEpsilon */
    int    num_extrapolation_coefficients;
    double extrapolation_coefficients[8];  /* MPIDForce::getExtrapolationCoefficients */
    int    precision;                 /* MPIDB200_MIXED | MPIDB200_DOUBLE (platform property "Precision") */
    int    solver;                    /* MPIDB200_SOLVER_*                         */
    int    device;                    /* CUDA device ordinal                       */
    int    frameless_alpha_fix;       /* 0: reference behaviour (atoms without a z anchor have zero
                                         lab-frame polarizability, MPIDReferenceForce.cpp:795-800) */
} mpidb200_config;

const char* mpidb200_last_error(void);

/* Fill a config with MPIDForce's defaults (openmmapi/src/MPIDForce.cpp:43-50). */
void mpidb200_default_config(mpidb200_config* cfg);

/* replaces CalcMPIDForceKernel::initialize (mpidKernels.h:68) -- part 1: allocate the engine */
int mpidb200_create(const mpidb200_config* cfg, mpidb200_handle* out);
void mpidb200_destroy(mpidb200_handle h);

/* replaces CalcMPIDForceKernel::initialize / copyParametersToContext (mpidKernels.h:68,96):
 * per-particle parameters, MPIDReferenceKernels.cpp:84-141 */
int mpidb200_set_particles(mpidb200_handle h, const double* charges, const double* dipoles,
                           const double* quadrupoles, const double* octopoles,
                           const int* axis_types, const int* atom_z, const int* atom_x, const int* atom_y,
                           const double* tholes, const double* alphas);

/* covalent maps as CSR: for MPIDForce::CovalentType t (0..7) atom i owns
 * indices[offsets[t*(N+1)+i] .. offsets[t*(N+1)+i+1])   (MPIDForce::getCovalentMaps, MPIDForce.h:322-339) */
int mpidb200_set_covalent_maps(mpidb200_handle h, const int* offsets, const int* indices);

/* periodic box vectors (ContextImpl::getPeriodicBoxVectors; MPIDReferenceKernels.cpp:188-198) */
int mpidb200_set_box(mpidb200_handle h, const double* a, const double* b, const double* c);

/* replaces CalcMPIDForceKernel::execute (mpidKernels.h:77): positions[3N] in, energy out, forces
 * ACCUMULATED into forces[3N] (MPIDReferenceKernels.cpp:225-239).  Host pointers; the copies are part
 * of the call.  forces may be NULL when include_forces is 0. */
int mpidb200_execute(mpidb200_handle h, const double* positions, int include_forces, int include_energy,
                     double* energy, double* forces);

/* Same evaluation with positions / forces already resident on the engine's device (double[3N]);
 * forces are accumulated.  Asynchronous on the engine's stream except for the solver's convergence
 * read-back; *energy is valid on return. */
int mpidb200_execute_device(mpidb200_handle h, const double* d_positions, int include_forces, int include_energy,
                            double* energy, double* d_forces);

/* Page-lock a caller-owned host array (cudaHostRegister) so that mpidb200_execute moves it by DMA without a staging
 * copy: positions straight from the caller's array, forces uploaded beside the evaluation, accumulated ON THE DEVICE and
 * written back by one DMA.  The platform kernel calls this once for the Context's position and force vectors (they
 * live as long as the Context: platforms/reference/src/MPIDReferenceKernels.cpp:43-66).  The buffer must stay
 * allocated until mpidb200_unpin_host_buffer / mpidb200_destroy.  Arrays that were not pinned still work (staged). */
int mpidb200_pin_host_buffer(mpidb200_handle h, void* buffer, unsigned long long bytes);
int mpidb200_unpin_host_buffer(mpidb200_handle h, void* buffer);

/* Partitioned host I/O for several ranks (after mpidb200_comm_init; collective: every rank sets the same value).
 * With enable = 1 the host arrays passed to mpidb200_execute / mpidb200_get_*_dipoles keep their full length 3N, but
 * rank r reads only the positions of ITS block of atoms from its array, gathers the other blocks from the other ranks
 * device to device (one all-gather over NVLink), and accumulates only its block of the forces into its array: every
 * process then moves 1/R of the bytes over its PCIe link, and the union of the R force blocks is the result the
 * reference's single-process kernel returns (platforms/reference/src/MPIDReferenceKernels.cpp:229-238).  The energy
 * is complete on every rank.  The block is atoms [first_atom, first_atom + num_atoms) with
 * first_atom = rank*ceil(N/R); enable = 0 (default): every rank reads all positions and returns all forces. */
int mpidb200_set_host_io_partition(mpidb200_handle h, int enable);
int mpidb200_get_host_io_block(mpidb200_handle h, int* first_atom, int* num_atoms);

/* The same evaluation on the device-resident data of an OpenMM CudaContext, for a kernel registered on the "CUDA"
 * platform -- no host copies at all:
 *   d_posq            cu.getPosq().getDevicePointer(): float4 (posq_is_double = 0) or double4 (1) per atom, in the
 *                     context's reordered atom order (reference: platforms/cuda/src/MPIDCudaKernels.cpp:216)
 *   d_posq_correction cu.getPosqCorrection() in mixed precision, else NULL
 *   d_atom_index      cu.getAtomIndexArray(): slot i holds atom d_atom_index[i] (MPIDCudaKernels.cpp:1089)
 *   d_force_buffer    cu.getForce(): signed 64-bit fixed point, scale 2^32, [x | y | z] x padded_num_atoms by slot;
 *                     our forces are ADDED with atomicAdd like the reference's kernels do
 *                     (platforms/cuda/src/kernels/multipoleElectrostatics.cu:708-710)
 * Run it on the context's stream with mpidb200_set_stream(h, cu.getCurrentStream()). */
int mpidb200_execute_cuda_context(mpidb200_handle h, const void* d_posq, int posq_is_double, const void* d_posq_correction,
                                  const int* d_atom_index, int padded_num_atoms, int include_forces, int include_energy,
                                  double* energy, void* d_force_buffer);

/* Run the engine on a caller-owned CUDA stream (e.g. the host framework's current stream) instead of its
 * own; pass NULL to return to the private stream.  The caller keeps ownership. */
int mpidb200_set_stream(mpidb200_handle h, void* cuda_stream);

/* replaces getInducedDipoles / getLabFramePermanentDipoles / getTotalDipoles (mpidKernels.h:79-84):
 * evaluate at `positions` and return double[3N].  which: 0 induced, 1 lab-frame permanent, 2 total */
int mpidb200_get_dipoles(mpidb200_handle h, const double* positions, int which, double* out);

/* replaces getSystemMultipoleMoments (mpidKernels.h:94): 13 values, Debye-based units
 * (MPIDReferenceForce.cpp:2349-2462).  masses[N] give the centre used as origin. */
int mpidb200_get_system_multipole_moments(mpidb200_handle h, const double* positions, const double* masses, double* out13);

/* replaces getElectrostaticPotential (mpidKernels.h:86): potential at num_points points
 * (MPIDReferenceForce.cpp:2464-2537; no octopole term, as in the reference). */
int mpidb200_get_electrostatic_potential(mpidb200_handle h, const double* positions, int num_points,
                                         const double* points, double* out);

/* replaces getPMEParameters (mpidKernels.h:103) */
int mpidb200_get_pme_parameters(mpidb200_handle h, double* alpha, int* nx, int* ny, int* nz);

/* Solver and timing statistics of the last execute: iterations, final epsilon, and CUDA-event
 * milliseconds per stage (see MPIDB200_STAGE_*). */
enum {
    MPIDB200_STAGE_SORT = 0,       /* wrap, cell sort, lab-frame moments                     */
    MPIDB200_STAGE_NLIST,          /* neighbour-list count + scan + fill                     */
    MPIDB200_STAGE_FIXED_SPREAD,   /* permanent multipoles -> grid (incl. grid clear)        */
    MPIDB200_STAGE_FFT,            /* every R2C FFT + convolution + C2R FFT of the call      */
    MPIDB200_STAGE_FIXED_GATHER,   /* 35 potential derivatives of the permanent grid         */
    MPIDB200_STAGE_FIXED_REAL,     /* real-space permanent field (+ mu = alpha.E)            */
    MPIDB200_STAGE_IND_SPREAD,     /* induced dipoles -> grid, all passes                    */
    MPIDB200_STAGE_IND_GATHER,     /* induced potential derivatives, all passes              */
    MPIDB200_STAGE_IND_REAL,       /* real-space induced field kernels, all passes           */
    MPIDB200_STAGE_SOLVER,         /* stream joins, field combination, collectives, DIIS/OPT */
    MPIDB200_STAGE_ELECTROSTATICS, /* pair energy / force / torque                           */
    MPIDB200_STAGE_FINISH,         /* reciprocal terms, torque mapping, output               */
    MPIDB200_NUM_STAGES
};
int mpidb200_get_stats(mpidb200_handle h, int* iterations, double* epsilon, double* stage_ms, long long* num_pairs);
/* Ordinary in-cutoff pairs (i<j) of the last execute by site class: out3[0] full-full (quasi-internal-frame kernel),
 * out3[1] full x bare-charge (Cartesian gather kernel), out3[2] charge-charge.  "bare charge" = a site with no
 * permanent dipole/quadrupole/octopole and no polarizability. */
int mpidb200_get_pair_class_counts(mpidb200_handle h, long long* out3);
/* per-stage CUDA-event timing: events are recorded on the engine's stream around each stage and read
 * after the call's final synchronisation (no extra synchronisation is added); off by default */
int mpidb200_set_profiling(mpidb200_handle h, int enabled);
/* Per-kernel timing: while enabled every launch of an evaluation runs alone (device drained first) between two CUDA
 * events and the durations are accumulated per kernel name; mpidb200_get_kernel_profile returns them as CSV text
 * "kernel,launches,total_us,evaluations" (call with buffer == NULL to get the size).  Measurement aid: slow. */
int mpidb200_set_kernel_profiling(mpidb200_handle h, int enabled);
int mpidb200_get_kernel_profile(mpidb200_handle h, char* buffer, long long capacity, long long* needed);
/* Work the kernels of the last execute did (this rank's share), for the rooflines:
 * out8 = { ordinary pairs, full x full, full x bare charge, charge x charge, polarizable x polarizable pairs,
 *          directed site x neighbour evaluations of the permanent-field kernel, covalently scaled pairs, polarizable sites } */
int mpidb200_get_work_counts(mpidb200_handle h, long long* out8);
/* Neighbour-list reuse: out2[0] = evaluations that sorted and searched (built the skin-padded candidate list),
 * out2[1] = evaluations that reused the order and the candidates (positions moved less than skin/2 since the build;
 * MPIDB200_SKIN, default 0.1 nm; MPIDB200_NO_LIST_REUSE=1 disables).  The pair set is exact either way. */
int mpidb200_get_list_stats(mpidb200_handle h, long long* out2);
/* Test hook: one reciprocal pass (R2C, influence function, C2R; unnormalised) of a caller-supplied real grid
 * [nx][ny][nz] in place, through the hand-written transform kernels (use_library = 0) or cuFFT (1).  Mixed precision,
 * one rank.  reference stage: fftpack_exec_3d + performMPIDReciprocalConvolution, MPIDReferenceForce.cpp:2931-2933, 3329-3366 */
int mpidb200_debug_reciprocal_pass(mpidb200_handle h, float* host_grid, int use_library);
/* FP32 FMA throughput of `device` measured with a register-resident FMA-chain kernel (TFLOP/s, 2 flop per FMA):
 * the denominator of the pair kernels' rooflines. */
int mpidb200_measure_fp32_peak(int device, double* tflops, double* seconds_per_launch);
/* number of kernel launches issued by the last execute */
long long mpidb200_last_launch_count(mpidb200_handle h);

/* Neighbour list of the last execute, for the bit-exactness tests: pairs (i<j, original indices) with
 * class 0 = ordinary, 1 = excluded (1-2, 1-3), 2 = 1-4.  Call with pairs == NULL to get the count. */
int mpidb200_get_pair_list(mpidb200_handle h, long long capacity, int* pairs_i, int* pairs_j, int* pair_class, long long* count);

/* ---- multi-GPU (one engine per rank / GPU) -------------------------------------------------------
 * Real-space rows and PME atoms are partitioned by atom block; partial fields and the charge grid are
 * summed across ranks with NCCL on the engine's stream.  unique_id is the 128-byte ncclUniqueId made by
 * mpidb200_nccl_unique_id on rank 0 and distributed by the host (e.g. torch.distributed broadcast). */
int mpidb200_nccl_unique_id(unsigned char* out128);
int mpidb200_comm_init(mpidb200_handle h, int rank, int num_ranks, const unsigned char* unique_id128);

#ifdef __cplusplus
}
#endif
#endif
