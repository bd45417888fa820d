/* laps_b200 — C ABI of the B200-native LAPS hot path (pseudo-spectral RHS + RK3 step).
 *
 * The reference (chenshihelio/LAPS) has no FFI layer: its Fortran driver calls FFTW through
 * `include 'fftw3.f03'` (src_compressible/fftw.f90:8) and MPI through `include 'mpif.h'`
 * (parallel.f90:3).  This header is the boundary a LAPS-style driver binds through ISO_C_BINDING
 * in place of those calls; every entry point names the reference call site it replaces
 * (file:line relative to src_compressible/).  INTEGRATION.md shows the Fortran interface module.
 *
 * Conventions: all functions return 0 on success, non-zero on failure (message via
 * laps_last_error); no exceptions cross the boundary; one host thread per handle; every device
 * allocation belongs to the library, every host buffer to the caller.  Every call makes the handle's
 * CUDA device current for the calling thread and leaves it current (as cudaSetDevice would).  Host arrays use the
 * reference layout uu(ix,iy,iz,ivar): x fastest, then local y, local z, variable slowest.
 */
#ifndef LAPS_B200_H
#define LAPS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAPS_ABI_VERSION 6
#define LAPS_MAX_RANKS 8

typedef struct laps_solver* laps_handle;

/* Flat mirror of the namelists read in mhd.f90:30-53 that the hot path depends on. */
typedef struct laps_params {
  int32_t abi_version;           /* must be LAPS_ABI_VERSION */
  int32_t nx, ny, nz;            /* &grid */
  double Lx, Ly, Lz;
  double adiabatic_index;        /* &phys */
  int32_t if_resis, if_resis_exp;
  double resistivity;
  int32_t if_visc, if_visc_exp;
  double viscosity;
  int32_t if_conserve_background; /* &numerical */
  double cfl;
  int32_t dealias_option;        /* 1: spherical 1/3 truncation, 2: compact filter, 0: none */
  double afx, afy, afz;
  int32_t if_AEB, if_corotating; /* &AEB */
  double radius0, Ur0, corotating_angle;
  int32_t if_hall;               /* &Hall */
  double ion_inertial_length;
  int32_t rank, nranks;          /* slab decomposition = ndim_parallel=1 (parallel.f90:56-58) */
  int32_t device;                /* CUDA device ordinal for this rank */
  /* 2D tree (src_compressible/2D/): ndim = 2 with nz = 1 runs the (nx, ny) algorithm of 2D/mhdrhs.f90,
   * 2D/mhd.f90:296-406 (vardt), 2D/dealiasing.f90 (dealias_option 3 = square truncation) on one GPU;
   * ndim = 0 or 3 is the 3D tree.  if_z_radial: 2D/mhd.f90:44 (&AEB); if_limit_dt_increase:
   * 2D/mhd.f90:23,396-404 (&numerical).  if_corotating (2D/mhdrhs.f90:282-288, 2D/AEBmod.f90:101-106) is supported in both
   * 2D trees (not together with if_z_radial, 2D/mhd.f90:62-67). */
  int32_t ndim;
  int32_t if_z_radial;
  int32_t if_limit_dt_increase;
  /* Incompressible tree (src_incompressible/, 3D): incompressible = 1 runs evolve of
   * src_incompressible/mhd.f90:318-367 — J and grad u (12 inverse transforms), the pressure projection
   * (mhdrhs.f90:393-518), E = -u x B (+ Hall), calc_rhs (:117-232), update_rho_p — and vardt of :356-457.
   * uu(8) is the PRESSURE in this tree (mhdinit.f90:210) and uu_prim has the velocity only.
   * rho0: the namelist background density (mhdinit.f90:15) that calc_gradient_velocity_real divides by. */
  int32_t incompressible;
  double rho0;
  /* 2D compressible tree only (2D/mhd.f90:43, &pert): a driver-supplied forcing of B_z, see laps_set_external_force. */
  int32_t if_external_force;
} laps_params;

/* Local extents as decompose_1d (parallel.f90:326-349) assigns them in slab mode. */
typedef struct laps_extents {
  int32_t nx, ny, nz, nxh;       /* nxh = nx/2+1 */
  int32_t z_offset, z_size;      /* real space: z in Zj(rank)   (zj_offset/zj_size) */
  int32_t y_offset, y_size;      /* Fourier space: ky in Yj(rank) (yj_offset/yj_size) */
  int32_t y_stride;              /* this rank's Fourier rows are ky = y_offset + j * y_stride, j < y_size.  1: the reference's
                                  * contiguous slabs (1-3 ranks, 2D, unmasked dealiasing, or LAPS_TUNE_CYCLIC=0).  nranks: rows
                                  * dealt round-robin (default from 4 ranks on with dealias options 1/3, or LAPS_TUNE_CYCLIC=1):
                                  * the rows the dealiasing mask keeps are then spread evenly over the ranks; real space, every
                                  * result and every driver-facing array are unaffected */
} laps_extents;

/* parallel_start + fftw_initialize + grid_initialize + arrays_initialize + AEB_initialize +
 * dealias_initialize (mhd.f90:58-99).  Grid sizes per axis: 2^k in [16, 2048], 3 * 2^k in [48, 1536], 5 * 2^k in
 * [80, 1280]; nz (3D) / ny (2D trees) may also be 8.  Anything else fails here with a message (the reference's FFTW
 * plans, fftw.f90:27-33, take any length). */
int laps_create(const laps_params* params, laps_handle* out);
/* parallel_end + fftw_finalize (mhd.f90:291-293). */
int laps_destroy(laps_handle h);
const char* laps_last_error(laps_handle h); /* h may be NULL: error of the last failed laps_create */
int laps_get_extents(laps_handle h, laps_extents* out);

/* Multi-rank wiring (replaces the communicators and subarray datatypes of parallel.f90:77-211).
 * The FFT passes on either side of transpose_yz / transpose_zy (parallel.f90:273-324) store
 * straight into the owning rank's buffer over NVLink and order themselves with device-side
 * flags, so the only host-side step is an exchange of addresses at start-up: each rank exports
 * an opaque blob (CUDA IPC handles of its exchange buffers), the driver all-gathers the blobs
 * (MPI_Allgather in a Fortran driver) and hands the concatenation, ordered by rank, back.
 * After that, laps_set_primitive, laps_evolve, laps_step, laps_vardt, laps_max_divb, laps_rms,
 * laps_invariants, laps_fft_forward and laps_fft_inverse are COLLECTIVE: every rank must call
 * them in the same order (exactly as every MPI rank of the reference does).  The driver must
 * synchronise the ranks (MPI_Barrier) before any of them calls laps_destroy.
 * Failure model: a rank that returns an error from a collective, dies, or calls the collectives in a different
 * order is FATAL FOR THE WHOLE JOB, as a failed MPI rank is for the reference.  The inter-rank waits are bounded
 * (LAPS_XCHG_TIMEOUT_S, default 120 s per wait): a rank that runs out of budget, or that fails on the host side
 * between two waits, raises an abort word on every rank; every rank's next host-side wait then returns an error
 * ("slab exchange aborted", laps_last_error), and every later call on those handles fails until laps_destroy.
 * Not needed when nranks == 1. */
#define LAPS_PEER_BLOB_BYTES 256
int laps_export_peer_blob(laps_handle h, void* blob /* LAPS_PEER_BLOB_BYTES */);
int laps_import_peer_blobs(laps_handle h, const void* blobs /* nranks * LAPS_PEER_BLOB_BYTES */);
/* Single-process multi-GPU: wire the handles of all ranks created in this process to each other
 * (peer access instead of IPC).  Each handle must then be driven by its own host thread (the collectives of the
 * ranks wait for each other); the handles may also share a device (tests/test_gpu_multirank.py runs 2-8 ranks on
 * one GPU this way — which needs CUDA_MODULE_LOADING=EAGER and one hardware queue per stream,
 * CUDA_DEVICE_MAX_CONNECTIONS=32, see tests/local_ranks.py; with one GPU per handle neither matters). */
int laps_connect_local(laps_handle* handles, int32_t nranks);

/* initial_calc_conserve_variable + transform_uu_real_to_fourier (mhd.f90:121-122):
 * uu_local = primitive rho,ux,uy,uz,bx,by,bz,p as background/perturbation_initialize or
 * read_restart leave them (mhdinit.f90:183-1036, restart.f90:17-63). */
int laps_set_primitive(laps_handle h, const double* uu_local);
/* The same from a mode table instead of a host array: field_v(x) = background[v] + sum_m Re(coef[v][m]
 * exp(i k_m.x)) for v = rho, ux, uy, uz, bx, by, bz and p = background[7] — the function the reference's
 * ipert = 6/7 hooks evaluate by summing cosines point by point (mhdinit.f90:487-829, O(modes x N^3)).  Here it is a
 * sparse spectrum and one inverse transform on the device; nothing but the table crosses PCIe.
 * k = int32 [nmodes][3] integer wave vectors with kx >= 0, coef = complex128 pairs [7][nmodes]. */
int laps_set_primitive_modes(laps_handle h, int32_t nmodes, const int32_t* k, const double* coef, const double* background);
/* evolve_radius(time) (mhd.f90:102,248; AEBmod.f90:56-73): radius, tau_exp, k_square. */
int laps_set_time(laps_handle h, double time);
/* vardt (mhd.f90:136,285,328-429): CFL limit, global min, 2 % hysteresis on *dt_inout, rkt_init. */
int laps_vardt(laps_handle h, double* dt_inout);
/* rkt_init(dt) alone (rktmod.f90:15-32) — fixed-dt runs and tests. */
int laps_rkt_init(laps_handle h, double dt);
/* evolve (mhd.f90:245,298-326): three RK stages with the coefficients armed by vardt/rkt_init.
 * Asynchronous: returns after enqueueing (multi-rank: after the last inter-rank barrier). */
int laps_evolve(laps_handle h);
/* One iteration of the Principal loop body, mhd.f90:245-248,285:
 * evolve; time += dt; evolve_radius(time); vardt.  Updates *time_inout and *dt_inout. */
int laps_step(laps_handle h, double* time_inout, double* dt_inout);
int laps_sync(laps_handle h);
/* The CUDA stream (cudaStream_t) every kernel of this handle is launched on, so that a harness can
 * bracket calls with its own CUDA events. */
int laps_get_stream(laps_handle h, void** stream_out);

/* calc_max_divB (mhd.f90:157,522-570). */
int laps_max_divb(laps_handle h, double* out);
/* Incompressible tree: calc_max_divV (src_incompressible/mhd.f90:620-668), max |k.(rho u)^| / rho0. */
int laps_max_divv(laps_handle h, double* out);
/* calc_divB_real + calc_max_divB_real and calc_divV_real + calc_max_divV_real
 * (src_incompressible/mhdrhs.f90:532-648, mhd.f90:672-732), the pair the incompressible driver prints at
 * dtrms cadence: out[0] = max |div B|, out[1] = max |div (rho u)| / rho0, both in REAL space. */
int laps_max_div_real(laps_handle h, double out[2]);
/* checkNan (src_compressible/2D/mhd.f90:242-245,563-591; src_incompressible/2D/mhd.f90:745-773; any tree here):
 * *is_nan = 1 if any uu(ix,iy,iz,1:8) on any rank is a NaN (MPI_Allreduce MAX), else 0.  The driver keeps the
 * dstep_checknan cadence and the stop. */
int laps_check_nan(laps_handle h, int32_t* is_nan);
/* 2D compressible tree with if_external_force: the field external_force(1:nx, 1:ny, 1, 1) of the user routine
 * calc_external_force_real (2D/mhdrhs.f90:480-531), which stays in the driver.  Every stage transforms it with the
 * fluxes (2D/mhdrhs.f90:216-251) and adds it to fnl(7) (:370-372) until the next call replaces it; it depends on
 * `time` only, so once per step (before laps_evolve / laps_step) is enough.  Zero until the first call.  Blocking. */
int laps_set_external_force(laps_handle h, const double* force_local);
/* Current rho0 (update_rho_p compounds it after every evolve in the expanding box, AEBmod.f90:123-134). */
int laps_get_rho0(laps_handle h, double* rho0);
/* calc_rms (mhdrms.f90:53-126): out = uu_ave(8), uu_rms(8), rho_u2(3); the driver keeps the
 * rms.dat formatting of mhdrms.f90:25,48. */
int laps_rms(laps_handle h, double out[19]);
/* Not in the reference (it has no energy / cross-helicity diagnostic): mean of uu(8), mean of
 * u.B, max |k.B^|. */
int laps_invariants(laps_handle h, double out[3]);

/* Host copy of the state for output_uu / restart (mhdoutput.f90:95-123): conserved uu (8 fields)
 * and uu_prim (ux,uy,uz,p); either pointer may be NULL. */
int laps_get_state(laps_handle h, double* uu_local, double* uu_prim_local);
/* The 8-field array output_uu writes (mhdoutput.f90:95-123): primitive != 0 -> rho, u, B, p
 * (output_primitive = .true.), else the conserved uu.  Same layout as uu_local. */
int laps_get_output(laps_handle h, double* out_local, int32_t primitive);
/* The same without blocking the run: the output array is packed into a device snapshot, stream-ordered behind the steps
 * enqueued so far, and copied to out_local on a copy stream while the steps that follow run (a dump then costs the run
 * one pointwise sweep instead of a PCIe transfer: output_uu of the reference blocks every rank in MPI-IO,
 * mhdoutput.f90:95-131).  out_local must stay valid — pinned, for the copy to overlap — until laps_output_wait returns;
 * one request in flight per handle (a second one waits for the first on the device).  The snapshot buffer (8 real
 * fields) is allocated at the first request. */
int laps_get_output_async(laps_handle h, double* out_local, int32_t primitive);
int laps_output_wait(laps_handle h);
/* Spectral state uu_fourier as complex128 pairs in the library's internal layout
 * [ivar][kx][ky_local][kz] (kz fastest) — for parity tests. */
int laps_get_spectral(laps_handle h, double* uu_fourier_local);

/* Unit-level transforms mirroring fftw.f90:42-71 / 73-103 (+136-222) for parity tests:
 * real [nfields][z_local][y][x]  <->  spectral [nfields][kx][ky_local][kz] (complex128 pairs). */
int laps_fft_forward(laps_handle h, const double* real_fields, int32_t nfields, double* spec_out);
int laps_fft_inverse(laps_handle h, const double* spec_in, int32_t nfields, double* real_out);
/* Index map of transpose_yz (parallel.f90:185-210,273-297) as this library realises it: for every
 * element of this rank's post-y-pass block (order: kx, ky, z_local with z_local fastest) the
 * destination rank and the destination linear offset in that rank's [kx][ky_local][z] block.
 * out = int64 pairs (rank, offset), nxh*ny*z_size of them. */
int laps_transpose_yz_indexmap(laps_handle h, int64_t* out);
/* The inverse direction, transpose_zy (parallel.f90:300-324), as the z pass realises it: for every element of this
 * rank's spectral block (order: kx, ky_local, z with z fastest) the rank that owns z and the linear offset in that
 * rank's [kx][ky][z_local] block.  out = int64 pairs (rank, offset), nxh*y_size*nz of them. */
int laps_transpose_zy_indexmap(laps_handle h, int64_t* out);

/* Work the passes skip exactly.  With dealias_option 1 (or 3 in 2D) every mode with kx >= *nkx, or with
 * *kymax < ky < ny - *kymax, is zeroed by the mask at the end of every stage (dealiasing.f90:87-99), so
 * the passes of a stage neither compute nor move those columns; the state is bit-identical to the
 * unpruned computation (tests/…::test_mask_pruning_is_bit_exact).  *nky_local = this rank's surviving
 * ky rows.  No pruning: *nkx = nx/2+1, *kymax = ny/2, *nky_local = y_size. */
int laps_get_pruning(laps_handle h, int32_t* nkx, int32_t* kymax, int32_t* nky_local);
/* The same, finer: with the spherical mask of option 1 the surviving (kx, ky) columns of this rank lie inside a
 * circle (*live_columns of nxh * y_size; the y and z passes visit only those), and inside a surviving column the
 * z pass neither loads nor stores the state / RK-history entries of masked kz (*live_modes of nxh * y_size * nz
 * are touched).  Bit-identical to the unpruned computation as well. */
int laps_get_pruning_counts(laps_handle h, int64_t* live_columns, int64_t* live_modes);

/* Fields transformed per RK stage: *nf forward (real fluxes -> spectra), *ni inverse (state + current density).
 * The reference transforms 18 (+1 with the expanding box) and 8 (+3 with the Hall term).  Here
 *  - the three mass fluxes are not transformed when dealiasing is on: calc_flux sets them to uu(2:4)
 *    (mhdrhs.f90:58-60), whose spectrum is the state itself;
 *  - the momentum flux tensor is symmetric, so F7, F10, F11 (mhdrhs.f90:69,74,75) read the spectra of their
 *    transposes F5, F6, F9 (:64,65,70);
 *  - the 2D tree drops the z fluxes (kz = 0).
 * 13 and 11 for 3D Hall-MHD in the expanding box.  *spec_rows = state rows updated by the main z-pass launch
 * (the continuity row runs with the current-density tasks when its fluxes come from the state). */
int laps_get_field_counts(laps_handle h, int32_t* nf, int32_t* ni, int32_t* spec_rows);

/* Device-time of the last laps_evolve/laps_step in milliseconds (CUDA events on the compute
 * stream), and the number of kernel launches it issued. */
int laps_last_step_ms(laps_handle h, float* ms, int32_t* launches);
/* Per-kernel device time of the last instrumented evolve: names and milliseconds of up to `cap`
 * launches (laps_set_profiling(h,1) inserts events around every launch). */
int laps_set_profiling(laps_handle h, int32_t on);
int laps_get_profile(laps_handle h, char* names /* cap*32 */, float* ms, int32_t cap, int32_t* count);
/* The ALGORITHMIC HBM bytes of the same launches, in the same order (what each pass has to read
 * and write once, the exactly skipped columns and modes left out; 0 for the one-CTA flag kernels) — the numerators of
 * the roofline figures bench.py prints. */
int laps_get_profile_bytes(laps_handle h, double* bytes, int32_t cap, int32_t* count);
/* Device memory this handle allocated (state, work and exchange buffers, tables). */
int laps_get_footprint(laps_handle h, int64_t* device_bytes);
/* Measurement helper: the LAPS_TUNE_* switches that select between equivalent kernels / launch shapes ("rhs", "rcg",
 * "cgz", "z", "spec", ...), settable on a live handle so that one process can time the alternatives on one state. */
int laps_set_tune(laps_handle h, const char* name, int32_t value);

#ifdef __cplusplus
}
#endif
#endif /* LAPS_B200_H */
