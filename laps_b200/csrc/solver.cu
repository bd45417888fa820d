// Host side of the library: handle, device memory plan, host-built tables (in the reference's
// FP64 operation order), launch sequencing of one RK stage, and the extern "C" ABI of
// include/laps_b200.h.
#include "../../include/laps_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <unistd.h>

#include "exchange.cuh"
#include "fft_passes.cuh"
#include "flux_fwd_x.cuh"
#include "pointwise.cuh"
#include "spectral_incomp.cuh"
#include "spectral_rhs.cuh"
#include "spectral_z.cuh"

using namespace laps;

namespace {

std::string g_create_error;

#define LAPS_CK(S, call)                                                                       \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      (S)->err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" +      \
                 std::to_string(__LINE__) + ")";                                               \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)
#define LAPS_TRY(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)

constexpr double kPi = 3.141592653589793;  // mhdinit.f90:7


// ---- per-size tile shapes --------------------------------------------------------------------
constexpr int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
constexpr int p2floor(int v) { return v < 2 ? 1 : 2 * p2floor(v / 2); }   // tiles hold a power-of-two number of lines (line lengths with an odd factor)
constexpr int cap_threads(int t, int nt) { return t * nt > 1024 ? p2floor(1024 / nt) : t; }
constexpr int tlx(int N) { return cap_threads(clampi(p2floor(256 / (N / 8)), 4, 8), N / 8); }   // complex lines per x-pass CTA
constexpr int tly(int N) { return cap_threads(clampi(p2floor(1024 / (N / 8)), 4, 8), N / 8); }  // lines per y-pass CTA (8 up to 1024-point lines: 128-byte chunks on the strided side)
constexpr int rcg(int N) { return clampi(p2floor(64 / (N / 8)), 1, 32); }                       // columns per CTA of the pipelined RHS z pass
constexpr int cgz(int N) { return clampi(p2floor(128 / (N / 8)), 1, 32); }                      // columns per z-pass CTA

// Line lengths with compiled transforms: powers of two, and 3 * 2^k, 5 * 2^k (FFTW plans any length, fftw.f90:27-33;
// these are the ones a 2/3-rule grid is usually given).  The odd-factor lengths run the composite transform of
// fft_core.cuh in every pass and the unpipelined z pass (k_spec_z).
#define LAPS_FOR_SIZES(X) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) \
  X(48) X(96) X(192) X(384) X(768) X(1536) X(80) X(160) X(320) X(640) X(1280)

bool size_supported(int n) {
  switch (n) {
#define X(N) case N:
    LAPS_FOR_SIZES(X)
#undef X
    return true;
    default: return false;
  }
}

struct ProfEntry { char name[32]; cudaEvent_t e0, e1; double bytes; };

}  // namespace

struct laps_solver {
  laps_params p;       // INTERNAL parameters: in the 2D tree the grid is held as (nx, 1, ny), see laps_create
  laps_params user;    // as given by the driver
  bool two_d = false;
  // incompressible tree (src_incompressible/): pressure projection, uu(8) = p, no energy equation
  bool incomp = false;
  double rho0 = 1.0;     // namelist background density (mhdinit.f90:15), compounded by update_rho_p (AEBmod.f90:123-134)
  double p0 = 1.0;
  double* G = nullptr;   // [9][npts] grad u (incompressible)
  bool retransform = false;   // re-derive the spectrum from the real fields at the start of every stage (mhd.f90:325)
  int xz = 0, xy = 0;  // line counts (planes, lines per plane) the x passes run over
  int nx, ny, nz, nxh, P, rank;
  int zoffs[LAPS_MAX_RANKS], zlens[LAPS_MAX_RANKS], yoffs[LAPS_MAX_RANKS], ylens[LAPS_MAX_RANKS];
  int nzl, nyl, zo, yo;
  int ystride = 1;       // global ky of local Fourier row kyl: yo + kyl * ystride (1: reference slabs; P: cyclic rows)
  bool cyclic = false;   // ky rows dealt round-robin to the ranks (balances the dealiased z pass; default from 4 ranks on)
  size_t npts;   // nx*ny*nzl      (real points per field)
  size_t ncol;   // nxh*nyl        (spectral columns)
  size_t csz;    // ncol*nz        (spectral elements per field)
  size_t w1sz;   // nxh*nzl*ny     (post-x-pass elements per field)
  int nf, ni;    // forward / inverse field counts per stage

  double* uu = nullptr;   // [8][npts] conserved
  double* J = nullptr;    // [3][npts]
  double* ext = nullptr;  // [npts] external_force(:,:,1,1) of the 2D compressible tree (2D/mhdinit.f90:146-148)
  int ext_slot = -1;      // its forward-field slot (transformed with the fluxes, 2D/mhdrhs.f90:216-251)
  void* bufX = nullptr;   // F (real fluxes) | V2
  void* bufY = nullptr;   // W1 | V1 (peer-written)
  void* bufZ = nullptr;   // W2 (peer-written)
  void* bufT = nullptr;   // two-stream schedule: local staging of the forward y pass, [peer][f][kx][kyl_p][zl]
  size_t bytesX = 0, bytesY = 0, bytesZ = 0, bytesT = 0;
  PeerTable tabT;         // the staging blocks, addressed like tabW2
  int peer_nA[LAPS_MAX_RANKS], peer_b0[LAPS_MAX_RANKS];   // every peer's rows inside the rectangle |ky| <= kymax
  cudaEvent_t ev_push[4] = {nullptr};   // chunk c of the forward fields has arrived on every rank
  cudaEvent_t ev_zrow[3] = {nullptr};   // the z-pass stores of row group g have arrived on every rank
  bool x_inflight = false;              // the exchange stream still works on the buffers (front half enqueued)
  cplx *uA = nullptr, *uB = nullptr, *rk = nullptr;   // u_fourier ping/pong, fnl_rk
  cplx *tw_x = nullptr, *tw_y = nullptr, *tw_z = nullptr;
  double *d_tab = nullptr;  // all 1-D tables in one allocation
  double *kxr, *kyr, *kze, *ksq_x, *ksq_y, *ksq_z, *dax, *day, *daz, *kzr;
  std::vector<double> h_tab;
  std::vector<double> wnx, wny, wnz;
  double* d_partial = nullptr;
  double* d_scal = nullptr;
  double* h_scal = nullptr;  // pinned
  int nblk = 0;

  // expanding box (AEBmod.f90)
  double Ur = 0, radius = 0, tau = 0, cosa = 1, sina = 0;
  bool ksq_initial = true;
  bool j_stale = true;
  double cc1[3] = {0, 0, 0}, dd1[3] = {0, 0, 0}, tstep[3] = {0, 0, 0};
  double dt = 0;
  bool have_state = false;

  PeerTable tabW2, tabV1;
  // slab exchange (exchange.cuh): this rank's flag/mailbox block, the peers' mappings, the epoch
  XchgBlock* xblk = nullptr;
  XchgPeers xp;
  unsigned long long epoch[kXchgChannels] = {0, 0, 0, 0};
  bool wired = false;
  unsigned long long* h_abort = nullptr;   // pinned: the device writes the abort code here (exchange.cuh)
  bool dead = false;                       // the slab exchange was aborted: every later call fails
  // Work skipped exactly (dealias option 1 / 3): every mode with kx >= nkx, or with kymax < ky < ny-kymax,
  // is zeroed by the mask at the end of each stage, so the passes of a stage neither compute nor
  // move those columns.  nkx = nxh and kymax = ny/2 mean "no pruning".
  int nkx = 0, kymax = 0, tune_prune = 1;
  int pr_nkyl = 0, pr_nA = 0, pr_a0 = 0, pr_b0 = 0;   // this rank's surviving ky rows (see ZParams)
  // spherical mask (option 1): per-kx largest surviving |ky| and the list of this rank's surviving columns
  int* d_kymax_x = nullptr;   // [nxh]
  int* d_colmap = nullptr;    // [pr_ncol]
  int* d_colkx = nullptr;     // [pr_ncol] kx of each entry of d_colmap
  int pr_ncol = 0;            // this rank's surviving columns
  long long pr_cols_all = 0;  // surviving (kx, ky) columns of ALL ranks (what the y passes of this rank's z slab visit)
  size_t dev_bytes = 0;       // device memory allocated by this handle
  long long pr_modes = 0;     // this rank's surviving modes (kx, ky, kz), for the traffic model of bench.py
  bool kzprune = false;       // masked modes are neither loaded nor stored by the z pass (see ZParams::kzprune)
  bool spectrum_full = true;   // the state still holds masked columns (fresh from laps_set_primitive)
  int slot[19];        // field slot the z pass reads each flux from (F1..F18, expand_term), < 0: not transformed
  int fslot[19];       // field slot calc_flux stores each flux to, < 0: not stored (differs from slot[] for the
                       // off-diagonal momentum fluxes that share the slot of their transpose, see sym_tensor)
  // The momentum flux tensor rho u_i u_j - B_i B_j + ptot delta_ij is symmetric: calc_flux forms both
  // rho u_y * u_x (mhdrhs.f90:64) and rho u_x * u_y (:69) and the reference transforms both.  Here the
  // transposed entries F7, F10, F11 read the spectrum of F5, F6, F9: 3 forward transforms fewer per stage,
  // equal to the reference up to the rounding of (m_y/rho) m_x against (m_x/rho) m_y.
  bool sym_tensor = true;
  int tune_cgz = 0, tune_z = 3, tune_rhs = 1, tune_rcg = 0;
  // The three mass fluxes are the momentum itself: their spectra are taken from the state (kZMass) instead of
  // being re-transformed.  Off with dealias_option 0, where the state keeps non-Hermitian Nyquist content that
  // the reference's real-space round trip would drop.
  bool mass_from_state = false;
  // laps_step runs the dt-independent front half of the NEXT step's first stage (with the CFL sweep fused into
  // calc_flux) before it returns: front_ready tells the next laps_evolve to skip it.  Anything that changes the
  // state, the radius or the work buffers in between clears it.
  bool front_ready = false;
  int tune_spec = 1;     // LAPS_TUNE_SPEC=0: laps_step = evolve; set_time; vardt with the separate k_cfl sweep
  int tune_zchunk = 0;   // LAPS_TUNE_ZCHUNK: z planes per interleaved calc_flux / forward-x launch pair (0 = whole slab)
  int tune_fusex = -1;   // calc_flux fused into the forward x pass (LAPS_TUNE_FUSEX): -1 = library default, 0 = never, 1 = whenever possible (nx <= 512)
  int num_sms = 148;
  double da_thresh = 0;
  void* ipc_opened[LAPS_MAX_RANKS][3];   // mappings obtained with cudaIpcOpenMemHandle
  cudaStream_t stream = nullptr;    // main stream: every API call is ordered on it
  // Second stream for the passes that store into the peers' buffers over NVLink (forward y pass, z passes) and the flag
  // barriers that order them, so that they run beside the HBM-only passes of the other field chunk (stage_overlap).
  // `cs` is the stream the pass launchers enqueue on (StreamScope), `cch` its flag channel (exchange.cuh).
  cudaStream_t xstream = nullptr;
  cudaStream_t cs = nullptr;
  int cch = 0;
  int cap_warps = 0;                 // > 0: the exchange-side launchers hold their grids to this many warps per SM (grid-stride loops)
  // asynchronous output (laps_get_output_async): device snapshot of the output array + a copy stream
  cudaStream_t ostream = nullptr;
  double* snap = nullptr;            // [8][npts], allocated at the first request
  cudaEvent_t ev_snap = nullptr, ev_out = nullptr;
  bool out_pending = false;
  cudaEvent_t ev_link[16] = {nullptr};
  int ev_next = 0;
  int tune_tly = 0;                  // LAPS_TUNE_TLY=4: half-height tiles in the y passes
  int tune_screen = 1;               // LAPS_TUNE_SCREEN=0: the signal speeds of vardt at every point (see cfl_may_raise)
  int tune_overlap = -1;             // LAPS_TUNE_OVERLAP: -1 = default (on from 2 ranks on), 0 = one stream, 1 = two streams
  int ovl_push_ctas = 2, ovl_chunks = 3;   // CTAs per SM of the transpose kernel; forward field chunks
  int ovl_y_warps = 16, ovl_z_warps = 12;  // form 1: warps per SM given to the exchange-side y pass / z passes
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_scal = nullptr;
  int launches = 0;
  bool profiling = false;
  std::vector<ProfEntry> prof;
  std::string err;
};

typedef laps_solver S;

namespace {

void decompose_1d(int n, int np, int* off, int* len) {  // parallel.f90:326-349
  const int normal = n / np;
  off[0] = 0; len[0] = normal;
  for (int i = 1; i < np; ++i) {
    off[i] = off[i - 1] + len[i - 1];
    len[i] = (i < np - 1) ? normal : n - off[i];
  }
}

std::vector<double> wave_numbers(int n, double L) {  // mhdinit.f90:79-110
  std::vector<double> k(n);
  for (int i = 1; i <= n; ++i) {
    if (i <= n / 2 + 1) k[i - 1] = 2 * kPi * (i - 1) / L;
    else k[i - 1] = 2 * kPi * (i - 1 - n) / L;
  }
  return k;
}

int pow2_part(int n) { int m = 1; while (n % (2 * m) == 0) m *= 2; return m; }
// entries of the twiddle table of an n-point line: W_n, and for n = P * 2^L with P odd > 1 also W_(2^L) behind it
// (the sub-transforms of the composite Fft, fft_core.cuh)
size_t twiddle_len(int n) { const int m = pow2_part(n); return m == n ? (size_t)n : (size_t)n + m; }

std::vector<cplx> twiddle_table(int n) {
  std::vector<cplx> t(twiddle_len(n));
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (size_t off = 0, len = n; off < t.size(); off += len, len = pow2_part(n)) {
    const int nn = (int)len;
    cplx* tt = t.data() + off;
    for (int m = 0; m < nn; ++m) {
      long double a = -two_pi * (long double)m / (long double)nn;
      tt[m] = mk((double)cosl(a), (double)sinl(a));
    }
    // exact values on the axes
    tt[0] = mk(1.0, 0.0);
    if (nn % 4 == 0) { tt[nn / 4] = mk(0.0, -1.0); tt[nn / 2] = mk(-1.0, 0.0); tt[3 * nn / 4] = mk(0.0, 1.0); }
  }
  return t;
}

double filter_1d(double k, double L, int n, double af) {  // dealiasing.f90:36-43
  const double aj = (5.0 + 6.0 * af) / 8.0;
  const double bj = (1.0 + 2.0 * af) / 2.0;
  const double cj = -(1.0 - 2 * af) / 8.0;
  const double w = k * L / n;
  return (aj + bj * std::cos(w) + cj * std::cos(2 * w)) / (1 + 2 * af * std::cos(w));
}

// (Re)build the radius-dependent tables and upload all 1-D tables.
int upload_tables(S* s) {
  const laps_params& p = s->p;
  const int nxh = s->nxh, ny = s->ny, nz = s->nz;
  s->h_tab.assign((size_t)3 * (nxh + ny + nz) + nz, 0.0);
  double* t = s->h_tab.data();
  double* kxr = t;            double* kyr = kxr + nxh;    double* kze = kyr + ny;
  double* ksq_x = kze + nz;   double* ksq_y = ksq_x + nxh; double* ksq_z = ksq_y + ny;
  double* dax = ksq_z + nz;   double* day = dax + nxh;     double* daz = day + ny;
  double* kzr = daz + nz;     // raw line wave numbers (2D tree with if_corotating)
  const double r0 = p.radius0, r = s->radius;
  for (int i = 0; i < nz; ++i) kzr[i] = s->wnz[i];
  for (int i = 0; i < nxh; ++i) kxr[i] = s->wnx[i];
  for (int i = 0; i < ny; ++i) kyr[i] = s->wny[i];
  for (int i = 0; i < nz; ++i) kze[i] = s->wnz[i] * r0 / r;          // mhdrhs.f90:192
  for (int i = 0; i < nxh; ++i) ksq_x[i] = s->wnx[i] * s->wnx[i];
  if (!s->ksq_initial && s->two_d && p.if_z_radial)                   // 2D/AEBmod.f90:109-111
    for (int i = 0; i < nxh; ++i) { const double a = s->wnx[i] * r0 / r; ksq_x[i] = a * a; }
  if (s->ksq_initial) {                                               // mhdinit.f90:114-122
    for (int i = 0; i < ny; ++i) ksq_y[i] = s->wny[i] * s->wny[i];
    for (int i = 0; i < nz; ++i) ksq_z[i] = s->wnz[i] * s->wnz[i];
  } else if (p.if_corotating && s->two_d) {                           // 2D/AEBmod.f90:101-106: the line axis carries ky
    // k_square = kx^2 c1 + ky^2 c2 + kx ky 2 cos sin (1 - (R0/R)^2): the first term from ksq_x (times c1 on the device), the
    // second from this table, the cross term per mode (corot2d_cross)
    const double c2 = s->sina * s->sina + (s->cosa * r0 / r) * (s->cosa * r0 / r);
    for (int i = 0; i < ny; ++i) ksq_y[i] = 0.0;
    for (int i = 0; i < nz; ++i) ksq_z[i] = (s->wnz[i] * s->wnz[i]) * c2;
  } else if (p.if_corotating) {                                       // AEBmod.f90:103-110
    for (int i = 0; i < ny; ++i) ksq_y[i] = s->wny[i] * s->wny[i];
    for (int i = 0; i < nz; ++i) { const double a = s->wnz[i] * r0 / r; ksq_z[i] = a * a; }
  } else {                                                            // AEBmod.f90:112-114
    for (int i = 0; i < ny; ++i) { const double a = s->wny[i] * r0 / r; ksq_y[i] = a * a; }
    for (int i = 0; i < nz; ++i) { const double a = s->wnz[i] * r0 / r; ksq_z[i] = a * a; }
  }
  if (p.dealias_option == 1) {                                        // dealiasing.f90:91-93
    for (int i = 0; i < nxh; ++i) { const double a = s->wnx[i] * p.Lx / (2 * kPi * s->nx); dax[i] = a * a; }
    for (int i = 0; i < ny; ++i) { const double a = s->wny[i] * p.Ly / (2 * kPi * ny); day[i] = a * a; }
    for (int i = 0; i < nz; ++i) { const double a = s->wnz[i] * p.Lz / (2 * kPi * nz); daz[i] = a * a; }
  } else if (p.dealias_option == 2) {                                 // dealiasing.f90:31-64
    for (int i = 0; i < nxh; ++i) dax[i] = filter_1d(s->wnx[i], p.Lx, s->nx, p.afx);
    for (int i = 0; i < ny; ++i) day[i] = filter_1d(s->wny[i], p.Ly, ny, p.afy);
    for (int i = 0; i < nz; ++i) daz[i] = filter_1d(s->wnz[i], p.Lz, nz, p.afz);
    if (s->two_d) day[0] = 1.0;                                       // 2D/dealiasing.f90:96: filtx * filty only
  } else if (p.dealias_option == 3) {                                 // 2D/dealiasing.f90:102-117 (square): per-axis flags
    for (int i = 0; i < nxh; ++i) dax[i] = std::fabs(s->wnx[i] * p.Lx / (2 * kPi * s->nx)) > (1.0 / 3.0) ? 1.0 : 0.0;
    for (int i = 0; i < ny; ++i) day[i] = 0.0;
    for (int i = 0; i < nz; ++i) daz[i] = std::fabs(s->wnz[i] * p.Lz / (2 * kPi * nz)) > (1.0 / 3.0) ? 1.0 : 0.0;
  }
  LAPS_CK(s, cudaMemcpyAsync(s->d_tab, t, s->h_tab.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  // the host vector is reused on the next call: make the copy complete before returning
  LAPS_CK(s, cudaStreamSynchronize(s->stream));
  return 0;
}

void aeb_calc(S* s) {  // AEBmod.f90:46-54
  s->tau = s->radius / s->Ur;
}

// ---- instrumentation ----------------------------------------------------------------------------
// `bytes`: ALGORITHMIC HBM bytes of the launch (what the pass has to read and write once, with
// the exactly skipped columns / modes left out) — the numerator of the roofline figures bench.py prints.
struct LaunchScope {
  S* s; int idx = -1;
  LaunchScope(S* s_, const char* name, double bytes = 0.0) : s(s_) {
    ++s->launches;
    if (s->profiling) {
      ProfEntry pe;
      pe.bytes = bytes;
      std::snprintf(pe.name, sizeof(pe.name), "%s", name);
      cudaEventCreate(&pe.e0); cudaEventCreate(&pe.e1);
      cudaEventRecord(pe.e0, s->cs);
      s->prof.push_back(pe);
      idx = (int)s->prof.size() - 1;
    }
  }
  ~LaunchScope() { if (idx >= 0) cudaEventRecord(s->prof[idx].e1, s->cs); }
};

// launches inside the scope go to `st` with flag channel `ch` and (cap > 0) at most `cap` warps per SM per exchange-side launch
int capped_ctas(const S* s, int nthreads) { return std::max(1, s->num_sms * s->cap_warps / std::max(1, nthreads / 32)); }
struct StreamScope {
  S* s; cudaStream_t prev; int prev_ch, prev_cap;
  StreamScope(S* s_, cudaStream_t st, int ch, int cap) : s(s_), prev(s_->cs), prev_ch(s_->cch), prev_cap(s_->cap_warps) { s->cs = st; s->cch = ch; s->cap_warps = cap; }
  ~StreamScope() { s->cs = prev; s->cch = prev_ch; s->cap_warps = prev_cap; }
};

int check_launch(S* s, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { s->err = std::string(what) + ": " + cudaGetErrorString(e); return 1; }
  return 0;
}

// ---- algorithmic bytes of the passes (per launch) ---------------------------------------------
// one real field of this rank's slab; the post-x-pass columns (kx < nkx survive) and the post-y-pass columns (the
// (kx, ky) columns of ALL ranks that survive) of one field over this rank's z slab
double bytes_real(const S* s) { return 8.0 * (double)s->npts; }
double bytes_xcols(const S* s, bool prune) { return 16.0 * (prune ? s->nkx : s->nxh) * (double)s->nzl * s->ny; }
double bytes_ycols(const S* s, bool prune) { return 16.0 * (prune ? (double)s->pr_cols_all : (double)s->nxh * s->ny) * s->nzl; }
double bytes_z(const S* s, const ZParams& z, int ntasks) {
  const double line = 16.0 * (double)z.ncolc * z.nz;                 // one z line of every visited column
  const double modes = z.kzprune ? 16.0 * (double)s->pr_modes : line;   // state / RK-history entries actually touched
  double b = 0.0;
  // a flux spectrum that several rows of the launch combine (the off-diagonal momentum fluxes, the components of E)
  // comes from HBM once: the CTAs of one column group run together and share it through L2
  unsigned long long seen = 0;
  auto in = [&](int f) { if (f >= 0 && !((seen >> f) & 1ull)) { seen |= 1ull << f; b += line; } };
  for (int i = 0; i < ntasks; ++i) {
    const ZTask& t = z.task[i];
    const double out = t.gout >= 0 ? line : 0.0;
    switch (t.kind) {
      case kZRhs: in(t.fa); in(t.fb); in(t.fx); in(t.fc); in(t.fc2);
                  b += out + modes * (1 + (z.read_rk ? 1 : 0) + 1 + (z.write_rk ? 1 : 0)); break;
      case kZForwardOnly: in(t.fa); b += line; break;
      case kZInverseOnly: b += line + out; break;
      case kZCurrent: b += modes + out; break;            // each B component is read by two of the three tasks: once from HBM
      case kZGrad: b += line / 3.0 + out; break;          // the three derivatives of one component share its line
      case kZDiv: b += 3 * line + out; break;
      case kZMass: b += modes * (4 + (z.read_rk ? 1 : 0) + 1 + (z.write_rk ? 1 : 0)) + out; break;
      default: break;
    }
  }
  return b;
}

// Every entry point makes the handle's device current for the calling thread (a driver thread that owns several
// handles, or a fresh thread, has another device current) and leaves it current, as cudaSetDevice itself would:
// restoring "the previous device" would create a context on device 0 in every one-process-per-GPU rank that never
// selected a device of its own.
struct DeviceGuard {
  bool ok = true;
  explicit DeviceGuard(int device) {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != device) ok = cudaSetDevice(device) == cudaSuccess;
  }
};
#define LAPS_ENTER(S)                                                                           \
  DeviceGuard guard_((S)->p.device);                                                            \
  if (!guard_.ok) { (S)->err = "cudaSetDevice(" + std::to_string((S)->p.device) + ") failed"; return 1; } \
  if ((S)->dead) { if ((S)->err.empty()) (S)->err = "the slab exchange of this handle was aborted"; return 1; }

// The abort word of the slab exchange (exchange.cuh): checked after every host-side wait.
int check_abort(S* s) {
  if (!s->h_abort) return 0;
  const unsigned long long code = *reinterpret_cast<volatile unsigned long long*>(s->h_abort);
  if (code == 0) return 0;
  const int who = (int)(code & 0xff) - 1;
  const unsigned long long why = code >> 8;
  s->dead = true;
  s->err = std::string("slab exchange aborted: ") +
           (why == kXchgTimeout ? "a wait for the peers' flags ran out of its budget (LAPS_XCHG_TIMEOUT_S) on rank " :
            why == kXchgHostFailure ? "the host side failed between two collectives on rank " : "abort raised by rank ") +
           std::to_string(who) + "; the state of every rank is undefined and the job must be torn down";
  return 1;
}
// This rank cannot go on (an error between two collectives): release the peers, which would otherwise wait for it.
void poison_peers(S* s) {
  if (s->P > 1 && s->wired && !s->dead) {
    LAPS_LAUNCH(k_xchg_abort, dim3(1), dim3(32), 0, s->stream, s->xp, kXchgHostFailure);
    (void)cudaGetLastError();
    s->dead = true;
  }
}

// ---- pass launchers ---------------------------------------------------------------------------
// Opt in to the large dynamic shared memory size and ask for exactly the shared-memory carve-out
// that `ctas` resident CTAs need: the rest of the 256 KB stays L1, which the strided sides of the
// passes rely on (measured: the maximum carve-out costs the y passes 25 % of their bandwidth, the
// default heuristic leaves the x and z passes one CTA short).
template <class K>
cudaError_t prepare_kernel(K kernel, size_t smem, int ctas) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int pct = (int)((ctas * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
  if (pct > 100) pct = 100;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}
// planes < 0: the whole slab; otherwise `in` holds the `planes` z planes starting at plane zl0 of the slab
template <int N, int TL>
int do_fwd_x_tl(S* s, const double* in, size_t fstride, int nfields, cplx* W1, bool prune, int zl0, int planes, bool scoped) {
  char name[32]; std::snprintf(name, sizeof(name), "fwd_x%d", nfields);
  typedef Tile<N, TL> T;
  LAPS_CK(s, prepare_kernel(k_fwd_x<N, TL>, T::SMEM, T::MINB));
  const int np = planes < 0 ? s->xz : planes;
  dim3 grid((unsigned)(np * (s->xy / (2 * TL))), (unsigned)nfields);
  if (scoped) {
    LaunchScope ls(s, name, nfields * ((double)np * s->xy * (8.0 * N + 16.0 * (prune ? s->nkx : s->nxh))));
    LAPS_LAUNCH((k_fwd_x<N, TL>), grid, dim3(T::NTHREADS), T::SMEM, s->cs, in, fstride, W1, s->xz, s->xy, s->tw_x,
                1.0 / N, prune ? s->nkx : s->nxh, zl0);
  } else {
    ++s->launches;
    LAPS_LAUNCH((k_fwd_x<N, TL>), grid, dim3(T::NTHREADS), T::SMEM, s->cs, in, fstride, W1, s->xz, s->xy, s->tw_x,
                1.0 / N, prune ? s->nkx : s->nxh, zl0);
  }
  return check_launch(s, "k_fwd_x");
}
// A CTA of the x passes takes 2 TL real lines of one plane; planes with fewer lines than the default tile (a 2D grid
// with ny = 8) use the half-height tile.
template <int N>
int do_fwd_x(S* s, const double* in, size_t fstride, int nfields, cplx* W1, bool prune, int zl0, int planes, bool scoped) {
  constexpr int TL = tlx(N);
  if (s->xy % (2 * TL) == 0) return do_fwd_x_tl<N, TL>(s, in, fstride, nfields, W1, prune, zl0, planes, scoped);
  if constexpr (TL == 8) {
    if (s->xy % 8 == 0) return do_fwd_x_tl<N, 4>(s, in, fstride, nfields, W1, prune, zl0, planes, scoped);
  }
  s->err = "ny must be a multiple of " + std::to_string(TL == 8 ? 8 : 2 * TL); return 1;
}

constexpr int kFuseGroups = 10;   // thread groups (fluxes in flight) per CTA of k_flux_fwd_x: 19 fluxes in two rounds
template <int N>
int do_flux_fwd_x(S* s, const FusedFluxParams& fp) {
  if constexpr (N <= 512 && Geom<N>::POW2) {
    typedef FTile<N, kFuseGroups> T;
    constexpr int ctas = (227 * 1024) / (int)(T::SMEM + 1024) < 1 ? 1 : ((227 * 1024) / (int)(T::SMEM + 1024) > 2 ? 2 : (227 * 1024) / (int)(T::SMEM + 1024));
    LAPS_CK(s, prepare_kernel(k_flux_fwd_x<N, kFuseGroups>, T::SMEM, ctas));
    LaunchScope ls(s, "flux_fwd_x", (8 + (s->p.if_hall ? 3 : 0)) * bytes_real(s) + s->nf * bytes_xcols(s, true));
    dim3 grid((unsigned)(s->nzl * (s->ny / 2)));
    LAPS_LAUNCH((k_flux_fwd_x<N, kFuseGroups>), grid, dim3(T::NTHREADS), T::SMEM, s->cs, fp);
    return check_launch(s, "k_flux_fwd_x");
  } else {
    s->err = "k_flux_fwd_x: power-of-two lines of up to 512 points only"; return 1;
  }
}

// f0: first field slot of the launch (W1 points at it; the peers' W2 bases are advanced to it here)
// staged: the lines go to this rank's staging blocks (tabT, z = this slab only) instead of the owners' W2 (k_xchg_push moves them)
template <int N, int TL>
int do_fwd_y_tl(S* s, const cplx* W1, int nfields, bool prune, int f0, int zbase, int zcount, bool staged) {
  if (zcount < 0) { zbase = 0; zcount = s->nzl; }
  PeerTable tabW2 = staged ? s->tabT : s->tabW2;
  const int dnz = staged ? s->nzl : s->nz, dzo = staged ? 0 : s->zo;
  for (int q = 0; q < s->P; ++q)
    if (tabW2.base[q]) tabW2.base[q] += (size_t)f0 * s->nxh * tabW2.len[q] * dnz;
  char name[32]; std::snprintf(name, sizeof(name), "fwd_y%d", nfields);
  typedef Tile<N, TL> T;
  LAPS_CK(s, prepare_kernel(k_fwd_y<N, TL>, T::SMEM, T::MINB));
  LaunchScope ls(s, name, nfields * (bytes_xcols(s, prune) + bytes_ycols(s, prune)) * ((double)zcount / s->nzl));
  const int ztiles = (zcount + TL - 1) / TL;
  const int ntiles = ztiles * (prune ? s->nkx : s->nxh);
  int gx = ntiles;
  if (s->cap_warps > 0) gx = std::max(1, std::min(ntiles, capped_ctas(s, T::NTHREADS) / nfields));
  dim3 grid((unsigned)gx, (unsigned)nfields);
  LAPS_LAUNCH((k_fwd_y<N, TL>), grid, dim3(T::NTHREADS), T::SMEM, s->cs, W1, tabW2, s->nzl, dnz, dzo,
              s->tw_y, 1.0 / N, s->nxh, prune ? s->kymax : N, prune ? (const int*)s->d_kymax_x : (const int*)nullptr, ntiles,
              zbase, zcount);
  return check_launch(s, "k_fwd_y");
}

template <int N>
int do_fwd_y(S* s, const cplx* W1, int nfields, bool prune, int f0, int zbase, int zcount, bool staged) {
  if constexpr (Geom<N>::POW2 && N >= 256 && N <= 1024 && tly(N) == 8) {   // tuning knob: half-height tiles (twice the CTAs per SM, 64-byte chunks)
    if (s->tune_tly == 4) return do_fwd_y_tl<N, 4>(s, W1, nfields, prune, f0, zbase, zcount, staged);
  }
  if constexpr (Geom<N>::POW2 && N >= 128 && N <= 512) {   // tuning knob: 16 planes per tile, 256-byte runs (measured at 8 GPUs: no gain over 128-byte runs)
    if (s->tune_tly == 16) return do_fwd_y_tl<N, 16>(s, W1, nfields, prune, f0, zbase, zcount, staged);
  }
  return do_fwd_y_tl<N, tly(N)>(s, W1, nfields, prune, f0, zbase, zcount, staged);
}

template <int N, int TL>
int do_inv_y_tl(S* s, const cplx* V1, cplx* V2, int nfields, bool prune) {
  char name[32]; std::snprintf(name, sizeof(name), "inv_y%d", nfields);
  typedef Tile<N, TL> T;
  LAPS_CK(s, prepare_kernel(k_inv_y<N, TL>, T::SMEM, T::MINB));
  LaunchScope ls(s, name, nfields * (bytes_xcols(s, prune) + bytes_ycols(s, prune)));
  const int ztiles = (s->nzl + TL - 1) / TL;
  dim3 grid((unsigned)(ztiles * (prune ? s->nkx : s->nxh)), (unsigned)nfields);
  LAPS_LAUNCH((k_inv_y<N, TL>), grid, dim3(T::NTHREADS), T::SMEM, s->cs, V1, V2, s->nzl, s->tw_y, s->nxh,
              prune ? s->kymax : N, prune ? (const int*)s->d_kymax_x : (const int*)nullptr);
  return check_launch(s, "k_inv_y");
}

template <int N>
int do_inv_y(S* s, const cplx* V1, cplx* V2, int nfields, bool prune) {
  if constexpr (Geom<N>::POW2 && N >= 256 && N <= 1024 && tly(N) == 8) {
    // 1024-point lines: 8-line tiles are 1024-thread CTAs, one per SM.  The forward pass wants them (128-byte runs in the
    // peers' buffers: 662 against 427 GB/s of NVLink egress at 1024^3 on 8 GPUs), the inverse pass, whose strided side is
    // local, runs better with two 4-line CTAs per SM (0.63 against 0.50 of the HBM peak, profiles/r02_multi_gpu.md)
    if (s->tune_tly == 4 || (N == 1024 && s->tune_tly == 0)) return do_inv_y_tl<N, 4>(s, V1, V2, nfields, prune);
  }
  return do_inv_y_tl<N, tly(N)>(s, V1, V2, nfields, prune);
}

template <int N, int TL>
int do_inv_x_tl(S* s, const cplx* V2, const RealDst& dst, int nfields, bool prune) {
  char name[32]; std::snprintf(name, sizeof(name), "inv_x%d", nfields);
  typedef Tile<N, TL> T;
  LAPS_CK(s, prepare_kernel(k_inv_x<N, TL>, T::SMEM, T::MINB));
  LaunchScope ls(s, name, nfields * (bytes_real(s) + bytes_xcols(s, prune)));
  dim3 grid((unsigned)(s->xz * (s->xy / (2 * TL))), (unsigned)nfields);
  LAPS_LAUNCH((k_inv_x<N, TL>), grid, dim3(T::NTHREADS), T::SMEM, s->cs, V2, dst, s->xz, s->xy, s->tw_x,
              prune ? s->nkx : s->nxh);
  return check_launch(s, "k_inv_x");
}
template <int N>
int do_inv_x(S* s, const cplx* V2, const RealDst& dst, int nfields, bool prune) {
  constexpr int TL = tlx(N);
  if (s->xy % (2 * TL) == 0) return do_inv_x_tl<N, TL>(s, V2, dst, nfields, prune);
  if constexpr (TL == 8) {
    if (s->xy % 8 == 0) return do_inv_x_tl<N, 4>(s, V2, dst, nfields, prune);
  }
  s->err = "ny must be a multiple of " + std::to_string(TL == 8 ? 8 : 2 * TL); return 1;
}

template <int N, int CG>
int do_spec_z_cg(S* s, const ZParams& zp, int ntasks, const char* name) {
  typedef ZTile<N, CG> T;
  LAPS_CK(s, prepare_kernel(k_spec_z<N, CG>, T::SMEM, T::MINB));
  LaunchScope ls(s, name, bytes_z(s, zp, ntasks));
  if (zp.ncolc == 0) return 0;   // this rank owns no surviving column
  const int ngroups = (zp.ncolc + CG - 1) / CG;
  int gx = ngroups;
  if (s->cap_warps > 0) gx = std::max(1, std::min(ngroups, capped_ctas(s, T::NTHREADS) / ntasks));
  dim3 grid((unsigned)gx, (unsigned)ntasks);
  LAPS_LAUNCH((k_spec_z<N, CG>), grid, dim3(T::NTHREADS), T::SMEM, s->cs, zp, ngroups);
  return check_launch(s, "k_spec_z");
}

template <int N>
int do_spec_z(S* s, const ZParams& zp, int ntasks, const char* name) {
  if constexpr (N == 512) {  // tuning knob for the benchmark grid (columns per CTA)
    if (s->tune_cgz == 1) return do_spec_z_cg<N, 1>(s, zp, ntasks, name);
    if (s->tune_cgz == 4) return do_spec_z_cg<N, 4>(s, zp, ntasks, name);
  }
  return do_spec_z_cg<N, cgz(N)>(s, zp, ntasks, name);
}

template <int N, int CG, int NQ>
int do_rhs_z_cg(S* s, const ZParams& zp, int ntasks) {
  typedef RTile<N, CG, NQ> T;
  LAPS_CK(s, prepare_kernel(k_rhs_z<N, CG, NQ>, T::SMEM, T::MINB));
  LaunchScope ls(s, "spec_z", bytes_z(s, zp, ntasks));
  if (zp.ncolc == 0) return 0;   // this rank owns no surviving column
  const int ngroups = (zp.ncolc + CG - 1) / CG;
  const long long nitems = (long long)ngroups * ntasks;
  long long wave = (long long)s->num_sms * T::MINB;
  if (s->cap_warps > 0) wave = std::min(wave, (long long)capped_ctas(s, T::NTHREADS));
  dim3 grid((unsigned)std::min(nitems, wave));
  LAPS_LAUNCH((k_rhs_z<N, CG, NQ>), grid, dim3(T::NTHREADS), T::SMEM, s->cs, zp, ntasks, ngroups);
  return check_launch(s, "k_rhs_z");
}

template <int N>
int do_rhs_z(S* s, const ZParams& zp, int ntasks) {
  if constexpr (!Geom<N>::POW2) return do_spec_z<N>(s, zp, ntasks, "spec_z");   // the pipelined kernel drives the power-of-two stages itself
  else {
  if constexpr (N == 512) {  // tuning knob for the benchmark grid (columns per CTA)
    if (s->tune_rcg == 2) return do_rhs_z_cg<N, 2, 2>(s, zp, ntasks);
  }
  if (s->tune_rhs == 2) return do_rhs_z_cg<N, rcg(N), 1>(s, zp, ntasks);   // one landing line, more resident columns
  return do_rhs_z_cg<N, rcg(N), 2>(s, zp, ntasks);
  }
}

template <int N>
int do_incomp_z(S* s, const ZParams& zp) {
  constexpr int CG = rcg(N);
  typedef ITile<N, CG> T;
  LAPS_CK(s, prepare_kernel(k_incomp_z<N, CG>, T::SMEM, T::MINB));
  const double iline = 16.0 * (double)zp.ncolc * zp.nz, imodes = zp.kzprune ? 16.0 * (double)s->pr_modes : iline;
  LaunchScope ls(s, "incomp_z", (3 + 5) * iline + imodes * (5 + (zp.read_rk ? 5 : 0) + 5 + (zp.write_rk ? 5 : 0)));
  if (zp.ncolc == 0) return 0;
  dim3 grid((unsigned)((zp.ncolc + CG - 1) / CG));
  LAPS_LAUNCH((k_incomp_z<N, CG>), grid, dim3(T::NTHREADS), T::SMEM, s->cs, zp);
  return check_launch(s, "k_incomp_z");
}

#define LAPS_DISPATCH(n, fn, ...)                                   \
  switch (n) {                                                      \
    case 16: return fn<16>(__VA_ARGS__);                            \
    case 32: return fn<32>(__VA_ARGS__);                            \
    case 64: return fn<64>(__VA_ARGS__);                            \
    case 128: return fn<128>(__VA_ARGS__);                          \
    case 256: return fn<256>(__VA_ARGS__);                          \
    case 512: return fn<512>(__VA_ARGS__);                          \
    case 1024: return fn<1024>(__VA_ARGS__);                        \
    case 2048: return fn<2048>(__VA_ARGS__);                        \
    case 48: return fn<48>(__VA_ARGS__);                        \
    case 96: return fn<96>(__VA_ARGS__);                        \
    case 192: return fn<192>(__VA_ARGS__);                      \
    case 384: return fn<384>(__VA_ARGS__);                      \
    case 768: return fn<768>(__VA_ARGS__);                      \
    case 1536: return fn<1536>(__VA_ARGS__);                    \
    case 80: return fn<80>(__VA_ARGS__);                        \
    case 160: return fn<160>(__VA_ARGS__);                      \
    case 320: return fn<320>(__VA_ARGS__);                      \
    case 640: return fn<640>(__VA_ARGS__);                      \
    case 1280: return fn<1280>(__VA_ARGS__);                    \
    default: s->err = "unsupported line length"; return 1;          \
  }

// The line axis of the fused spectral pass (nz; ny in the 2D trees) may also be 8 points long — one register-resident
// radix-8 stage, one thread per line (the 2D input the reference ships is 256 x 8, src_compressible/2D/mhd.input:12-13).
#define LAPS_DISPATCH_Z(n, fn, ...)                                 \
  switch (n) {                                                      \
    case 8: return fn<8>(__VA_ARGS__);                              \
    case 16: return fn<16>(__VA_ARGS__);                            \
    case 32: return fn<32>(__VA_ARGS__);                            \
    case 64: return fn<64>(__VA_ARGS__);                            \
    case 128: return fn<128>(__VA_ARGS__);                          \
    case 256: return fn<256>(__VA_ARGS__);                          \
    case 512: return fn<512>(__VA_ARGS__);                          \
    case 1024: return fn<1024>(__VA_ARGS__);                        \
    case 2048: return fn<2048>(__VA_ARGS__);                        \
    case 48: return fn<48>(__VA_ARGS__);                        \
    case 96: return fn<96>(__VA_ARGS__);                        \
    case 192: return fn<192>(__VA_ARGS__);                      \
    case 384: return fn<384>(__VA_ARGS__);                      \
    case 768: return fn<768>(__VA_ARGS__);                      \
    case 1536: return fn<1536>(__VA_ARGS__);                    \
    case 80: return fn<80>(__VA_ARGS__);                        \
    case 160: return fn<160>(__VA_ARGS__);                      \
    case 320: return fn<320>(__VA_ARGS__);                      \
    case 640: return fn<640>(__VA_ARGS__);                      \
    case 1280: return fn<1280>(__VA_ARGS__);                    \
    default: s->err = "unsupported line length"; return 1;          \
  }

int fwd_x(S* s, const double* in, size_t fstride, int nfields, cplx* W1, bool prune, int zl0 = 0, int planes = -1, bool scoped = true) {
  LAPS_DISPATCH(s->nx, do_fwd_x, s, in, fstride, nfields, W1, prune, zl0, planes, scoped)
}
int fwd_y(S* s, const cplx* W1, int nfields, bool prune, int f0 = 0, int zbase = 0, int zcount = -1, bool staged = false) {
  LAPS_DISPATCH(s->ny, do_fwd_y, s, W1, nfields, prune, f0, zbase, zcount, staged)
}

// transpose_yz of the forward fields [f0, f0 + nfc) from the staging blocks to the owners' W2 (k_xchg_push), on s->cs
int push_y(S* s, int f0, int nfc) {
  PushParams pp; std::memset(&pp, 0, sizeof(pp));
  int maxrows = 0;
  for (int q = 0; q < s->P; ++q) {
    pp.src[q] = s->tabT.base[q]; pp.dst[q] = s->tabW2.base[q];
    pp.len[q] = s->ylens[q]; pp.yoff[q] = s->yoffs[q]; pp.nA[q] = s->peer_nA[q]; pp.b0[q] = s->peer_b0[q];
    maxrows = std::max(maxrows, pp.nA[q] + (pp.len[q] - pp.b0[q]));
  }
  pp.nparts = s->P; pp.rank = s->rank; pp.ystride = s->ystride; pp.ny = s->ny; pp.nxh = s->nxh; pp.nkx = s->nkx;
  pp.nzl = s->nzl; pp.nz = s->nz; pp.zoff = s->zo; pp.f0 = f0; pp.nfc = nfc; pp.maxrows = maxrows;
  pp.kymax_x = s->d_kymax_x; pp.kymax = s->kymax;
  char name[32]; std::snprintf(name, sizeof(name), "push_y%d", nfc);
  LaunchScope ls(s, name, nfc * bytes_ycols(s, true));
  const long long warps = (long long)nfc * s->nkx * maxrows;
  const int ctas = (int)std::max<long long>(1, std::min<long long>((warps + 7) / 8, (long long)s->num_sms * std::max(1, s->ovl_push_ctas)));
  LAPS_LAUNCH(k_xchg_push, dim3((unsigned)ctas), dim3(256), 0, s->cs, pp);
  return check_launch(s, "k_xchg_push");
}
int flux_fwd_x(S* s, const FusedFluxParams& fp) { LAPS_DISPATCH(s->nx, do_flux_fwd_x, s, fp) }
bool use_fused_flux(const S* s) {
  if (s->two_d || s->incomp || s->nx > 512 || (s->ny & 1)) return false;
  if (s->tune_fusex >= 0) return s->tune_fusex != 0;
  return false;   // measured on B200 (profiles/r01c): 19 ms per stage against 12 ms for k_flux + k_fwd_x; opt-in until it wins
}
int inv_y(S* s, const cplx* V1, cplx* V2, int nfields, bool prune) { LAPS_DISPATCH(s->ny, do_inv_y, s, V1, V2, nfields, prune) }

// Forward x (+y) passes of `nfields` real fields into the z-pass input buffer W2.  In the 2D tree
// (grid held as (nx, 1, ny)) the post-x-pass layout [f][kx][1][y] IS the z-pass layout
// [f][kx][ky_local=1][line], so the x pass writes W2 directly and there is no y pass.
int forward_xy(S* s, const double* in, size_t fstride, int nfields, bool prune) {
  if (s->two_d) return fwd_x(s, in, fstride, nfields, (cplx*)s->bufZ, prune);
  LAPS_TRY(fwd_x(s, in, fstride, nfields, (cplx*)s->bufY, prune));
  return fwd_y(s, (const cplx*)s->bufY, nfields, prune);
}
int inv_x(S* s, const cplx* V2, const RealDst& dst, int nfields, bool prune) { LAPS_DISPATCH(s->nx, do_inv_x, s, V2, dst, nfields, prune) }
int spec_z(S* s, const ZParams& zp, int ntasks, const char* name) { LAPS_DISPATCH_Z(s->nz, do_spec_z, s, zp, ntasks, name) }
int rhs_z(S* s, const ZParams& zp, int ntasks) { LAPS_DISPATCH_Z(s->nz, do_rhs_z, s, zp, ntasks) }
int incomp_z(S* s, const ZParams& zp) { LAPS_DISPATCH_Z(s->nz, do_incomp_z, s, zp) }

// buffers (see the memory plan in DESIGN.md)
double* buf_F(S* s) { return (double*)s->bufX; }
cplx* buf_V2(S* s) { return (cplx*)s->bufX; }
cplx* buf_W1(S* s) { return (cplx*)s->bufY; }
cplx* buf_V1(S* s) { return (cplx*)s->bufY; }
cplx* buf_W2(S* s) { return (cplx*)s->bufZ; }
// uu_prim (4 real fields) for laps_get_state / laps_get_output: the flux work area is free between two API calls
// (the real fluxes are dead once the forward x pass has run, which is stream-ordered before this use)
double* prim_scratch(S* s) { return (double*)s->bufX; }

void fill_zparams(S* s, ZParams& z, bool prune = false) {
  std::memset(&z, 0, sizeof(z));
  if (prune) {
    z.nkyl = s->pr_nkyl; z.nA = s->pr_nA; z.a0 = s->pr_a0; z.b0 = s->pr_b0; z.ncolc = s->nkx * s->pr_nkyl;
    if (s->d_colmap) { z.colmap = s->d_colmap; z.colkx = s->d_colkx; z.ncolc = s->pr_ncol; }
    z.kzprune = s->kzprune ? 1 : 0;
  } else { z.nkyl = s->nyl; z.nA = s->nyl; z.a0 = 0; z.b0 = 0; z.ncolc = (int)s->ncol; }
  const laps_params& p = s->p;
  z.nxh = s->nxh; z.ny = s->ny; z.nyl = s->nyl; z.yoff = s->yo; z.ystride = s->ystride; z.nz = s->nz; z.ncol = (int)s->ncol;
  z.W2 = buf_W2(s); z.fstride = s->csz;
  z.u_in = s->uA; z.u_out = s->uB; z.fnl_rk = s->rk;
  z.V1 = s->tabV1; z.tw = s->tw_z;
  z.kxr = s->kxr; z.kyr = s->kyr; z.kze = s->kze;
  z.ksq_x = s->ksq_x; z.ksq_y = s->ksq_y; z.ksq_z = s->ksq_z;
  z.dax = s->dax; z.day = s->day; z.daz = s->daz;
  z.radius0 = p.radius0; z.radius = s->radius; z.cosa = s->cosa; z.sina = s->sina; z.tau = s->tau;
  const double q = p.radius0 / s->radius;
  z.ksq_c1 = s->cosa * s->cosa + (s->sina * p.radius0 / s->radius) * (s->sina * p.radius0 / s->radius);
  z.ksq_c2 = s->sina * s->sina + (s->cosa * p.radius0 / s->radius) * (s->cosa * p.radius0 / s->radius);
  z.ksq_c3 = 1 - q * q;
  z.corot_k = (p.if_AEB && p.if_corotating) ? 1 : 0;
  z.corot2d = (s->two_d && z.corot_k) ? 1 : 0;
  z.kzr = s->kzr; z.ksq_cross = z.ksq_c3;
  z.corot_ksq = (!s->ksq_initial && p.if_corotating) ? 1 : 0;
  z.aeb = p.if_AEB;
  z.visc_imp = p.if_visc && !p.if_visc_exp; z.visc_exp = p.if_visc && p.if_visc_exp;
  z.resis_imp = p.if_resis && !p.if_resis_exp; z.resis_exp = p.if_resis && p.if_resis_exp;
  z.conserve_bg = p.if_conserve_background;
  z.nu = p.viscosity; z.eta = p.resistivity;
  z.dealias_option = p.dealias_option;
  z.scale = 1.0 / s->nz;
  z.da_thresh = s->da_thresh;
  z.mode2d = s->two_d; z.z_radial = s->two_d && p.if_AEB && p.if_z_radial; z.bg_all_kz = s->two_d;
  z.tune = s->tune_z;
  z.aeb_p = 2.0 * p.adiabatic_index;   // src_incompressible/mhdrhs.f90:202-203
}

ZTask blank_task() {
  ZTask t; std::memset(&t, 0, sizeof(t));
  t.fa = t.fb = t.fx = t.fc = t.fc2 = -1; t.cf1 = 1.0; t.gout = -1;
  return t;
}

ZTask rhs_task(int v, int gout, int fa, double ca, int fb, double cb, int fx, double cx, double sg, int fc, double sc) {
  ZTask t = blank_task();
  t.kind = kZRhs; t.v = v; t.gout = gout;
  t.fa = fa; t.ca = ca; t.fb = fb; t.cb = cb; t.fx = fx; t.cx = cx; t.sg = sg; t.fc = fc; t.sc = sc;
  static const double aebc[8] = {2.0, 2.0, 3.0, 3.0, 2.0, 1.0, 1.0, 0.0};
  t.aeb_c = aebc[v];
  t.diff = (v >= 1 && v <= 3) ? 1 : ((v >= 4 && v <= 6) ? 2 : 0);
  return t;
}

// J^ = i k x B^ from the state `u`, inverse z (mhdrhs.f90:296-339); with irk >= 0 also the continuity row of
// stage irk (kZMass: reads the stage's input state u_A, writes u_B(1), fnl_rk(1) and the inverse-z output).
int launch_current_tasks(S* s, const cplx* u, bool prune, bool want_j = true, int irk = -1) {
  ZParams z; fill_zparams(s, z, prune);
  z.u_in = u;
  int n = 0;
  if (want_j)
    for (int j = 0; j < 3; ++j) {
      ZTask t = blank_task();
      t.kind = kZCurrent; t.jcomp = j; t.gout = 8 + j; t.fa = t.fb = t.fx = t.fc = -1;
      z.task[n++] = t;
    }
  if (irk >= 0) {
    z.u_old = s->uA; z.u_out = s->uB;
    z.cc = s->cc1[irk]; z.dd = s->dd1[irk]; z.dt_irk = s->tstep[irk];
    z.read_rk = (irk > 0); z.write_rk = (irk < 2);
    ZTask t = blank_task();
    t.kind = kZMass; t.v = 0; t.gout = 0; t.aeb_c = 2.0; t.fa = t.fb = t.fx = t.fc = -1;
    z.task[n++] = t;
  }
  if (n == 0) return 0;
  return spec_z(s, z, n, "curl_b_inv_z");
}

// Every rank has finished the passes enqueued so far (stands where the reference's blocking
// mpi_sendrecv loops of transpose_yz/zy return, parallel.f90:273-324).  Device side, asynchronous.
int host_barrier(S* s) {
  if (s->P > 1) {
    if (!s->wired) { s->err = "nranks > 1 but the ranks are not connected (laps_import_peer_blobs / laps_connect_local)"; return 1; }
    ++s->epoch[s->cch];
    LaunchScope ls(s, "xchg_barrier");
    LAPS_LAUNCH(k_xchg_barrier, dim3(1), dim3(32), 0, s->cs, s->xp, s->cch, s->epoch[s->cch]);
    return check_launch(s, "k_xchg_barrier");
  }
  return 0;
}

struct RealSlots { double* ptr[20]; };
RealSlots dst_state_and_current(S* s) {   // V1 slot -> real field: 0-7 uu, 8-10 J, 11-19 grad u (incompressible)
  RealSlots d; std::memset(&d, 0, sizeof(d));
  for (int v = 0; v < 8; ++v) d.ptr[v] = s->uu + (size_t)v * s->npts;
  for (int j = 0; j < 3; ++j) d.ptr[8 + j] = s->J ? s->J + (size_t)j * s->npts : nullptr;
  for (int j = 0; j < 9; ++j) d.ptr[11 + j] = s->G ? s->G + (size_t)j * s->npts : nullptr;
  return d;
}

// inverse y and x passes for V1 slots [g0, g0+n) into the real fields [r0, r0+n) of dst_state_and_current (r0 < 0: r0 = g0)
int inverse_yx(S* s, int g0, int n, bool prune, int r0 = -1) {
  if (r0 < 0) r0 = g0;
  const size_t vs = (size_t)s->nxh * s->ny * s->nzl;
  RealSlots d = dst_state_and_current(s);
  RealDst d2;
  std::memset(&d2, 0, sizeof(d2));
  for (int i = 0; i < n; ++i) d2.ptr[i] = d.ptr[r0 + i];
  if (s->two_d) return inv_x(s, buf_V1(s) + (size_t)g0 * vs, d2, n, prune);   // [g][kx][1][line] is already the x-pass layout
  LAPS_TRY(inv_y(s, buf_V1(s) + (size_t)g0 * vs, buf_V2(s) + (size_t)g0 * vs, n, prune));
  return inv_x(s, buf_V2(s) + (size_t)g0 * vs, d2, n, prune);
}

// J from the current spectral state when the cached one is stale (first stage after
// set_primitive or after the radius changed).
int refresh_current(S* s) {
  if (!s->p.if_hall || !s->j_stale) return 0;
  // the state may still hold masked columns here (first stage after laps_set_primitive): no pruning then
  const bool prune = !s->spectrum_full;
  // The tasks below store into the peers' V1 buffers: every peer must have finished the inverse y pass of the
  // previous stage (which reads V1) first.  Between two steps driven as evolve; vardt the allreduce of vardt
  // orders this; laps_step's speculative front half and back-to-back laps_evolve calls have no such call in between.
  LAPS_TRY(host_barrier(s));
  LAPS_TRY(launch_current_tasks(s, s->uA, prune));
  LAPS_TRY(host_barrier(s));
  LAPS_TRY(inverse_yx(s, 8, 3, prune));
  s->j_stale = false;
  return 0;
}

// transform_uu_real_to_fourier (fftw.f90:42-71): u_A = FFT(uu)
int spectrum_from_real(S* s, bool prune) {
  LAPS_TRY(forward_xy(s, s->uu, s->npts, 8, prune));
  LAPS_TRY(host_barrier(s));
  ZParams z; fill_zparams(s, z, prune);
  z.u_out = s->uA;
  for (int v = 0; v < 8; ++v) {
    ZTask t = blank_task();
    t.kind = kZForwardOnly; t.v = v; t.gout = -1; t.fa = v; t.fb = t.fx = t.fc = -1;
    z.task[v] = t;
  }
  LAPS_TRY(spec_z(s, z, 8, "fwd_z"));
  return host_barrier(s);
}

// One RK stage of the incompressible tree (src_incompressible/mhd.f90:323-364).
int stage_incomp(S* s, int irk) {
  const laps_params& p = s->p;
  const bool prune = !s->spectrum_full;
  if (s->retransform) LAPS_TRY(spectrum_from_real(s, prune));            // mhd.f90:325
  {  // calc_current_density_real + calc_gradient_velocity_real (mhdrhs.f90:244-391): 12 inverse transforms
    // (their stores go into the peers' V1 buffers, which the peers read in the inverse y pass that ended the
    // previous stage: wait for every rank to be past it)
    LAPS_TRY(host_barrier(s));
    ZParams z; fill_zparams(s, z, prune);
    z.u_in = s->uA;
    for (int j = 0; j < 3; ++j) {
      ZTask t = blank_task();
      t.kind = kZCurrent; t.jcomp = j; t.gout = j; t.fa = t.fb = t.fx = t.fc = -1;
      z.task[j] = t;
    }
    for (int b = 0; b < 3; ++b)
      for (int a = 0; a < 3; ++a) {
        ZTask t = blank_task();
        t.kind = kZGrad; t.v = 1 + b; t.jcomp = a; t.cx = s->rho0; t.gout = 3 + 3 * b + a; t.fa = t.fb = t.fx = t.fc = -1;
        z.task[3 + 3 * b + a] = t;
      }
    LAPS_TRY(spec_z(s, z, 12, "deriv_inv_z"));
    LAPS_TRY(host_barrier(s));
    LAPS_TRY(inverse_yx(s, 0, 12, prune, 8));   // V1 slots 0-11 -> J (3), grad u (9)
  }
  {  // calc_flux_for_pressure + calc_flux (mhdrhs.f90:393-437, 25-84)
    FluxIncParams f;
    f.uu = s->uu; f.J = s->J; f.G = s->G; f.F = buf_F(s); f.npts = s->npts;
    f.hall = p.if_hall; f.di = p.ion_inertial_length;
    LaunchScope ls(s, "flux", (7 + 3 + 9 + 6) * bytes_real(s));   // reads rho, rho u, B (the pressure is not an input), J, grad u; writes Fp, E
    LAPS_LAUNCH(k_flux_incomp, dim3((unsigned)s->nblk), dim3(256), 0, s->stream, f);
    LAPS_TRY(check_launch(s, "k_flux_incomp"));
  }
  LAPS_TRY(forward_xy(s, buf_F(s), s->npts, 6, true));
  LAPS_TRY(host_barrier(s));
  {
    ZParams z; fill_zparams(s, z, true);
    z.cc = s->cc1[irk]; z.dd = s->dd1[irk]; z.dt_irk = s->tstep[irk];
    z.read_rk = (irk > 0); z.write_rk = (irk < 2);
    // projection rows: rho u, p, rho (slots 0-2 = Fp)
    LAPS_TRY(incomp_z(s, z));
    // dB/dt = curl E (mhdrhs.f90:175-181), slots 3-5 = E
    if (!s->two_d) {
      z.task[0] = rhs_task(4, 4, -1, 0.0, 5, 1.0, -1, 0.0, -1.0, 4, +1.0);
      z.task[1] = rhs_task(5, 5, 5, 1.0, -1, 0.0, -1, 0.0, +1.0, 3, -1.0);
      z.task[2] = rhs_task(6, 6, 4, -1.0, 3, 1.0, -1, 0.0, +1.0, -1, 0.0);
    } else {   // kz = 0 (2D/mhdrhs.f90:228-233): fnl5 = -ky Ez ; fnl6 = kx Ez ; fnl7 = ky Ex - kx Ey; ky rides the line axis
      z.task[0] = rhs_task(4, 4, -1, 0.0, -1, 0.0, -1, 0.0, -1.0, 5, -1.0);
      z.task[1] = rhs_task(5, 5, 5, 1.0, -1, 0.0, -1, 0.0, +1.0, -1, 0.0);
      z.task[2] = rhs_task(6, 6, 4, -1.0, -1, 0.0, -1, 0.0, +1.0, 3, +1.0);
      if (z.corot2d) {   // if_corotating (src_incompressible/2D/mhdrhs.f90:196-201): see the compressible 2D rows in stage()
        const double cq = s->cosa, sq = s->sina * s->radius / p.radius0;
        z.task[0] = rhs_task(4, 4, -1, 0.0, 5, 1.0, -1, 0.0, -1.0, 5, -1.0);  z.task[0].cf1 = cq;
        z.task[1] = rhs_task(5, 5, 5, 1.0, -1, 0.0, -1, 0.0, +1.0, 5, +1.0);  z.task[1].cf1 = sq;
        z.task[2] = rhs_task(6, 6, 4, -1.0, 3, 1.0, -1, 0.0, +1.0, 3, +1.0);
        z.task[2].cf1 = cq; z.task[2].fc2 = 4; z.task[2].cf2 = -sq;
      }
    }
    LAPS_TRY(spec_z(s, z, 3, "spec_z"));
  }
  LAPS_TRY(host_barrier(s));
  LAPS_TRY(inverse_yx(s, 0, 8, true));
  if (s->spectrum_full) {
    if (s->nkx < s->nxh || s->kymax < s->ny / 2 || s->kzprune || s->d_colmap)
      LAPS_CK(s, cudaMemsetAsync(s->uA, 0, 8 * s->csz * sizeof(cplx), s->stream));
    s->spectrum_full = false;
  }
  std::swap(s->uA, s->uB);
  return 0;
}

int reduce_launch(S* s, int nrows, int op, double init);

void fill_cfl_params(S* s, CflParams& c) {
  const laps_params& p = s->p;
  c.uu = s->uu; c.npts = s->npts; c.gamma = p.adiabatic_index; c.di = p.ion_inertial_length;
  const double dx = p.Lx / s->nx, dy = s->two_d ? p.Lz / s->nz : p.Ly / s->ny, dz = p.Lz / s->nz;
  c.dmin = s->two_d ? std::min(dx, dy) : std::min(std::min(dx, dy), dz);
  const bool rfloor = s->two_d && p.if_resis && p.if_resis_exp;      // 2D/mhd.f90:361-364
  c.floor_x = rfloor ? p.resistivity / dx : 0.0;
  c.floor_y = rfloor ? p.resistivity / dy : 0.0;
  c.hall = p.if_hall; c.partial = s->d_partial; c.screen = s->tune_screen;
  c.row_stride = s->nblk; c.boff = 0;
}

// `to` waits for everything enqueued on `from` so far (events from a small rotating pool: a wait captures the event's
// state at the call, so re-recording an event later is harmless).
int link_streams(S* s, cudaStream_t from, cudaStream_t to) {
  cudaEvent_t e = s->ev_link[s->ev_next];
  s->ev_next = (s->ev_next + 1) % 16;
  LAPS_CK(s, cudaEventRecord(e, from));
  LAPS_CK(s, cudaStreamWaitEvent(to, e, 0));
  return 0;
}

// Two-stream schedule of a stage (3D compressible tree, several ranks).  Every pass runs on the main stream at the
// speed it has on one GPU; what the reference's transpose_yz does (parallel.f90:273-297: every rank blocked in
// mpi_sendrecv) is a light copy kernel on the exchange stream (k_xchg_push) that moves chunk c of the forward fields to
// their owners while the x and y passes of chunk c + 1 — and the first z-pass rows, which need only the first chunk —
// compute.  The flag barriers that order the exchange live on the exchange stream as well (their own channel), so the
// main stream never waits inside one: it waits for the EVENT recorded behind it, usually long after it has fired.
// The inverse direction (transpose_zy) stays fused into the z passes: their stores are 1 KB runs that NVLink takes at
// more than 800 GB/s while the kernel is bound by its own arithmetic.
// Returns 0 (one stream), 1 (two streams, the y pass and the z passes store straight into the peers' buffers from the
// exchange stream, with their grids held to a share of every SM, beside the x passes / inverse passes of the main
// stream) or 2 (two streams, transpose_yz as the copy kernel described above).  Measured on 8 B200s at 512^3
// (profiles/r02_multi_gpu.md): 10.7 ms per step on one stream, 10.4 with form 1, 11.4 with form 2; on 4: 19.5 / 18.9 /
// 21.0; on 2 the forms are within 1 % of each other.  NVLink carries about 1 GB per stage per GPU at 550-660 GB/s —
// half of the stage's compute time — and whatever runs beside the transfers loses most of what the overlap wins (the
// passes share SMs, L2 and HBM with them).  Default: form 1 from 4 ranks on, one stream below.
int use_overlap(const S* s) {
  if (s->two_d || s->incomp || s->ext_slot >= 0 || !s->xstream) return 0;
  if (use_fused_flux(s) || s->tune_zchunk > 0) return 0;
  int mode = s->tune_overlap >= 0 ? s->tune_overlap : (s->P >= 4 ? 1 : 0);   // default: form 1 from 4 ranks on
  if (mode == 2 && !s->bufT) mode = 1;
  return mode;
}

// Forward field chunks of the two-stream schedule: chunk 0 = the fluxes of the density / momentum rows (all the first
// z-pass row group needs: a prefix of the slots, which follow the flux numbering), the rest in ovl_chunks - 1 parts.
int forward_chunks(const S* s, int (&f0)[4], int (&n)[4]) {
  int cA = 0;
  for (int j = 0; j < 12; ++j) if (s->slot[j] >= 0) cA = std::max(cA, s->slot[j] + 1);
  const int parts = std::max(1, std::min(3, s->ovl_chunks - 1));
  int nc = 0;
  if (cA > 0 && cA < s->nf) { f0[nc] = 0; n[nc] = cA; ++nc; } else cA = 0;
  const int rem = s->nf - cA;
  for (int i = 0, at = cA; i < parts && at < s->nf; ++i) {
    const int m = rem / parts + (i < rem % parts ? 1 : 0);
    if (m == 0) continue;
    f0[nc] = at; n[nc] = m; ++nc; at += m;
  }
  return nc;
}

// The exchange stream is still moving forward fields (a front half enqueued by laps_step): make the main stream wait
// for it before anything else touches the work buffers.
int settle_exchange(S* s) {
  if (s->x_inflight) {
    for (int c = 0; c < 4; ++c) if (s->ev_push[c]) LAPS_CK(s, cudaStreamWaitEvent(s->stream, s->ev_push[c], 0));
    s->x_inflight = false;
  }
  return 0;
}

// The part of a stage that does not depend on the time step: J refresh, calc_flux, forward x and y passes.
// with_cfl: the CFL maxima of vardt (mhd.f90:352-416) are taken inside the calc_flux sweep (k_flux<true>) and
// left in d_partial — used by laps_step, which runs this for the NEXT step before it knows the next dt.
int stage_front(S* s, bool with_cfl) {
  const laps_params& p = s->p;
  LAPS_TRY(refresh_current(s));
  if (use_fused_flux(s)) {  // calc_flux + the forward x pass in one kernel (flux_fwd_x.cuh), then the y pass
    FusedFluxParams fp;
    std::memset(&fp, 0, sizeof(fp));
    fp.uu = s->uu; fp.J = s->J; fp.W1 = buf_W1(s); fp.npts = s->npts;
    fp.nzl = s->nzl; fp.ny = s->ny; fp.nkx = s->nkx; fp.tw = s->tw_x; fp.scale = 1.0 / s->nx;
    fp.hall = p.if_hall; fp.aeb = p.if_AEB; fp.gamma = p.adiabatic_index; fp.di = p.ion_inertial_length; fp.tau = s->tau;
    fp.nflux = s->nf;
    for (int j = 0; j < 19; ++j) if (s->fslot[j] >= 0) fp.id[s->fslot[j]] = j;
    LAPS_TRY(flux_fwd_x(s, fp));
    LAPS_TRY(fwd_y(s, buf_W1(s), s->nf, true));
  } else if (s->tune_zchunk > 0 && !s->two_d) {
    // calc_flux and the forward x pass interleaved over z chunks small enough for the chunk's fluxes to stay in
    // the L2 between the two kernels (nf x chunk planes x 8 nx ny bytes): the fluxes need not reach HBM at all
    FluxParams f;
    f.uu = s->uu; f.J = s->J; f.F = buf_F(s); f.npts = s->npts;
    f.hall = p.if_hall; f.aeb = p.if_AEB; f.gamma = p.adiabatic_index; f.di = p.ion_inertial_length; f.tau = s->tau;
    f.z_radial = 0;
    for (int j = 0; j < 19; ++j) f.slot[j] = s->fslot[j];
    const size_t plane = (size_t)s->nx * s->ny;
    const int cz = std::min(s->tune_zchunk, s->nzl);
    f.fstride = (size_t)cz * plane;
    LaunchScope ls(s, "flux+fwd_x", (8 + (p.if_hall ? 3 : 0) + 2 * s->nf) * bytes_real(s) + s->nf * bytes_xcols(s, true));
    for (int z0 = 0; z0 < s->nzl; z0 += cz) {
      const int nzc = std::min(cz, s->nzl - z0);
      f.in_off = (size_t)z0 * plane; f.count = (size_t)nzc * plane;
      const unsigned gb = (unsigned)std::min<size_t>((size_t)s->nblk, (f.count + 255) / 256);
      ++s->launches;
      LAPS_LAUNCH(k_flux<false>, dim3(gb), dim3(256), 0, s->stream, f);
      LAPS_TRY(check_launch(s, "k_flux"));
      LAPS_TRY(fwd_x(s, buf_F(s), f.fstride, s->nf, buf_W1(s), true, z0, nzc, false));
    }
    LAPS_TRY(fwd_y(s, buf_W1(s), s->nf, true));
  } else {
  {  // calc_flux (mhdrhs.f90:21-124; 2D/mhdrhs.f90:23-128)
    FluxParams f;
    f.uu = s->uu; f.J = s->J; f.F = buf_F(s); f.npts = s->npts;
    f.in_off = 0; f.count = s->npts; f.fstride = s->npts;
    f.hall = p.if_hall; f.aeb = p.if_AEB; f.gamma = p.adiabatic_index; f.di = p.ion_inertial_length; f.tau = s->tau;
    f.z_radial = s->two_d && p.if_z_radial;
    for (int j = 0; j < 19; ++j) f.slot[j] = s->fslot[j];
    int stored = 0;
    for (int j = 0; j < 19; ++j) stored += s->fslot[j] >= 0;
    LaunchScope ls(s, with_cfl ? "flux+cfl" : "flux", (8 + (p.if_hall ? 3 : 0) + stored) * bytes_real(s));
    if (with_cfl) {
      fill_cfl_params(s, f.cfl);
      LAPS_LAUNCH(k_flux<true>, dim3((unsigned)s->nblk), dim3(256), 0, s->stream, f);
    } else {
      LAPS_LAUNCH(k_flux<false>, dim3((unsigned)s->nblk), dim3(256), 0, s->stream, f);
    }
    LAPS_TRY(check_launch(s, "k_flux"));
  }
  if (with_cfl) LAPS_TRY(reduce_launch(s, 3, 2, 0.0));   // the maxima travel to the host while the passes below run
  if (s->ext_slot >= 0)   // calc_external_force_real (2D/mhdrhs.f90:129-131): the driver's field, transformed with the fluxes
    LAPS_CK(s, cudaMemcpyAsync(buf_F(s) + (size_t)s->ext_slot * s->npts, s->ext, s->npts * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
  // transform_flux_real_to_fourier (mhdrhs.f90:128-172)
  if (!use_overlap(s)) {
    LAPS_TRY(forward_xy(s, buf_F(s), s->npts, s->nf, true));
  } else if (use_overlap(s) == 1) {
    // x pass of chunk c on the main stream, y pass of chunk c (stores into the peers' W2) on the exchange stream
    const int nc = std::max(1, std::min(s->ovl_chunks, s->nf));
    for (int c = 0, f0 = 0; c < nc; ++c) {
      const int n = s->nf / nc + (c < s->nf % nc ? 1 : 0);
      LAPS_TRY(fwd_x(s, buf_F(s) + (size_t)f0 * s->npts, s->npts, n, buf_W1(s) + (size_t)f0 * s->w1sz, true));
      LAPS_TRY(link_streams(s, s->stream, s->xstream));
      {
        StreamScope sc(s, s->xstream, 1, s->P > 1 ? s->ovl_y_warps : 0);
        LAPS_TRY(fwd_y(s, buf_W1(s) + (size_t)f0 * s->w1sz, n, true, f0));
      }
      f0 += n;
    }
    LAPS_TRY(link_streams(s, s->xstream, s->stream));   // the main stream owns the buffers again (API calls are ordered on it)
  } else {
    int f0[4], n[4];
    const int nc = forward_chunks(s, f0, n);
    for (int c = 0; c < nc; ++c) {
      LAPS_TRY(fwd_x(s, buf_F(s) + (size_t)f0[c] * s->npts, s->npts, n[c], buf_W1(s) + (size_t)f0[c] * s->w1sz, true));
      LAPS_TRY(fwd_y(s, buf_W1(s) + (size_t)f0[c] * s->w1sz, n[c], true, f0[c], 0, -1, true));   // into the staging blocks
      LAPS_TRY(link_streams(s, s->stream, s->xstream));
      {
        StreamScope sc(s, s->xstream, 1, 0);
        LAPS_TRY(push_y(s, f0[c], n[c]));       // transpose_yz of this chunk, beside the passes of the next one
        LAPS_TRY(host_barrier(s));              // ... and it has arrived on every rank
      }
      LAPS_CK(s, cudaEventRecord(s->ev_push[c], s->xstream));
    }
    for (int c = nc; c < 4; ++c) LAPS_CK(s, cudaEventRecord(s->ev_push[c], s->xstream));
    s->x_inflight = true;
  }
  }
  return 0;
}

// The CFL sweep can ride on calc_flux only on the plain path (one k_flux launch over the whole slab).
bool can_speculate(const S* s) {
  // (the external force is user data of the step's own time: the driver sets it between two steps)
  return s->tune_spec && !s->incomp && s->ext_slot < 0 && !use_fused_flux(s) && !(s->tune_zchunk > 0 && !s->two_d);
}

// End of a stage: the buffer just read becomes the output buffer of the next stage.
int stage_finish(S* s, bool want_j) {
  if (s->spectrum_full) {
    // First stage after laps_set_primitive: the buffer just read still holds the unmasked initial
    // spectrum; it becomes the output buffer of the next stage, which writes surviving columns only.
    if (s->nkx < s->nxh || s->kymax < s->ny / 2 || s->kzprune || s->d_colmap)
      LAPS_CK(s, cudaMemsetAsync(s->uA, 0, 8 * s->csz * sizeof(cplx), s->stream));
    s->spectrum_full = false;
  }
  std::swap(s->uA, s->uB);
  s->j_stale = s->p.if_hall && !want_j;
  return 0;
}

// Back half of a stage in the two-stream schedule (see use_overlap).  `z` holds the `ntasks` RHS rows of the stage
// (v = 1..7, or 0..7 when the continuity row is not taken from the state); rows [0, na) = density / momentum, which need
// the first forward chunk only, the rest = B and energy.
int stage_back_push(S* s, int irk, ZParams& z, int ntasks) {
  const laps_params& p = s->p;
  const bool want_j = p.if_hall && !(irk == 2 && s->Ur != 0.0);
  const int na = ntasks - 4;                      // rows up to momentum z
  const int g0a = z.task[0].gout;                 // V1 slots of group A are g0a .. g0a + na - 1, of group B 4 .. 7
  auto z_rows = [&](int first, int n) -> int {   // RHS rows [first, first + n): stores into the peers' V1 (transpose_zy fused)
    ZParams zz = z;
    for (int i = 0; i < n; ++i) zz.task[i] = z.task[first + i];
    if (s->tune_rhs) return rhs_z(s, zz, n);
    return spec_z(s, zz, n, "spec_z");
  };
  auto arrived = [&](int g) -> int {             // flag barrier behind the rows just enqueued, on the exchange stream
    LAPS_TRY(link_streams(s, s->stream, s->xstream));
    {
      StreamScope sc(s, s->xstream, 1, 0);
      LAPS_TRY(host_barrier(s));
    }
    LAPS_CK(s, cudaEventRecord(s->ev_zrow[g], s->xstream));
    return 0;
  };
  LAPS_CK(s, cudaStreamWaitEvent(s->stream, s->ev_push[0], 0));    // chunk 0 is in W2 on every rank (and nobody reads V1 any more)
  LAPS_TRY(z_rows(0, na));
  LAPS_TRY(arrived(0));
  for (int c = 1; c < 4; ++c) LAPS_CK(s, cudaStreamWaitEvent(s->stream, s->ev_push[c], 0));
  s->x_inflight = false;
  LAPS_TRY(z_rows(na, 4));
  LAPS_TRY(arrived(1));
  const bool group_c = want_j || s->mass_from_state;
  if (group_c) {  // J for the next stage's calc_flux + the continuity row (reads the rows just updated)
    LAPS_TRY(launch_current_tasks(s, s->uB, true, want_j, s->mass_from_state ? irk : -1));
    LAPS_TRY(arrived(2));
  }
  LAPS_CK(s, cudaStreamWaitEvent(s->stream, s->ev_zrow[0], 0));
  LAPS_TRY(inverse_yx(s, g0a, na, true));
  LAPS_CK(s, cudaStreamWaitEvent(s->stream, s->ev_zrow[1], 0));
  LAPS_TRY(inverse_yx(s, 4, 4, true));
  if (group_c) {
    LAPS_CK(s, cudaStreamWaitEvent(s->stream, s->ev_zrow[2], 0));
    if (s->mass_from_state) LAPS_TRY(inverse_yx(s, 0, 1, true));
    if (want_j) LAPS_TRY(inverse_yx(s, 8, 3, true));
  }
  return stage_finish(s, want_j);
}

// Back half of a stage, two-stream form 1 (see use_overlap): the z passes store into the peers' V1 from the exchange
// stream, the inverse y and x passes of the previous row group run on the main stream beside them.
int stage_back_direct(S* s, int irk, ZParams& z, int ntasks) {
  const laps_params& p = s->p;
  const bool want_j = p.if_hall && !(irk == 2 && s->Ur != 0.0);
  const int cap = s->P > 1 ? s->ovl_z_warps : 0;
  const int na = ntasks - 4;                      // rows up to momentum z
  const int g0a = z.task[0].gout;                 // V1 slots of group A are g0a .. g0a + na - 1, of group B 4 .. 7
  auto z_rows = [&](int first, int n) -> int {   // RHS rows [first, first + n) on the exchange stream, then the flag barrier
    StreamScope sc(s, s->xstream, 1, cap);
    ZParams zz = z;
    for (int i = 0; i < n; ++i) zz.task[i] = z.task[first + i];
    if (s->tune_rhs) LAPS_TRY(rhs_z(s, zz, n));
    else LAPS_TRY(spec_z(s, zz, n, "spec_z"));
    return host_barrier(s);
  };
  LAPS_TRY(link_streams(s, s->stream, s->xstream));
  {  // every rank's forward y pass has landed in W2; every rank has finished reading V1 (previous stage's inverse y pass)
    StreamScope sc(s, s->xstream, 1, cap);
    LAPS_TRY(host_barrier(s));
  }
  LAPS_TRY(z_rows(0, na));
  LAPS_TRY(link_streams(s, s->xstream, s->stream));
  LAPS_TRY(z_rows(na, 4));                                   // beside ...
  LAPS_TRY(inverse_yx(s, g0a, na, true));                    // ... the inverse y, x passes of group A (main stream)
  LAPS_TRY(link_streams(s, s->xstream, s->stream));
  {  // J for the next stage's calc_flux + the continuity row (reads the rows just updated: same stream, after them)
    StreamScope sc(s, s->xstream, 1, cap);
    LAPS_TRY(launch_current_tasks(s, s->uB, true, want_j, s->mass_from_state ? irk : -1));
    if (want_j || s->mass_from_state) LAPS_TRY(host_barrier(s));
  }
  LAPS_TRY(inverse_yx(s, 4, 4, true));                       // group B beside the current / continuity tasks
  LAPS_TRY(link_streams(s, s->xstream, s->stream));
  if (s->mass_from_state) LAPS_TRY(inverse_yx(s, 0, 1, true));
  if (want_j) LAPS_TRY(inverse_yx(s, 8, 3, true));
  return stage_finish(s, want_j);
}

int stage(S* s, int irk) {
  const laps_params& p = s->p;
  if (s->incomp) return stage_incomp(s, irk);
  if (irk == 0 && s->front_ready) s->front_ready = false;   // laps_step has already run this stage's front half
  else LAPS_TRY(stage_front(s, false));
  if (!use_overlap(s)) LAPS_TRY(host_barrier(s));
  {  // z-pass + calc_rhs + rkt + dealias + inverse z
    ZParams z; fill_zparams(s, z, true);
    z.cc = s->cc1[irk]; z.dd = s->dd1[irk]; z.dt_irk = s->tstep[irk];
    z.read_rk = (irk > 0); z.write_rk = (irk < 2);
    const int* L = s->slot;   // flux index (0-based: F1..F18, expand_term) -> field slot
    const int X = p.if_AEB ? L[18] : -1;
    if (!s->two_d) {
      //                 v  g  fa     ca   fb     cb   fx cx   sg   fc     sc
      z.task[0] = rhs_task(0, 0, L[0], 1.0, L[1], 1.0, -1, 0.0, -1.0, L[2], -1.0);
      z.task[1] = rhs_task(1, 1, L[3], 1.0, L[4], 1.0, -1, 0.0, -1.0, L[5], -1.0);
      z.task[2] = rhs_task(2, 2, L[6], 1.0, L[7], 1.0, -1, 0.0, -1.0, L[8], -1.0);
      z.task[3] = rhs_task(3, 3, L[9], 1.0, L[10], 1.0, -1, 0.0, -1.0, L[11], -1.0);
      // dB/dt = curl E (mhdrhs.f90:223-228): fnl5 = kz F14 - ky F15 ; fnl6 = kx F15 - kz F13 ; fnl7 = ky F13 - kx F14
      z.task[4] = rhs_task(4, 4, -1, 0.0, L[14], 1.0, -1, 0.0, -1.0, L[13], +1.0);
      z.task[5] = rhs_task(5, 5, L[14], 1.0, -1, 0.0, -1, 0.0, +1.0, L[12], -1.0);
      z.task[6] = rhs_task(6, 6, L[13], -1.0, L[12], 1.0, -1, 0.0, +1.0, -1, 0.0);
      // energy: -(kx F16 + ky F17 + kz F18) + X  (mhdrhs.f90:231-233,250)
      z.task[7] = rhs_task(7, 7, L[15], 1.0, L[16], 1.0, X, -1.0, -1.0, L[17], -1.0);
    } else {
      // 2D tree (2D/mhdrhs.f90:290-317), kz = 0: the x derivative is applied before the line transform
      // (kx is constant along a line), the y derivative after it (the line axis carries ky).
      z.task[0] = rhs_task(0, 0, L[0], 1.0, -1, 0.0, -1, 0.0, -1.0, L[1], -1.0);
      z.task[1] = rhs_task(1, 1, L[3], 1.0, -1, 0.0, -1, 0.0, -1.0, L[4], -1.0);
      z.task[2] = rhs_task(2, 2, L[6], 1.0, -1, 0.0, -1, 0.0, -1.0, L[7], -1.0);
      z.task[3] = rhs_task(3, 3, L[9], 1.0, -1, 0.0, -1, 0.0, -1.0, L[10], -1.0);
      // fnl5 = -ky F15 ; fnl6 = kx F15 ; fnl7 = ky F13 - kx F14
      z.task[4] = rhs_task(4, 4, -1, 0.0, -1, 0.0, -1, 0.0, -1.0, L[14], -1.0);
      z.task[5] = rhs_task(5, 5, L[14], 1.0, -1, 0.0, -1, 0.0, +1.0, -1, 0.0);
      z.task[6] = rhs_task(6, 6, L[13], -1.0, -1, 0.0, -1, 0.0, +1.0, L[12], +1.0);
      z.task[7] = rhs_task(7, 7, L[15], 1.0, -1, 0.0, X, -1.0, -1.0, L[16], -1.0);
      if (z.corot2d) {
        // if_corotating (2D/mhdrhs.f90:282-288): kx_eff = kx cos + ky sin, ky_eff = (-kx sin + ky cos) R0/R.  The kx parts
        // are the column constants the kernel applies to fa, fb before the line transform; the ky parts share the factor
        // (i ky R0/R) the kernel applies after it:  ky sin F^a + ky cos (R0/R) F^b = (ky R0/R) [(sin R/R0) F^a + cos F^b].
        const double cq = s->cosa, sq = s->sina * s->radius / p.radius0;
        auto div_row = [&](int v, int fxa, int fya, int X_, double cxv) {   // -(kx_eff F^a + ky_eff F^b) (+ cx X^)
          ZTask t = rhs_task(v, v, fxa, 1.0, fya, 1.0, X_, cxv, -1.0, fya, -1.0);
          t.cf1 = cq; t.fc2 = fxa; t.cf2 = sq;
          return t;
        };
        z.task[0] = div_row(0, L[0], L[1], -1, 0.0);
        z.task[1] = div_row(1, L[3], L[4], -1, 0.0);
        z.task[2] = div_row(2, L[6], L[7], -1, 0.0);
        z.task[3] = div_row(3, L[9], L[10], -1, 0.0);
        z.task[7] = div_row(7, L[15], L[16], X, -1.0);
        // fnl5 = -ky_eff F15 ; fnl6 = kx_eff F15 ; fnl7 = ky_eff F13 - kx_eff F14
        z.task[4] = rhs_task(4, 4, -1, 0.0, L[14], 1.0, -1, 0.0, -1.0, L[14], -1.0);  z.task[4].cf1 = cq;
        z.task[5] = rhs_task(5, 5, L[14], 1.0, -1, 0.0, -1, 0.0, +1.0, L[14], +1.0);  z.task[5].cf1 = sq;
        z.task[6] = rhs_task(6, 6, L[13], -1.0, L[12], 1.0, -1, 0.0, +1.0, L[12], +1.0);
        z.task[6].cf1 = cq; z.task[6].fc2 = L[13]; z.task[6].cf2 = -sq;
      }
      if (s->ext_slot >= 0) { z.task[6].fx = s->ext_slot; z.task[6].cx = 1.0; }   // fnl(7) += external_force_fourier(1), 2D/mhdrhs.f90:370-372
      if (p.if_AEB && p.if_z_radial) {   // 2D/mhdrhs.f90:324-343
        static const double c2[8] = {2.0, 3.0, 3.0, 2.0, 1.0, 1.0, 2.0, 0.0};
        for (int v = 0; v < 8; ++v) z.task[v].aeb_c = c2[v];
      }
    }
    const int t0 = s->mass_from_state ? 1 : 0;   // the continuity row is a kZMass task of the launch below
    if (t0) for (int v = 1; v < 8; ++v) z.task[v - 1] = z.task[v];
    if (use_overlap(s) == 1) return stage_back_direct(s, irk, z, 8 - t0);
    if (use_overlap(s) == 2) return stage_back_push(s, irk, z, 8 - t0);
    // (the persistent kernel carries one field in its (i k_line) term: the 2D tree with if_corotating goes through k_spec_z)
    if (s->tune_rhs && !z.corot2d) LAPS_TRY(rhs_z(s, z, 8 - t0));
    else LAPS_TRY(spec_z(s, z, 8 - t0, "spec_z"));
  }
  // J for the next stage's calc_flux.  After the last stage of a step in the expanding box the
  // driver moves the radius (evolve_radius, mhd.f90:248) and with it the wave vectors J is built
  // from (mhdrhs.f90:313-326), so that J would be discarded: leave it to refresh_current.
  const bool want_j = p.if_hall && !(irk == 2 && s->Ur != 0.0);
  LAPS_TRY(launch_current_tasks(s, s->uB, true, want_j, s->mass_from_state ? irk : -1));
  LAPS_TRY(host_barrier(s));
  LAPS_TRY(inverse_yx(s, 0, want_j ? 11 : 8, true));
  return stage_finish(s, want_j);
}

// reduce_launch enqueues the final reduction (+ the inter-rank allreduce) and the copy of the scalars to the host,
// reduce_wait blocks until that copy has landed (later launches in the stream keep the GPU busy meanwhile).
int reduce_launch(S* s, int nrows, int op /*0 sum,1 min,2 max*/, double init) {
  LaunchScope ls(s, "reduce");
  if (op == 0) LAPS_LAUNCH((k_reduce_final<OpSum>), dim3((unsigned)nrows), dim3(256), 0, s->stream, s->d_partial, s->nblk, s->d_scal, init);
  if (op == 1) LAPS_LAUNCH((k_reduce_final<OpMin>), dim3((unsigned)nrows), dim3(256), 0, s->stream, s->d_partial, s->nblk, s->d_scal, init);
  if (op == 2) LAPS_LAUNCH((k_reduce_final<OpMax>), dim3((unsigned)nrows), dim3(256), 0, s->stream, s->d_partial, s->nblk, s->d_scal, init);
  LAPS_TRY(check_launch(s, "k_reduce_final"));
  if (s->P > 1) {  // mpi_allreduce (mhd.f90:419,567; mhdrms.f90:96,98,122), combined in rank order
    if (!s->wired) { s->err = "nranks > 1 but the ranks are not connected"; return 1; }
    ++s->epoch[0];
    LAPS_LAUNCH(k_xchg_allreduce, dim3(1), dim3(32), 0, s->stream, s->xp, 0, s->epoch[0], s->d_scal, nrows, op);
    LAPS_TRY(check_launch(s, "k_xchg_allreduce"));
  }
  LAPS_CK(s, cudaMemcpyAsync(s->h_scal, s->d_scal, nrows * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  LAPS_CK(s, cudaEventRecord(s->ev_scal, s->stream));
  return 0;
}
int reduce_wait(S* s) {
  LAPS_CK(s, cudaEventSynchronize(s->ev_scal));
  return check_abort(s);
}
int reduce_final(S* s, int nrows, int op, double init) {
  LAPS_TRY(reduce_launch(s, nrows, op, init));
  return reduce_wait(s);
}

int require_state(S* s) {
  if (!s->have_state) { s->err = "no state: call laps_set_primitive first"; return 1; }
  return 0;
}

}  // namespace

// =================================================================================================
extern "C" {

const char* laps_last_error(laps_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int laps_create(const laps_params* params, laps_handle* out) {
  if (!params || !out) { g_create_error = "null argument"; return 1; }
  *out = nullptr;
  const laps_params& u = *params;
  if (u.abi_version != LAPS_ABI_VERSION) { g_create_error = "laps_params.abi_version mismatch"; return 1; }
  const bool two_d = u.ndim == 2;
  if (u.ndim != 0 && u.ndim != 2 && u.ndim != 3) { g_create_error = "ndim must be 2 or 3"; return 1; }
  laps_params p = u;   // internal view
  if (two_d) {
    // 2D tree (src_compressible/2D/): the (nx, ny, 1) grid is held as (nx, 1, ny) so that the
    // reference's y lines are the contiguous lines of the fused spectral pass; the reference's
    // ky tables become the internal z tables (Ly -> Lz, afy -> afz).
    if (u.nz != 1) { g_create_error = "ndim = 2 needs nz = 1"; return 1; }
    if (u.nranks != 1) { g_create_error = "the 2D tree runs on one GPU (nranks = 1)"; return 1; }
    if (u.if_AEB && u.if_corotating && u.if_z_radial) { g_create_error = "if_z_radial and if_corotating exclude each other (2D/mhd.f90:62-67)"; return 1; }
    if (!size_supported(u.nx) || !(size_supported(u.ny) || u.ny == 8)) { g_create_error = "nx, ny must be 2^k, 3 * 2^k or 5 * 2^k in [16, 2048] (ny = 8 too)"; return 1; }
    p.ny = 1; p.nz = u.ny; p.Ly = 1.0; p.Lz = u.Ly; p.afz = u.afy;
    if (u.dealias_option < 0 || u.dealias_option > 3) { g_create_error = "dealias_option must be 0..3 in the 2D tree"; return 1; }
  } else {
    if (!size_supported(p.nx) || !size_supported(p.ny) || !(size_supported(p.nz) || p.nz == 8)) {
      g_create_error = "nx, ny, nz must be 2^k, 3 * 2^k or 5 * 2^k in [16, 2048] (nz = 8 too)"; return 1;
    }
    if (p.dealias_option < 0 || p.dealias_option > 2) { g_create_error = "dealias_option must be 0, 1 or 2"; return 1; }
    p.if_z_radial = 0; p.if_limit_dt_increase = 0;
  }
  if (u.if_external_force && (!two_d || u.incompressible)) {
    g_create_error = "if_external_force exists only in the 2D compressible tree (src_compressible/2D/mhd.f90:43)"; return 1;
  }
  if (u.incompressible) {
    if (two_d && u.if_z_radial) { g_create_error = "if_z_radial does not exist in src_incompressible/2D"; return 1; }
    if (!(u.rho0 > 0.0)) { g_create_error = "incompressible: rho0 must be positive (mhdinit.f90:15)"; return 1; }
  }
  if (p.nranks < 1 || p.nranks > LAPS_MAX_RANKS || p.rank < 0 || p.rank >= p.nranks) {
    g_create_error = "bad rank/nranks (1..8 ranks, slab decomposition)"; return 1;
  }
  if (p.nz / p.nranks < 1 || (!two_d && p.ny / p.nranks < 1)) { g_create_error = "more ranks than planes"; return 1; }
  S* s = new S();
  s->p = p;
  s->user = u;
  s->two_d = two_d;
  s->incomp = u.incompressible != 0;
  s->rho0 = u.rho0; s->p0 = 1.0;
  auto fail = [&](const std::string& m) { g_create_error = m; laps_destroy(s); return 1; };
#ifndef LAPS_EMU_BUILD
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("no CUDA device available (this library has no CPU path)");
  if (p.device < 0 || p.device >= ndev) return fail("bad device ordinal");
  if (cudaSetDevice(p.device) != cudaSuccess) return fail("cudaSetDevice failed");
  if (cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, p.device) != cudaSuccess) return fail("cudaDeviceGetAttribute failed");
#endif
  s->nx = p.nx; s->ny = p.ny; s->nz = p.nz; s->nxh = p.nx / 2 + 1; s->P = p.nranks; s->rank = p.rank;
  decompose_1d(s->nz, s->P, s->zoffs, s->zlens);   // zj_offset/zj_size (parallel.f90:102)
  decompose_1d(s->ny, s->P, s->yoffs, s->ylens);   // yj_offset/yj_size (parallel.f90:101)
  {  // Fourier-row ownership.  Default: the reference's contiguous ky slabs up to 3 ranks; from 4 ranks on, when a
     // dealiasing mask (options 1, 3) removes the middle of the ky axis and those rows are skipped, the rows are dealt
     // round-robin so that every rank owns the same share of the surviving ones (with slabs the ranks in the middle of
     // the spectrum idle: at 8 ranks two of them own no surviving row).  LAPS_TUNE_CYCLIC=0/1 forces either.
    int prune = 1;
    if (const char* e = std::getenv("LAPS_TUNE_PRUNE")) prune = std::atoi(e);
    s->cyclic = s->P >= 4 && !two_d && prune != 0 && (p.dealias_option == 1 || p.dealias_option == 3);
    if (const char* e = std::getenv("LAPS_TUNE_CYCLIC")) s->cyclic = std::atoi(e) != 0 && s->P > 1 && !two_d;
  }
  if (s->cyclic) {   // row ky belongs to rank ky % P; rank q holds rows q, q + P, q + 2P, ...
    for (int q = 0; q < s->P; ++q) { s->yoffs[q] = q; s->ylens[q] = (s->ny - q + s->P - 1) / s->P; }
    s->ystride = s->P;
  }
  s->nzl = s->zlens[s->rank]; s->zo = s->zoffs[s->rank];
  s->nyl = s->ylens[s->rank]; s->yo = s->yoffs[s->rank];
  s->npts = (size_t)s->nx * s->ny * s->nzl;
  s->ncol = (size_t)s->nxh * s->nyl;
  s->csz = s->ncol * s->nz;
  s->w1sz = (size_t)s->nxh * s->nzl * s->ny;
  s->xz = two_d ? 1 : s->nzl; s->xy = two_d ? s->nzl : s->ny;
  s->mass_from_state = !s->incomp && p.dealias_option != 0;
  if (const char* e = std::getenv("LAPS_TUNE_MASS")) s->mass_from_state = !s->incomp && std::atoi(e) != 0;
  if (const char* e = std::getenv("LAPS_TUNE_SYM")) s->sym_tensor = std::atoi(e) != 0;
  {  // field slots of the fluxes; the 2D tree never uses the z fluxes F3,F6,F9,F12,F18 (kz = 0)
    int n = 0;
    for (int j = 0; j < 19; ++j) {
      const bool zflux = (j == 2 || j == 5 || j == 8 || j == 11 || j == 17);
      const bool on = j == 18 ? (p.if_AEB != 0) : !(two_d && zflux) && !(s->mass_from_state && j < 3);
      s->slot[j] = on ? n++ : -1;
      s->fslot[j] = s->slot[j];
      const int twin = j == 6 ? 4 : (j == 9 ? 5 : (j == 10 ? 8 : -1));   // (row, dir) -> (dir, row)
      if (on && s->sym_tensor && twin >= 0 && s->slot[twin] >= 0) { s->slot[j] = s->slot[twin]; s->fslot[j] = -1; --n; }
    }
    if (two_d && !u.incompressible && u.if_external_force) s->ext_slot = n++;   // external_force_fourier(:,:,:,1)
    s->nf = n;
  }
  s->ni = 8 + (p.if_hall ? 3 : 0);
  if (s->incomp) { s->nf = 6; s->ni = 12; }   // Fp (3) + E (3) forward; J (3) + grad u (9) / the state (8) inverse
  {  // the mask test "sqrt(s) > 1./3." (dealiasing.f90:94) as a threshold on s: sqrt is correctly rounded
     // and monotonic, so { s : sqrt(s) > c } = { s >= T } with T the smallest double that passes
    const double c = 1.0 / 3.0;
    double t = c * c;
    while (std::sqrt(t) > c) t = std::nextafter(t, 0.0);
    while (!(std::sqrt(t) > c)) t = std::nextafter(t, 1.0);
    s->da_thresh = t;
  }
  if (const char* e = std::getenv("LAPS_TUNE_CGZ")) s->tune_cgz = std::atoi(e);
  if (const char* e = std::getenv("LAPS_TUNE_RHS")) s->tune_rhs = std::atoi(e);
  if (const char* e = std::getenv("LAPS_TUNE_RCG")) s->tune_rcg = std::atoi(e);
  if (const char* e = std::getenv("LAPS_TUNE_Z")) s->tune_z = std::atoi(e);
  if (const char* e = std::getenv("LAPS_TUNE_FUSEX")) s->tune_fusex = std::atoi(e);
  if (const char* e = std::getenv("LAPS_TUNE_ZCHUNK")) s->tune_zchunk = std::atoi(e);
  if (const char* e = std::getenv("LAPS_TUNE_SPEC")) s->tune_spec = std::atoi(e);
  // The reference re-derives uu_fourier from the real fields at the start of every stage
  // (src_incompressible/mhd.f90:325).  For a spectrum the dealiasing has band-limited (options 1, 2: the
  // Nyquist planes are removed) that round trip is the identity up to round-off and is skipped; with
  // dealias_option 0 it projects out the non-Hermitian Nyquist content the derivatives create, so it is done.
  s->retransform = s->incomp && p.dealias_option == 0;
  if (const char* e = std::getenv("LAPS_TUNE_RETRANSFORM")) s->retransform = s->incomp && std::atoi(e) != 0;
  s->nblk = (int)std::min<size_t>(148 * 8, (s->npts + 255) / 256);   // grid-stride loops: 8 CTAs per SM at most

  // expanding box: mhd.f90:88-91, AEBmod.f90:16-31
  s->Ur = p.if_AEB ? p.Ur0 : 0.0;
  s->radius = p.radius0;
  aeb_calc(s);
  const double ang = p.if_corotating ? p.corotating_angle : 0.0;
  s->cosa = std::cos(ang); s->sina = std::sin(ang);
  s->wnx = wave_numbers(s->nx, p.Lx); s->wny = wave_numbers(s->ny, p.Ly); s->wnz = wave_numbers(s->nz, p.Lz);

  if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) return fail("cudaStreamCreate failed");
  s->cs = s->stream;
#ifndef LAPS_EMU_BUILD
  {  // exchange stream: highest priority, so that its (grid-capped) passes get their share of every SM as CTAs retire
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&s->xstream, cudaStreamNonBlocking, hi) != cudaSuccess) return fail("cudaStreamCreateWithPriority failed");
    for (int i = 0; i < 16; ++i)
      if (cudaEventCreateWithFlags(&s->ev_link[i], cudaEventDisableTiming) != cudaSuccess) return fail("cudaEventCreate failed");
    for (int i = 0; i < 4; ++i)
      if (cudaEventCreateWithFlags(&s->ev_push[i], cudaEventDisableTiming) != cudaSuccess) return fail("cudaEventCreate failed");
    for (int i = 0; i < 3; ++i)
      if (cudaEventCreateWithFlags(&s->ev_zrow[i], cudaEventDisableTiming) != cudaSuccess) return fail("cudaEventCreate failed");
  }
#else
  s->xstream = (cudaStream_t)1;   // the emulator runs launches synchronously: the two-stream schedule is exercised as a sequence
#endif
  if (const char* e = std::getenv("LAPS_TUNE_OVERLAP")) s->tune_overlap = std::atoi(e);
  if (const char* e = std::getenv("LAPS_TUNE_SCREEN")) s->tune_screen = std::atoi(e);
  if (const char* e = std::getenv("LAPS_TUNE_TLY")) s->tune_tly = std::atoi(e);
  if (const char* e = std::getenv("LAPS_TUNE_OVL_PUSH")) s->ovl_push_ctas = std::atoi(e);
  if (const char* e = std::getenv("LAPS_TUNE_OVL_CHUNKS")) s->ovl_chunks = std::atoi(e);
  cudaEventCreate(&s->ev0); cudaEventCreate(&s->ev1); cudaEventCreate(&s->ev_scal);

  const int nmax = s->nf > s->ni ? s->nf : s->ni;
  // largest per-rank slab sizes of the exchanged layouts (the last rank holds the remainder)
  s->bytesX = std::max((size_t)s->nf * s->npts * sizeof(double), (size_t)s->ni * s->w1sz * sizeof(cplx));
  s->bytesY = (size_t)nmax * s->w1sz * sizeof(cplx);
  s->bytesZ = (size_t)std::max(s->nf, 8) * s->csz * sizeof(cplx);   // the 8 state fields pass through W2 in laps_set_primitive
  bool ok = true;
  auto alloc = [&](void** ptr, size_t bytes) { if (ok && cudaMalloc(ptr, bytes) != cudaSuccess) ok = false; else if (ok) s->dev_bytes += bytes; };
  alloc((void**)&s->uu, 8 * s->npts * sizeof(double));
  if (p.if_hall || s->incomp) alloc((void**)&s->J, 3 * s->npts * sizeof(double));
  if (s->incomp) alloc((void**)&s->G, 9 * s->npts * sizeof(double));
  if (s->ext_slot >= 0) alloc((void**)&s->ext, s->npts * sizeof(double));
  alloc(&s->bufX, s->bytesX); alloc(&s->bufY, s->bytesY); alloc(&s->bufZ, s->bytesZ);
  {  // staging blocks of the two-stream schedule (3D compressible tree on several ranks, or LAPS_TUNE_OVERLAP=1)
    int want = 0;
    if (const char* e = std::getenv("LAPS_TUNE_OVERLAP")) want = std::atoi(e) == 2 ? 1 : want;
    if (const char* e = std::getenv("LAPS_TUNE_STAGING")) want = std::atoi(e) != 0 ? 1 : want;   // (A/B on a live handle: laps_set_tune)
    if (want && !two_d && !s->incomp) {
      s->bytesT = (size_t)s->nf * s->w1sz * sizeof(cplx);
      alloc(&s->bufT, s->bytesT);
    }
  }
  alloc((void**)&s->uA, 8 * s->csz * sizeof(cplx));
  alloc((void**)&s->uB, 8 * s->csz * sizeof(cplx));
  alloc((void**)&s->rk, 8 * s->csz * sizeof(cplx));
  alloc((void**)&s->tw_x, twiddle_len(s->nx) * sizeof(cplx)); alloc((void**)&s->tw_y, twiddle_len(s->ny) * sizeof(cplx));
  alloc((void**)&s->tw_z, twiddle_len(s->nz) * sizeof(cplx));
  alloc((void**)&s->d_tab, ((size_t)3 * (s->nxh + s->ny + s->nz) + s->nz) * sizeof(double));
  alloc((void**)&s->d_partial, (size_t)32 * s->nblk * sizeof(double));
  alloc((void**)&s->d_scal, 64 * sizeof(double));
  alloc((void**)&s->xblk, sizeof(XchgBlock));
  if (!ok) return fail("device allocation failed (state + work buffers need about " +
                       std::to_string((s->bytesX + s->bytesY + s->bytesZ + 24 * s->csz * 16 + 11 * s->npts * 8) >> 20) + " MiB)");
  if (cudaMallocHost((void**)&s->h_scal, 64 * sizeof(double)) != cudaSuccess) return fail("cudaMallocHost failed");
  if (cudaMallocHost((void**)&s->h_abort, 64) != cudaSuccess) return fail("cudaMallocHost failed");
  std::memset(s->h_abort, 0, 64);
  if (s->ext && cudaMemsetAsync(s->ext, 0, s->npts * sizeof(double), s->stream) != cudaSuccess) return fail("cudaMemsetAsync failed");
  {
    double* t = s->d_tab;
    s->kxr = t; s->kyr = s->kxr + s->nxh; s->kze = s->kyr + s->ny;
    s->ksq_x = s->kze + s->nz; s->ksq_y = s->ksq_x + s->nxh; s->ksq_z = s->ksq_y + s->ny;
    s->dax = s->ksq_z + s->nz; s->day = s->dax + s->nxh; s->daz = s->day + s->ny;
    s->kzr = s->daz + s->nz;
  }
  for (int a = 0; a < 3; ++a) {
    const int n = a == 0 ? s->nx : (a == 1 ? s->ny : s->nz);
    cplx* d = a == 0 ? s->tw_x : (a == 1 ? s->tw_y : s->tw_z);
    std::vector<cplx> t = twiddle_table(n);
    if (cudaMemcpy(d, t.data(), t.size() * sizeof(cplx), cudaMemcpyHostToDevice) != cudaSuccess) return fail("twiddle upload failed");
  }
  if (upload_tables(s)) return fail(s->err);
  {  // columns the dealiasing mask removes entirely (see laps_solver::nkx)
    if (const char* e = std::getenv("LAPS_TUNE_PRUNE")) s->tune_prune = std::atoi(e);
    const double* dax = s->h_tab.data() + 2 * (s->nxh + s->ny + s->nz);
    const double* day = dax + s->nxh;
    s->nkx = s->nxh; s->kymax = s->ny / 2;
    if (s->tune_prune && (p.dealias_option == 1 || p.dealias_option == 3)) {
      // option 1: the test is fl(fl(tx+ty)+tz) >= T with non-negative terms, and floating-point addition
      // is monotonic, so tx >= T (or ty >= T) alone already removes the mode.  option 3: per-axis flags.
      auto gone = [&](double t) { return p.dealias_option == 1 ? t >= s->da_thresh : t != 0.0; };
      int nk = 0;
      while (nk < s->nxh && !gone(dax[nk])) ++nk;
      bool tail_gone = true;
      for (int i = nk; i < s->nxh; ++i) tail_gone = tail_gone && gone(dax[i]);
      if (tail_gone && nk >= 1) s->nkx = nk;
      if (!s->two_d) {
        int km = 0;
        while (km + 1 <= s->ny / 2 && !gone(day[km + 1])) ++km;
        bool mid_gone = true;
        for (int j = km + 1; j < s->ny - km; ++j) mid_gone = mid_gone && gone(day[j]);
        if (mid_gone) s->kymax = km;
      }
    }
    // this rank's surviving ky rows: run A = owned rows with ky <= kymax, run B = owned rows with ky >= ny - kymax
    // (local rows are in increasing ky for slabs and for cyclic ownership alike)
    auto ky_of = [&](int kyl) { return s->yo + kyl * s->ystride; };
    int nA = 0;
    while (nA < s->nyl && ky_of(nA) <= s->kymax) ++nA;
    int b0 = nA;
    while (b0 < s->nyl && ky_of(b0) < s->ny - s->kymax) ++b0;
    s->pr_a0 = 0; s->pr_nA = nA; s->pr_b0 = b0;
    s->pr_nkyl = nA + (s->nyl - b0);
    if (s->kymax >= s->ny / 2) { s->pr_nkyl = s->nyl; s->pr_nA = s->nyl; s->pr_a0 = 0; s->pr_b0 = 0; }
    s->pr_ncol = s->nkx * s->pr_nkyl;
    const double* daz = day + s->ny;
    const bool masked = s->tune_prune && (p.dealias_option == 1 || p.dealias_option == 3);
    int tune_circle = 1, tune_kz = 1;
    if (const char* e = std::getenv("LAPS_TUNE_CIRCLE")) tune_circle = std::atoi(e);
    if (const char* e = std::getenv("LAPS_TUNE_KZPRUNE")) tune_kz = std::atoi(e);
    s->kzprune = masked && tune_kz != 0;
    if (masked && p.dealias_option == 1 && tune_circle) {
      // A column (kx, ky) is dead for every kz iff it is dead at kz = 0: the test fl(fl(tx + ty) + tz) >= T is
      // monotonic in tz >= 0.  ty grows with |ky| (each operation of dealiasing.f90:92 is monotonic), so the
      // surviving rows of a kx column are |ky| <= kymax_x[kx].
      std::vector<int> kym(s->nxh, -1), cmap;
      for (int kx = 0; kx < s->nkx; ++kx) {
        int km = -1;
        while (km + 1 <= s->kymax && !(dax[kx] + day[km + 1] >= s->da_thresh)) ++km;
        kym[kx] = km;
        for (int r = 0; r < s->pr_nkyl; ++r) {
          const int kyl = r < s->pr_nA ? s->pr_a0 + r : s->pr_b0 + r - s->pr_nA;
          const int ky = ky_of(kyl);
          if (!(dax[kx] + day[ky] >= s->da_thresh)) cmap.push_back(kx * s->nyl + kyl);
        }
      }
      s->pr_ncol = (int)cmap.size();
      alloc((void**)&s->d_kymax_x, s->nxh * sizeof(int));
      alloc((void**)&s->d_colmap, std::max<size_t>(1, cmap.size()) * sizeof(int));
      alloc((void**)&s->d_colkx, std::max<size_t>(1, cmap.size()) * sizeof(int));
      if (!ok) return fail("device allocation failed (pruning tables)");
      std::vector<int> ckx(cmap.size());
      for (size_t i = 0; i < cmap.size(); ++i) ckx[i] = cmap[i] / s->nyl;
      if (cudaMemcpy(s->d_kymax_x, kym.data(), s->nxh * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
          (!cmap.empty() && (cudaMemcpy(s->d_colmap, cmap.data(), cmap.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
                             cudaMemcpy(s->d_colkx, ckx.data(), ckx.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess)))
        return fail("pruning table upload failed");
    }
    {  // surviving (kx, ky) columns of all ranks
      long long n = 0;
      for (int kx = 0; kx < s->nkx; ++kx)
        for (int ky = 0; ky < s->ny; ++ky) {
          if (ky > s->kymax && ky < s->ny - s->kymax) continue;
          if (masked && p.dealias_option == 1 && tune_circle && dax[kx] + day[ky] >= s->da_thresh) continue;
          ++n;
        }
      s->pr_cols_all = n;
    }
    {  // surviving modes of this rank (traffic model)
      long long m = 0;
      for (int kx = 0; kx < s->nxh; ++kx)
        for (int kyl = 0; kyl < s->nyl; ++kyl) {
          const double dxy = (p.dealias_option == 1 || p.dealias_option == 3) ? dax[kx] + day[ky_of(kyl)] : 0.0;
          for (int kz = 0; kz < s->nz; ++kz) {
            bool dead = false;
            if (masked && p.dealias_option == 1) dead = dxy + daz[kz] >= s->da_thresh;
            if (masked && p.dealias_option == 3) dead = dxy != 0.0 || daz[kz] != 0.0;
            m += dead ? 0 : 1;
          }
        }
      s->pr_modes = m;
    }
  }
  cudaMemsetAsync(s->rk, 0, 8 * s->csz * sizeof(cplx), s->stream);
#ifndef LAPS_EMU_BUILD
  if (s->tune_zchunk > 0 && !two_d && !s->incomp) {
    // keep the flux chunk resident: persisting L2 lines for the chunk buffer, streaming for everything else
    int persist = 1;
    if (const char* e = std::getenv("LAPS_TUNE_L2PERSIST")) persist = std::atoi(e);
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, p.device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, p.device);
    const size_t want = (size_t)s->nf * std::min(s->tune_zchunk, s->nzl) * s->nx * s->ny * sizeof(double);
    if (persist && max_persist > 0 && max_window > 0) {
      const size_t carve = std::min(want, (size_t)max_persist);
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
      cudaStreamAttrValue a;
      std::memset(&a, 0, sizeof(a));
      a.accessPolicyWindow.base_ptr = s->bufX;
      a.accessPolicyWindow.num_bytes = std::min(want, (size_t)max_window);
      a.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)a.accessPolicyWindow.num_bytes);
      a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      cudaStreamSetAttribute(s->stream, cudaStreamAttributeAccessPolicyWindow, &a);
      (void)cudaGetLastError();
    }
  }
#endif

  // single rank: the exchange tables point at this rank's own buffers
  std::memset(&s->tabW2, 0, sizeof(PeerTable)); std::memset(&s->tabV1, 0, sizeof(PeerTable));
  s->tabW2.nparts = s->tabV1.nparts = s->P;
  { int sh = 0; while ((1 << sh) < s->P) ++sh; s->tabW2.shift = s->tabV1.shift = sh; }
  s->tabW2.quot = s->ny / s->P; s->tabV1.quot = s->nz / s->P;
  for (int q = 0; q < s->P; ++q) {
    s->tabW2.off[q] = s->yoffs[q]; s->tabW2.len[q] = s->ylens[q];
    s->tabW2.cyclic = s->cyclic ? 1 : 0;
    s->tabV1.off[q] = s->zoffs[q]; s->tabV1.len[q] = s->zlens[q];
  }
  s->tabW2.base[s->rank] = buf_W2(s);
  s->tabV1.base[s->rank] = buf_V1(s);
  s->tabT = s->tabW2;     // same ownership tables; the blocks lie one behind the other in this rank's staging buffer
  {
    size_t at = 0;
    for (int q = 0; q < s->P; ++q) {
      s->tabT.base[q] = s->bufT ? (cplx*)s->bufT + at : nullptr;
      at += (size_t)s->nf * s->nxh * s->ylens[q] * s->nzl;
      // rows of peer q inside the rectangle |ky| <= kymax (run A: ky <= kymax, run B: ky >= ny - kymax)
      int nA = 0;
      while (nA < s->ylens[q] && s->yoffs[q] + nA * s->ystride <= s->kymax) ++nA;
      int b0 = nA;
      while (b0 < s->ylens[q] && s->yoffs[q] + b0 * s->ystride < s->ny - s->kymax) ++b0;
      if (s->kymax >= s->ny / 2) { nA = s->ylens[q]; b0 = s->ylens[q]; }
      s->peer_nA[q] = nA; s->peer_b0[q] = b0;
    }
  }
  std::memset(&s->xp, 0, sizeof(s->xp));
  std::memset(s->ipc_opened, 0, sizeof(s->ipc_opened));
  s->xp.rank = s->rank; s->xp.nranks = s->P;
  s->xp.blk[s->rank] = s->xblk;
  {  // budget of one inter-rank wait; a rank that exceeds it aborts the exchange on every rank (exchange.cuh)
    double sec = 120.0;
    if (const char* e = std::getenv("LAPS_XCHG_TIMEOUT_S")) sec = std::atof(e);
    if (!(sec > 0.0)) sec = 120.0;
    s->xp.timeout_ns = (unsigned long long)(sec * 1e9);
    s->xp.host_abort = s->h_abort;
  }
  if (cudaMemset(s->xblk, 0, sizeof(XchgBlock)) != cudaSuccess) return fail("cudaMemset failed");
  if (cudaStreamSynchronize(s->stream) != cudaSuccess) return fail("stream sync failed");
  *out = s;
  return 0;
}

int laps_destroy(laps_handle s) {
  if (!s) return 0;
  DeviceGuard guard_(s->p.device);
  if (s->stream) cudaStreamSynchronize(s->stream);   // bounded: the inter-rank waits time out (exchange.cuh)
#ifndef LAPS_EMU_BUILD
  if (s->xstream) cudaStreamSynchronize(s->xstream);
  if (s->ostream) cudaStreamSynchronize(s->ostream);
#endif
  for (int q = 0; q < LAPS_MAX_RANKS; ++q)
    for (int j = 0; j < 3; ++j)
      if (s->ipc_opened[q][j]) cudaIpcCloseMemHandle(s->ipc_opened[q][j]);
  cudaFree(s->xblk);
  cudaFree(s->uu); cudaFree(s->J); cudaFree(s->G); cudaFree(s->ext); cudaFree(s->bufX); cudaFree(s->bufY); cudaFree(s->bufZ); cudaFree(s->bufT);
  cudaFree(s->uA); cudaFree(s->uB); cudaFree(s->rk); cudaFree(s->tw_x); cudaFree(s->tw_y); cudaFree(s->tw_z);
  cudaFree(s->d_tab); cudaFree(s->d_partial); cudaFree(s->d_scal); cudaFree(s->d_kymax_x); cudaFree(s->d_colmap); cudaFree(s->d_colkx);
  if (s->h_scal) cudaFreeHost(s->h_scal);
  if (s->h_abort) cudaFreeHost(s->h_abort);
  for (auto& pe : s->prof) { cudaEventDestroy(pe.e0); cudaEventDestroy(pe.e1); }
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  if (s->ev_scal) cudaEventDestroy(s->ev_scal);
  cudaFree(s->snap);
  if (s->ev_snap) cudaEventDestroy(s->ev_snap);
  if (s->ev_out) cudaEventDestroy(s->ev_out);
  if (s->stream) cudaStreamDestroy(s->stream);
#ifndef LAPS_EMU_BUILD
  if (s->ostream) cudaStreamDestroy(s->ostream);
  if (s->xstream) cudaStreamDestroy(s->xstream);
  for (int i = 0; i < 16; ++i) if (s->ev_link[i]) cudaEventDestroy(s->ev_link[i]);
  for (int i = 0; i < 4; ++i) if (s->ev_push[i]) cudaEventDestroy(s->ev_push[i]);
  for (int i = 0; i < 3; ++i) if (s->ev_zrow[i]) cudaEventDestroy(s->ev_zrow[i]);
#endif
  delete s;
  return 0;
}

int laps_get_extents(laps_handle s, laps_extents* e) {
  if (!s || !e) return 1;
  e->nx = s->nx; e->ny = s->ny; e->nz = s->nz; e->nxh = s->nxh;
  e->z_offset = s->zo; e->z_size = s->nzl; e->y_offset = s->yo; e->y_size = s->nyl; e->y_stride = s->ystride;
  if (s->two_d) {  // the driver's view: uu(1:nx, 1:ny, 1, 1:8); spectral block (kx, all ky)
    e->ny = s->nz; e->nz = 1; e->z_offset = 0; e->z_size = 1; e->y_offset = 0; e->y_size = s->nz; e->y_stride = 1;
  }
  return 0;
}

static int sync_body(laps_handle s) {
  LAPS_CK(s, cudaStreamSynchronize(s->stream));
  return check_abort(s);
}

int laps_get_stream(laps_handle s, void** stream_out) {
  if (!s || !stream_out) return 1;
  *stream_out = (void*)s->stream;
  return 0;
}

static int finish_set_primitive(laps_handle s);

static int set_primitive_body(laps_handle s, const double* uu_local) {
  if (!s || !uu_local) return 1;
  s->front_ready = false;
  LAPS_TRY(settle_exchange(s));
  LAPS_CK(s, cudaMemcpyAsync(s->uu, uu_local, 8 * s->npts * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  return finish_set_primitive(s);
}

// initial_calc_conserve_variable + transform_uu_real_to_fourier on the primitive fields already in s->uu
static int finish_set_primitive(laps_handle s) {
  {
    LaunchScope ls(s, "prim_to_cons", (8 + 4) * bytes_real(s));
    LAPS_LAUNCH(k_prim_to_cons, dim3((unsigned)s->nblk), dim3(256), 0, s->stream, s->uu, s->npts, s->p.adiabatic_index, s->incomp ? 1 : 0);
    LAPS_TRY(check_launch(s, "k_prim_to_cons"));
  }
  LAPS_CK(s, cudaMemsetAsync(s->uB, 0, 8 * s->csz * sizeof(cplx), s->stream));   // masked columns of the first output
  s->spectrum_full = true;
  LAPS_TRY(spectrum_from_real(s, false));
  s->have_state = true;
  s->j_stale = true;
  LAPS_CK(s, cudaStreamSynchronize(s->stream));  // the caller may reuse uu_local
  return check_abort(s);
}

// Initial data given as a mode table instead of a host array: the reference's ipert = 6/7
// hooks sum cosines point by point, O(modes x N^3) (mhdinit.f90:487-829); the same field is a sparse spectrum
// and one inverse transform.  field_v(x) = background[v] + sum_m Re( coef[v][m] exp(i k_m . x) ), v = rho, u, B;
// p = background[7].  k = integer wave vectors (ikx >= 0), coef = complex128 pairs [7][nmodes].
static int set_primitive_modes_body(laps_handle s, int32_t nmodes, const int32_t* k, const double* coef, const double* background) {
  if (!s || nmodes < 0 || (nmodes > 0 && (!k || !coef)) || !background) return 1;
  s->front_ready = false;
  LAPS_TRY(settle_exchange(s));
  const int NYr = s->two_d ? s->nz : s->ny;       // the driver's ny
  std::vector<long long> idx;
  std::vector<cplx> val;                            // [8][nent]
  std::vector<std::vector<cplx>> rows(8);
  auto push = [&](long long i, const cplx* c8) {
    for (size_t e = 0; e < idx.size(); ++e)
      if (idx[e] == i) { for (int v = 0; v < 8; ++v) rows[v][e] = cadd(rows[v][e], c8[v]); return; }
    idx.push_back(i);
    for (int v = 0; v < 8; ++v) rows[v].push_back(c8[v]);
  };
  if (s->yo == 0) {  // the k = 0 mode carries the uniform background (ifield = 3)
    cplx c8[8];
    for (int v = 0; v < 8; ++v) c8[v] = mk(background[v], 0.0);
    push(0, c8);
  }
  for (int m = 0; m < nmodes; ++m) {
    const int ikx = k[3 * m], iky = k[3 * m + 1], ikz = k[3 * m + 2];
    if (ikx < 0 || ikx > s->nx / 2 || std::abs(iky) >= NYr / 2 || (!s->two_d && std::abs(ikz) >= s->nz / 2) || (s->two_d && ikz != 0)) {
      s->err = "laps_set_primitive_modes: mode " + std::to_string(m) + " is outside the grid's half spectrum"; return 1;
    }
    const int kyr = (iky + NYr) % NYr;
    // internal axes: 3D (kx, ky, kz); 2D tree (kx, 0, ky) — the line axis carries the driver's ky
    const int ky = s->two_d ? 0 : kyr, kz = s->two_d ? kyr : (ikz + s->nz) % s->nz;
    if (s->tabW2.owner(ky) != s->rank) continue;                // another rank owns this row
    // the c2r x pass doubles 0 < kx < nx/2 and takes the real part of the kx = 0 and kx = nx/2 columns as they are:
    // Re(c exp(i (ky y + kz z))) (-1)^ix is exactly the reference's cosine sampled at the Nyquist kx
    const double w = (ikx > 0 && ikx < s->nx / 2) ? 0.5 : 1.0;
    cplx c8[8];
    for (int v = 0; v < 7; ++v) c8[v] = mk(w * coef[2 * ((size_t)v * nmodes + m)], w * coef[2 * ((size_t)v * nmodes + m) + 1]);
    c8[7] = mk(0.0, 0.0);
    push(((long long)ikx * s->nyl + s->tabW2.local(ky, s->rank)) * s->nz + kz, c8);
  }
  const int nent = (int)idx.size();
  for (int v = 0; v < 8; ++v) val.insert(val.end(), rows[v].begin(), rows[v].end());
  LAPS_CK(s, cudaMemsetAsync(s->uB, 0, 8 * s->csz * sizeof(cplx), s->stream));
  long long* d_idx = nullptr; cplx* d_val = nullptr;
  // stream-ordered temporaries (cudaFree would synchronise the whole device: ranks that share one device in a test
  // would wait for each other's flag kernels); released also on the error returns
  struct Release { long long*& a; cplx*& b; cudaStream_t st; ~Release() { if (a) cudaFreeAsync(a, st); if (b) cudaFreeAsync(b, st); } } release{d_idx, d_val, s->stream};
  if (nent > 0) {
    LAPS_CK(s, cudaMallocAsync((void**)&d_idx, nent * sizeof(long long), s->stream));
    LAPS_CK(s, cudaMallocAsync((void**)&d_val, (size_t)8 * nent * sizeof(cplx), s->stream));
    LAPS_CK(s, cudaMemcpyAsync(d_idx, idx.data(), nent * sizeof(long long), cudaMemcpyHostToDevice, s->stream));
    LAPS_CK(s, cudaMemcpyAsync(d_val, val.data(), (size_t)8 * nent * sizeof(cplx), cudaMemcpyHostToDevice, s->stream));
    LaunchScope ls(s, "scatter_modes");
    LAPS_LAUNCH(k_scatter_modes, dim3((unsigned)((8 * nent + 255) / 256)), dim3(256), 0, s->stream, s->uB, s->csz,
                (const long long*)d_idx, (const cplx*)d_val, nent, 8);
    LAPS_TRY(check_launch(s, "k_scatter_modes"));
  }
  {  // inverse transform of the 8 sparse spectra straight into uu (as primitives)
    ZParams z; fill_zparams(s, z);
    z.u_in = s->uB;
    for (int v = 0; v < 8; ++v) {
      ZTask t = blank_task();
      t.kind = kZInverseOnly; t.v = v; t.gout = v; t.fa = t.fb = t.fx = t.fc = -1;
      z.task[v] = t;
    }
    LAPS_TRY(host_barrier(s));   // the peers may still be reading their V1 (inverse y pass of the last stage)
    LAPS_TRY(spec_z(s, z, 8, "inv_z"));
    LAPS_TRY(host_barrier(s));
    LAPS_TRY(inverse_yx(s, 0, 8, false));
  }
  LAPS_CK(s, cudaStreamSynchronize(s->stream));
  return finish_set_primitive(s);
}

static int set_time_body(laps_handle s, double time) {  // AEBmod.f90:56-73
  if (!s) return 1;
  s->front_ready = false;
  LAPS_TRY(settle_exchange(s));
  const double old = s->radius;
  s->radius = s->p.radius0 + s->Ur * time;
  aeb_calc(s);
  s->ksq_initial = false;
  if (s->radius != old) s->j_stale = true;
  return upload_tables(s);
}

int laps_rkt_init(laps_handle s, double dt) {  // rktmod.f90:15-32 (fnl_rk is never read in stage 1)
  if (!s) return 1;
  const double cc10 = 8.0 / 15.0, cc20 = 5.0 / 12.0, cc30 = 0.75;
  const double dd20 = -17. / 60., dd30 = -5. / 12.;
  const double ts1 = 8. / 15., ts2 = 2. / 15., ts3 = 1. / 3.;
  s->cc1[0] = cc10 * dt; s->dd1[0] = 0.0;
  s->cc1[1] = cc20 * dt; s->dd1[1] = dd20 * dt;
  s->cc1[2] = cc30 * dt; s->dd1[2] = dd30 * dt;
  s->tstep[0] = ts1 * dt; s->tstep[1] = ts2 * dt; s->tstep[2] = ts3 * dt;
  s->dt = dt;
  return 0;
}

// dt from the three global maxima in h_scal (mhd.f90:404-428), hysteresis, rkt_init
static int vardt_finish(laps_handle s, double* dt_inout) {
  const laps_params& p = s->p;
  const double dx = p.Lx / s->nx, dy = s->two_d ? p.Lz / s->nz : p.Ly / s->ny, dz = p.Lz / s->nz;
  const double rr = s->radius / p.radius0;
  double dtmin;
  if (s->two_d) {                          // 2D/mhd.f90:375-381
    double dtx = dx / s->h_scal[0];
    if (p.if_AEB && p.if_z_radial) dtx = dtx * rr;
    const double dty = dy / s->h_scal[1] * rr;
    dtmin = std::min(dtx, dty);
  } else {
    const double dtx = dx / s->h_scal[0];
    const double dty = dy / s->h_scal[1] * rr;
    const double dtz = dz / s->h_scal[2] * rr;
    dtmin = std::min(std::min(dtx, dty), dtz);
  }
  dtmin = dtmin * p.cfl;
  double dt = *dt_inout;
  if (s->two_d && p.if_limit_dt_increase) {   // 2D/mhd.f90:396-400
    if (dt == 0.0 || dt > 1.02 * dtmin) dt = dtmin;
  } else if (dt < 0.98 * dtmin || dt > 1.02 * dtmin) dt = dtmin;
  *dt_inout = dt;
  return laps_rkt_init(s, dt);
}

static int vardt_body(laps_handle s, double* dt_inout) {  // mhd.f90:328-429
  if (!s || !dt_inout) return 1;
  LAPS_TRY(require_state(s));
  CflParams c;
  fill_cfl_params(s, c);
  {
    LaunchScope ls(s, "cfl", 8 * bytes_real(s));
    if (s->incomp) LAPS_LAUNCH(k_cfl_incomp, dim3((unsigned)s->nblk), dim3(256), 0, s->stream, c);   // src_incompressible/mhd.f90:369-476
    else LAPS_LAUNCH(k_cfl, dim3((unsigned)s->nblk), dim3(256), 0, s->stream, c);
    LAPS_TRY(check_launch(s, "k_cfl"));
  }
  LAPS_TRY(reduce_final(s, 3, 2, 0.0));   // global maxima of the three signal speeds (mhd.f90:419 as max)
  return vardt_finish(s, dt_inout);
}

static int evolve_body(laps_handle s) {  // mhd.f90:298-326
  if (!s) return 1;
  LAPS_TRY(require_state(s));
  for (auto& pe : s->prof) { cudaEventDestroy(pe.e0); cudaEventDestroy(pe.e1); }
  s->prof.clear();
  s->launches = 0;
  LAPS_CK(s, cudaEventRecord(s->ev0, s->stream));
  for (int irk = 0; irk < 3; ++irk) LAPS_TRY(stage(s, irk));
  LAPS_CK(s, cudaEventRecord(s->ev1, s->stream));
  if (s->incomp) {  // update_rho_p (src_incompressible/mhd.f90:366, AEBmod.f90:123-134): compounds with the CURRENT radius
    const double q = s->p.radius0 / s->radius;
    s->rho0 = s->rho0 * std::pow(q, 2);
    s->p0 = s->p0 * std::pow(q, 2 * s->p.adiabatic_index);
  }
  return 0;
}

static int step_body(laps_handle s, double* time_inout, double* dt_inout) {  // mhd.f90:245-248,285
  if (!s || !time_inout || !dt_inout) return 1;
  LAPS_TRY(laps_evolve(s));
  *time_inout = *time_inout + s->dt;
  LAPS_TRY(laps_set_time(s, *time_inout));
  *dt_inout = s->dt;
  if (!can_speculate(s)) return laps_vardt(s, dt_inout);
  // vardt reads the same state as the next step's calc_flux: one sweep serves both.  The launches below are the
  // dt-independent front half of the next step's first stage; the host waits only for the three maxima.
  LAPS_TRY(stage_front(s, true));
  LAPS_CK(s, cudaEventRecord(s->ev1, s->stream));   // laps_last_step_ms: this step's evolve + the next step's front half
  LAPS_TRY(reduce_wait(s));
  s->front_ready = true;
  return vardt_finish(s, dt_inout);
}

static int last_step_ms_body(laps_handle s, float* ms, int32_t* launches) {
  if (!s) return 1;
  LAPS_CK(s, cudaEventSynchronize(s->ev1));
  if (ms) LAPS_CK(s, cudaEventElapsedTime(ms, s->ev0, s->ev1));
  if (launches) *launches = s->launches;
  return 0;
}

int laps_get_pruning(laps_handle s, int32_t* nkx, int32_t* kymax, int32_t* nky_local) {
  if (!s) return 1;
  if (nkx) *nkx = s->nkx;
  if (kymax) *kymax = s->kymax;
  if (nky_local) *nky_local = s->pr_nkyl;
  return 0;
}

int laps_get_pruning_counts(laps_handle s, int64_t* live_columns, int64_t* live_modes) {
  if (!s) return 1;
  if (live_columns) *live_columns = s->pr_ncol;
  if (live_modes) *live_modes = s->pr_modes;
  return 0;
}

int laps_get_field_counts(laps_handle s, int32_t* nf, int32_t* ni, int32_t* spec_rows) {
  if (!s) return 1;
  if (nf) *nf = s->nf;
  if (ni) *ni = s->ni;
  if (spec_rows) *spec_rows = s->incomp ? 3 : (s->mass_from_state ? 7 : 8);
  return 0;
}

int laps_set_profiling(laps_handle s, int32_t on) { if (!s) return 1; s->profiling = on != 0; return 0; }

static int get_profile_body(laps_handle s, char* names, float* ms, int32_t cap, int32_t* count) {
  if (!s || !count) return 1;
  LAPS_CK(s, cudaStreamSynchronize(s->stream));
#ifndef LAPS_EMU_BUILD
  if (s->xstream) LAPS_CK(s, cudaStreamSynchronize(s->xstream));   // the transpose kernels of a front half enqueued by laps_step
#endif
  int n = 0;
  for (auto& pe : s->prof) {
    if (n >= cap) break;
    if (names) std::memcpy(names + (size_t)n * 32, pe.name, 32);
    if (ms) LAPS_CK(s, cudaEventElapsedTime(ms + n, pe.e0, pe.e1));
    ++n;
  }
  *count = n;
  return 0;
}

int laps_get_profile_bytes(laps_handle s, double* bytes, int32_t cap, int32_t* count) {
  if (!s || !count) return 1;
  int n = 0;
  for (auto& pe : s->prof) {
    if (n >= cap) break;
    if (bytes) bytes[n] = pe.bytes;
    ++n;
  }
  *count = n;
  return 0;
}

int laps_get_footprint(laps_handle s, int64_t* device_bytes) {
  if (!s || !device_bytes) return 1;
  *device_bytes = (int64_t)s->dev_bytes;
  return 0;
}

// Measurement helper: the LAPS_TUNE_* switches that only select between equivalent kernels / launch shapes, settable on a
// live handle so that one process can time the alternatives on the same state.  Unknown names fail.
int laps_set_tune(laps_handle s, const char* name, int32_t value) {
  if (!s || !name) return 1;
  const std::string n(name);
  int* slot = n == "rhs" ? &s->tune_rhs : n == "rcg" ? &s->tune_rcg : n == "cgz" ? &s->tune_cgz : n == "z" ? &s->tune_z :
              n == "spec" ? &s->tune_spec : n == "overlap" ? &s->tune_overlap : n == "ovl_push" ? &s->ovl_push_ctas : n == "ovl_y" ? &s->ovl_y_warps : n == "ovl_z" ? &s->ovl_z_warps :
              n == "ovl_chunks" ? &s->ovl_chunks : n == "screen" ? &s->tune_screen : n == "tly" ? &s->tune_tly : nullptr;
  if (!slot) { s->err = "laps_set_tune: unknown switch '" + n + "'"; return 1; }
  *slot = value;
  s->front_ready = false;
  LAPS_TRY(settle_exchange(s));
  return 0;
}

static int max_div_fourier(laps_handle s, int v0, double* out) {
  DivbParams d;
  d.v0 = v0;
  d.u = s->uA; d.fstride = s->csz; d.ncol = (int)s->ncol; d.nz = s->nz; d.nyl = s->nyl; d.yoff = s->yo; d.ystride = s->ystride;
  d.kxr = s->kxr; d.kyr = s->kyr; d.kze = s->kze;
  d.radius0 = s->p.radius0; d.radius = s->radius; d.cosa = s->cosa; d.sina = s->sina;
  d.corot_k = (s->p.if_AEB && s->p.if_corotating) ? 1 : 0;
  d.mode2d = s->two_d; d.z_radial = s->two_d && s->p.if_AEB && s->p.if_z_radial;
  d.kzr = s->kzr;
  d.partial = s->d_partial;
  {
    LaunchScope ls(s, "divb", 3 * 16.0 * (double)s->csz);
    LAPS_LAUNCH(k_divb, dim3((unsigned)s->nblk), dim3(256), 0, s->stream, d);
    LAPS_TRY(check_launch(s, "k_divb"));
  }
  LAPS_TRY(reduce_final(s, 1, 2, 0.0));
  *out = s->h_scal[0];
  return 0;
}

static int max_divb_body(laps_handle s, double* out) {  // mhd.f90:522-570
  if (!s || !out) return 1;
  LAPS_TRY(require_state(s));
  return max_div_fourier(s, 4, out);
}

static int max_divv_body(laps_handle s, double* out) {  // src_incompressible/mhd.f90:620-668: max |k . (rho u)^| / rho0
  if (!s || !out) return 1;
  LAPS_TRY(require_state(s));
  LAPS_TRY(max_div_fourier(s, 1, out));
  *out = *out / (s->incomp ? s->rho0 : 1.0);
  return 0;
}

// calc_divB_real + calc_max_divB_real, calc_divV_real + calc_max_divV_real
// (src_incompressible/mhdrhs.f90:532-648, mhd.f90:672-732): maxima of |div B| and |div (rho u)/rho0| in REAL space.
static int max_div_real_body(laps_handle s, double out[2]) {
  if (!s || !out) return 1;
  LAPS_TRY(require_state(s));
  s->front_ready = false;   // the work buffers are used as scratch
  LAPS_TRY(settle_exchange(s));
  const bool prune = !s->spectrum_full;
  ZParams z; fill_zparams(s, z, prune);
  z.u_in = s->uA;
  for (int j = 0; j < 2; ++j) {
    ZTask t = blank_task();
    t.kind = kZDiv; t.v = j == 0 ? 4 : 1; t.cx = j == 0 ? 1.0 : (s->incomp ? s->rho0 : 1.0); t.gout = j; t.fa = t.fb = t.fx = t.fc = -1;
    z.task[j] = t;
  }
  LAPS_TRY(host_barrier(s));   // the peers may still be reading their V1 (inverse y pass of the last stage)
  LAPS_TRY(spec_z(s, z, 2, "div_inv_z"));
  LAPS_TRY(host_barrier(s));
  // the two real fields land in the flux work area (free between stages)
  RealDst d; std::memset(&d, 0, sizeof(d));
  d.ptr[0] = buf_F(s) + (size_t)(s->nf - 2) * s->npts; d.ptr[1] = buf_F(s) + (size_t)(s->nf - 1) * s->npts;
  if (!s->two_d) LAPS_TRY(inv_y(s, buf_V1(s), buf_V2(s), 2, prune));
  LAPS_TRY(inv_x(s, s->two_d ? buf_V1(s) : buf_V2(s), d, 2, prune));
  {
    LaunchScope ls(s, "absmax", 2 * bytes_real(s));
    LAPS_LAUNCH(k_absmax, dim3((unsigned)s->nblk), dim3(256), 0, s->stream, (const double*)d.ptr[0], s->npts, 2, s->d_partial);
    LAPS_TRY(check_launch(s, "k_absmax"));
  }
  LAPS_TRY(reduce_final(s, 2, 2, 0.0));
  out[0] = s->h_scal[0]; out[1] = s->h_scal[1];
  return 0;
}

// external_force(ix,iy,1,1) of the 2D compressible tree: the user routine calc_external_force_real
// (2D/mhdrhs.f90:480-531) stays in the driver, which hands its field over whenever it changes (it depends on
// `time` only, i.e. once per step); every stage transforms it with the fluxes and adds it to fnl(7).
static int set_external_force_body(laps_handle s, const double* force_local) {
  if (!s || !force_local) return 1;
  if (s->ext_slot < 0) { s->err = "laps_set_external_force: the handle was created without if_external_force"; return 1; }
  LAPS_CK(s, cudaMemcpyAsync(s->ext, force_local, s->npts * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  LAPS_CK(s, cudaStreamSynchronize(s->stream));   // the caller may reuse force_local
  return 0;
}

// checkNan (2D/mhd.f90:563-591, src_incompressible/2D/mhd.f90:745-773): is any uu(ix,iy,iz,1:nvar) a NaN, on any rank.
static int check_nan_body(laps_handle s, int32_t* is_nan) {
  if (!s || !is_nan) return 1;
  LAPS_TRY(require_state(s));
  {
    LaunchScope ls(s, "check_nan", 8 * bytes_real(s));
    LAPS_LAUNCH(k_nan_flag, dim3((unsigned)s->nblk), dim3(256), 0, s->stream, (const double*)s->uu, 8 * s->npts, s->d_partial);
    LAPS_TRY(check_launch(s, "k_nan_flag"));
  }
  LAPS_TRY(reduce_final(s, 1, 2, 0.0));   // MPI_MAX over the ranks
  *is_nan = s->h_scal[0] != 0.0 ? 1 : 0;
  return 0;
}

int laps_get_rho0(laps_handle s, double* rho0) {
  if (!s || !rho0) return 1;
  *rho0 = s->rho0;
  return 0;
}

static int moments(laps_handle s, double sums[18]) {
  {
    LaunchScope ls(s, "moments1", 8 * bytes_real(s));
    LAPS_LAUNCH(k_moments1, dim3((unsigned)s->nblk), dim3(256), 0, s->stream, s->uu, s->npts, s->p.adiabatic_index, s->d_partial, s->incomp ? 1 : 0);
    LAPS_TRY(check_launch(s, "k_moments1"));
  }
  LAPS_TRY(reduce_final(s, 18, 0, 0.0));
  for (int j = 0; j < 18; ++j) sums[j] = s->h_scal[j];
  return 0;
}

static int rms_body(laps_handle s, double out[19]) {  // mhdrms.f90:53-126
  if (!s || !out) return 1;
  LAPS_TRY(require_state(s));
  double sums[18];
  LAPS_TRY(moments(s, sums));
  const double n = (double)s->nx * s->ny * s->nz;   // size_grid (mhdrms.f90:20)
  for (int j = 0; j < 8; ++j) {
    const double ave = sums[j] / n, sq = sums[8 + j] / n;
    out[j] = ave;
    out[8 + j] = sq - ave * ave;
  }
  {
    LaunchScope ls(s, "moments2", 4 * bytes_real(s));
    LAPS_LAUNCH(k_moments2, dim3((unsigned)s->nblk), dim3(256), 0, s->stream, s->uu, s->npts, out[1], out[2], out[3], s->d_partial);
    LAPS_TRY(check_launch(s, "k_moments2"));
  }
  LAPS_TRY(reduce_final(s, 3, 0, 0.0));
  for (int j = 0; j < 3; ++j) out[16 + j] = s->h_scal[j] / n;
  return 0;
}

static int invariants_body(laps_handle s, double out[3]) {
  if (!s || !out) return 1;
  LAPS_TRY(require_state(s));
  double sums[18];
  LAPS_TRY(moments(s, sums));
  const double n = (double)s->nx * s->ny * s->nz;
  out[0] = sums[16] / n;
  out[1] = sums[17] / n;
  return laps_max_divb(s, &out[2]);
}

static int get_state_body(laps_handle s, double* uu_local, double* uu_prim_local) {
  if (!s) return 1;
  LAPS_TRY(require_state(s));
  if (uu_local) LAPS_CK(s, cudaMemcpyAsync(uu_local, s->uu, 8 * s->npts * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (uu_prim_local) {
    double* prim = prim_scratch(s);
    {
      LaunchScope ls(s, "cons_to_prim", (8 + 4) * bytes_real(s));
      LAPS_LAUNCH(k_cons_to_prim, dim3((unsigned)s->nblk), dim3(256), 0, s->stream, (const double*)s->uu, prim, s->npts, s->p.adiabatic_index, s->incomp ? 1 : 0);
      LAPS_TRY(check_launch(s, "k_cons_to_prim"));
    }
    LAPS_CK(s, cudaMemcpyAsync(uu_prim_local, prim, 4 * s->npts * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  }
  LAPS_CK(s, cudaStreamSynchronize(s->stream));
  return check_abort(s);
}

// The array output_uu writes (mhdoutput.f90:95-123): uu with rho u -> u and e -> p when output_primitive,
// else the conserved uu.  One 8-field device->host copy instead of the 12 fields of laps_get_state.
static int get_output_body(laps_handle s, double* out_local, int32_t primitive) {
  if (!s || !out_local) return 1;
  LAPS_TRY(require_state(s));
  const size_t fb = s->npts * sizeof(double);
  if (!primitive) {
    LAPS_CK(s, cudaMemcpyAsync(out_local, s->uu, 8 * fb, cudaMemcpyDeviceToHost, s->stream));
  } else {
    double* prim = prim_scratch(s);
    {
      LaunchScope ls(s, "cons_to_prim", (8 + 4) * bytes_real(s));
      LAPS_LAUNCH(k_cons_to_prim, dim3((unsigned)s->nblk), dim3(256), 0, s->stream, (const double*)s->uu, prim, s->npts, s->p.adiabatic_index, s->incomp ? 1 : 0);
      LAPS_TRY(check_launch(s, "k_cons_to_prim"));
    }
    LAPS_CK(s, cudaMemcpyAsync(out_local, s->uu, fb, cudaMemcpyDeviceToHost, s->stream));                                 // rho
    LAPS_CK(s, cudaMemcpyAsync(out_local + s->npts, prim, 3 * fb, cudaMemcpyDeviceToHost, s->stream));                     // u
    LAPS_CK(s, cudaMemcpyAsync(out_local + 4 * s->npts, s->uu + 4 * s->npts, 3 * fb, cudaMemcpyDeviceToHost, s->stream)); // B
    LAPS_CK(s, cudaMemcpyAsync(out_local + 7 * s->npts, prim + 3 * s->npts, fb, cudaMemcpyDeviceToHost, s->stream));      // p
  }
  LAPS_CK(s, cudaStreamSynchronize(s->stream));
  return check_abort(s);
}

// laps_get_output without the wait: the output array is packed into a device snapshot (stream-ordered behind the steps
// enqueued so far, 2 x 8R of HBM traffic) and leaves over PCIe on a copy stream while later steps run on the main one.
static int get_output_async_body(laps_handle s, double* out_local, int32_t primitive) {
  if (!out_local) return 1;
  LAPS_TRY(require_state(s));
  const size_t bytes = 8 * s->npts * sizeof(double);
  if (!s->snap) {
    LAPS_CK(s, cudaMalloc((void**)&s->snap, bytes));
    s->dev_bytes += bytes;
#ifndef LAPS_EMU_BUILD
    LAPS_CK(s, cudaStreamCreateWithFlags(&s->ostream, cudaStreamNonBlocking));
#endif
    LAPS_CK(s, cudaEventCreateWithFlags(&s->ev_snap, cudaEventDisableTiming));
    LAPS_CK(s, cudaEventCreateWithFlags(&s->ev_out, cudaEventDisableTiming));
  }
  if (s->out_pending) LAPS_CK(s, cudaStreamWaitEvent(s->stream, s->ev_out, 0));   // the previous copy still reads the snapshot
  {
    LaunchScope ls(s, "output_pack", 16 * bytes_real(s));
    LAPS_LAUNCH(k_output_pack, dim3((unsigned)s->nblk), dim3(256), 0, s->stream, (const double*)s->uu, s->snap, s->npts,
                s->p.adiabatic_index, s->incomp ? 1 : 0, primitive ? 1 : 0);
    LAPS_TRY(check_launch(s, "k_output_pack"));
  }
  LAPS_CK(s, cudaEventRecord(s->ev_snap, s->stream));
  LAPS_CK(s, cudaStreamWaitEvent(s->ostream, s->ev_snap, 0));
  LAPS_CK(s, cudaMemcpyAsync(out_local, s->snap, bytes, cudaMemcpyDeviceToHost, s->ostream));
  LAPS_CK(s, cudaEventRecord(s->ev_out, s->ostream));
  s->out_pending = true;
  return 0;
}

static int output_wait_body(laps_handle s) {
  if (!s->out_pending) return 0;
  LAPS_CK(s, cudaEventSynchronize(s->ev_out));
  s->out_pending = false;
  return check_abort(s);
}

static int get_spectral_body(laps_handle s, double* out) {
  if (!s || !out) return 1;
  LAPS_TRY(require_state(s));
  LAPS_CK(s, cudaMemcpyAsync(out, s->uA, 8 * s->csz * sizeof(cplx), cudaMemcpyDeviceToHost, s->stream));
  LAPS_CK(s, cudaStreamSynchronize(s->stream));
  return check_abort(s);
}

static int fft_forward_body(laps_handle s, const double* real_fields, int32_t nfields, double* spec_out) {
  if (!s || !real_fields || !spec_out) return 1;
  if (nfields < 1 || nfields > 8) { s->err = "laps_fft_forward: 1..8 fields per call"; return 1; }
  s->front_ready = false;   // the work buffers are used as scratch
  LAPS_TRY(settle_exchange(s));
  // uses the flux work buffers and u_B as scratch; the state (u_A, uu) is untouched
  LAPS_CK(s, cudaMemcpyAsync(buf_F(s), real_fields, (size_t)nfields * s->npts * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  LAPS_TRY(forward_xy(s, buf_F(s), s->npts, nfields, false));
  LAPS_TRY(host_barrier(s));
  ZParams z; fill_zparams(s, z);
  z.u_out = s->uB;
  for (int v = 0; v < nfields; ++v) {
    ZTask t = blank_task();
    t.kind = kZForwardOnly; t.v = v; t.gout = -1; t.fa = v; t.fb = t.fx = t.fc = -1;
    z.task[v] = t;
  }
  LAPS_TRY(spec_z(s, z, nfields, "fwd_z"));
  LAPS_CK(s, cudaMemcpyAsync(spec_out, s->uB, (size_t)nfields * s->csz * sizeof(cplx), cudaMemcpyDeviceToHost, s->stream));
  LAPS_CK(s, cudaMemsetAsync(s->uB, 0, 8 * s->csz * sizeof(cplx), s->stream));   // u_B must keep its masked columns zero
  LAPS_CK(s, cudaStreamSynchronize(s->stream));
  LAPS_TRY(host_barrier(s));
  return check_abort(s);
}

static int fft_inverse_body(laps_handle s, const double* spec_in, int32_t nfields, double* real_out) {
  if (!s || !spec_in || !real_out) return 1;
  if (nfields < 1 || nfields > 8) { s->err = "laps_fft_inverse: 1..8 fields per call"; return 1; }
  s->front_ready = false;   // the work buffers are used as scratch
  LAPS_TRY(settle_exchange(s));
  LAPS_CK(s, cudaMemcpyAsync(s->uB, spec_in, (size_t)nfields * s->csz * sizeof(cplx), cudaMemcpyHostToDevice, s->stream));
  ZParams z; fill_zparams(s, z);
  z.u_in = s->uB;
  for (int v = 0; v < nfields; ++v) {
    ZTask t = blank_task();
    t.kind = kZInverseOnly; t.v = v; t.gout = v; t.fa = t.fb = t.fx = t.fc = -1;
    z.task[v] = t;
  }
  LAPS_TRY(host_barrier(s));   // the peers may still be reading their V1 (inverse y pass of the last stage)
  LAPS_TRY(spec_z(s, z, nfields, "inv_z"));
  LAPS_CK(s, cudaMemsetAsync(s->uB, 0, 8 * s->csz * sizeof(cplx), s->stream));   // u_B must keep its masked columns zero
  LAPS_TRY(host_barrier(s));
  if (!s->two_d) LAPS_TRY(inv_y(s, buf_V1(s), buf_V2(s), nfields, false));
  // bufX holds V2 (the x pass's input) and uu must stay untouched: the real fields go to a temporary buffer
  double* tmp = nullptr;
  LAPS_CK(s, cudaMallocAsync((void**)&tmp, (size_t)nfields * s->npts * sizeof(double), s->stream));
  RealDst d; std::memset(&d, 0, sizeof(d));
  for (int v = 0; v < nfields; ++v) d.ptr[v] = tmp + (size_t)v * s->npts;
  int rc = inv_x(s, s->two_d ? buf_V1(s) : buf_V2(s), d, nfields, false);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(real_out, tmp, (size_t)nfields * s->npts * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) { s->err = std::string("laps_fft_inverse copy: ") + cudaGetErrorString(e); rc = 1; }
  }
  cudaFreeAsync(tmp, s->stream);
  return rc ? rc : check_abort(s);
}

int laps_transpose_yz_indexmap(laps_handle s, int64_t* out) {
  if (!s || !out) return 1;
  // destination of element (kx, ky, zl) of this rank's post-y-pass data, exactly as k_fwd_y stores it
  size_t i = 0;
  for (int kx = 0; kx < s->nxh; ++kx)
    for (int ky = 0; ky < s->ny; ++ky) {
      const int pq = s->tabW2.owner(ky);
      for (int zl = 0; zl < s->nzl; ++zl, ++i) {
        out[2 * i] = pq;
        out[2 * i + 1] = ((int64_t)kx * s->tabW2.len[pq] + s->tabW2.local(ky, pq)) * s->nz + s->zo + zl;
      }
    }
  return 0;
}

namespace {
// What a rank tells its peers about its exchange buffers (fits LAPS_PEER_BLOB_BYTES).
struct PeerBlob {
  uint32_t magic;
  int32_t rank, nranks, device;
  int64_t pid;
  uint64_t ptr[3];                 // bufY (V1), bufZ (W2), xblk as seen by the owning process
  cudaIpcMemHandle_t ipc[3];
};
static_assert(sizeof(PeerBlob) <= LAPS_PEER_BLOB_BYTES, "peer blob too large");
constexpr uint32_t kBlobMagic = 0x4c415053u;  // "LAPS"
}  // namespace

int laps_transpose_zy_indexmap(laps_handle s, int64_t* out) {
  if (!s || !out) return 1;
  // destination of element (kx, ky_local, z) of this rank's inverse-z output, exactly as the z pass stores it
  size_t i = 0;
  for (int kx = 0; kx < s->nxh; ++kx)
    for (int kyl = 0; kyl < s->nyl; ++kyl) {
      const int ky = s->yo + kyl * s->ystride;
      for (int z = 0; z < s->nz; ++z, ++i) {
        const int pq = s->tabV1.owner(z);
        out[2 * i] = pq;
        out[2 * i + 1] = ((int64_t)kx * s->ny + ky) * s->tabV1.len[pq] + (z - s->tabV1.off[pq]);
      }
    }
  return 0;
}

int laps_export_peer_blob(laps_handle s, void* blob) {
  if (!s || !blob) return 1;
  PeerBlob b; std::memset(&b, 0, sizeof(b));
  b.magic = kBlobMagic; b.rank = s->rank; b.nranks = s->P; b.device = s->p.device; b.pid = (int64_t)getpid();
  void* ptrs[3] = {s->bufY, s->bufZ, (void*)s->xblk};
  LAPS_ENTER(s);
  for (int j = 0; j < 3; ++j) {
    b.ptr[j] = (uint64_t)(uintptr_t)ptrs[j];
    LAPS_CK(s, cudaIpcGetMemHandle(&b.ipc[j], ptrs[j]));
  }
  std::memset(blob, 0, LAPS_PEER_BLOB_BYTES);
  std::memcpy(blob, &b, sizeof(b));
  return 0;
}

int laps_import_peer_blobs(laps_handle s, const void* blobs) {
  if (!s || !blobs) return 1;
  LAPS_ENTER(s);
  for (int q = 0; q < s->P; ++q) {
    if (q == s->rank) continue;
    PeerBlob b;
    std::memcpy(&b, (const char*)blobs + (size_t)q * LAPS_PEER_BLOB_BYTES, sizeof(b));
    if (b.magic != kBlobMagic || b.rank != q || b.nranks != s->P) { s->err = "laps_import_peer_blobs: blob " + std::to_string(q) + " is not rank " + std::to_string(q) + "'s export"; return 1; }
    void* ptrs[3];
    if (b.pid == (int64_t)getpid()) {
      // same process (one host thread per GPU): plain peer access to the owner's allocations
      if (b.device != s->p.device) {
        int can = 0;
        LAPS_CK(s, cudaDeviceCanAccessPeer(&can, s->p.device, b.device));
        if (!can) { s->err = "no peer access between devices " + std::to_string(s->p.device) + " and " + std::to_string(b.device); return 1; }
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { s->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); return 1; }
        (void)cudaGetLastError();
      }
      for (int j = 0; j < 3; ++j) ptrs[j] = (void*)(uintptr_t)b.ptr[j];
    } else {
      for (int j = 0; j < 3; ++j) {
        if (s->ipc_opened[q][j]) { ptrs[j] = s->ipc_opened[q][j]; continue; }
        LAPS_CK(s, cudaIpcOpenMemHandle(&ptrs[j], b.ipc[j], cudaIpcMemLazyEnablePeerAccess));
        s->ipc_opened[q][j] = ptrs[j];
      }
    }
    s->tabV1.base[q] = (cplx*)ptrs[0];
    s->tabW2.base[q] = (cplx*)ptrs[1];
    s->xp.blk[q] = (XchgBlock*)ptrs[2];
  }
  s->wired = true;
  return 0;
}

// Single process, one handle per GPU (each driven by its own host thread afterwards).
int laps_connect_local(laps_handle* handles, int32_t nranks) {
  if (!handles || nranks < 1 || nranks > LAPS_MAX_RANKS) return 1;
  std::vector<char> blobs((size_t)nranks * LAPS_PEER_BLOB_BYTES);
  for (int q = 0; q < nranks; ++q) {
    if (!handles[q] || handles[q]->P != nranks || handles[q]->rank != q) return 1;
    LAPS_TRY(laps_export_peer_blob(handles[q], blobs.data() + (size_t)q * LAPS_PEER_BLOB_BYTES));
  }
  for (int q = 0; q < nranks; ++q) LAPS_TRY(laps_import_peer_blobs(handles[q], blobs.data()));
  return 0;
}

// ---- entry points whose bodies are above: device selection, dead-handle check, and (collectives) release of the
// peers when this rank fails between two inter-rank waits
int laps_set_primitive(laps_handle s, const double* uu_local) {
  if (!s) return 1;
  LAPS_ENTER(s);
  const int rc = set_primitive_body(s, uu_local);
  if (rc) poison_peers(s);
  return rc;
}

int laps_set_primitive_modes(laps_handle s, int32_t nmodes, const int32_t* k, const double* coef, const double* background) {
  if (!s) return 1;
  LAPS_ENTER(s);
  const int rc = set_primitive_modes_body(s, nmodes, k, coef, background);
  if (rc) poison_peers(s);
  return rc;
}

int laps_vardt(laps_handle s, double* dt_inout) {
  if (!s) return 1;
  LAPS_ENTER(s);
  const int rc = vardt_body(s, dt_inout);
  if (rc) poison_peers(s);
  return rc;
}

int laps_evolve(laps_handle s) {
  if (!s) return 1;
  LAPS_ENTER(s);
  const int rc = evolve_body(s);
  if (rc) poison_peers(s);
  return rc;
}

int laps_step(laps_handle s, double* time_inout, double* dt_inout) {
  if (!s) return 1;
  LAPS_ENTER(s);
  const int rc = step_body(s, time_inout, dt_inout);
  if (rc) poison_peers(s);
  return rc;
}

int laps_max_divb(laps_handle s, double* out) {
  if (!s) return 1;
  LAPS_ENTER(s);
  const int rc = max_divb_body(s, out);
  if (rc) poison_peers(s);
  return rc;
}

int laps_max_divv(laps_handle s, double* out) {
  if (!s) return 1;
  LAPS_ENTER(s);
  const int rc = max_divv_body(s, out);
  if (rc) poison_peers(s);
  return rc;
}

int laps_max_div_real(laps_handle s, double out[2]) {
  if (!s) return 1;
  LAPS_ENTER(s);
  const int rc = max_div_real_body(s, out);
  if (rc) poison_peers(s);
  return rc;
}

int laps_check_nan(laps_handle s, int32_t* is_nan) {
  if (!s) return 1;
  LAPS_ENTER(s);
  const int rc = check_nan_body(s, is_nan);
  if (rc) poison_peers(s);
  return rc;
}

int laps_rms(laps_handle s, double out[19]) {
  if (!s) return 1;
  LAPS_ENTER(s);
  const int rc = rms_body(s, out);
  if (rc) poison_peers(s);
  return rc;
}

int laps_invariants(laps_handle s, double out[3]) {
  if (!s) return 1;
  LAPS_ENTER(s);
  const int rc = invariants_body(s, out);
  if (rc) poison_peers(s);
  return rc;
}

int laps_fft_forward(laps_handle s, const double* real_fields, int32_t nfields, double* spec_out) {
  if (!s) return 1;
  LAPS_ENTER(s);
  const int rc = fft_forward_body(s, real_fields, nfields, spec_out);
  if (rc) poison_peers(s);
  return rc;
}

int laps_fft_inverse(laps_handle s, const double* spec_in, int32_t nfields, double* real_out) {
  if (!s) return 1;
  LAPS_ENTER(s);
  const int rc = fft_inverse_body(s, spec_in, nfields, real_out);
  if (rc) poison_peers(s);
  return rc;
}

int laps_sync(laps_handle s) {
  if (!s) return 1;
  LAPS_ENTER(s);
  return sync_body(s);
}

int laps_set_time(laps_handle s, double time) {
  if (!s) return 1;
  LAPS_ENTER(s);
  return set_time_body(s, time);
}

int laps_last_step_ms(laps_handle s, float* ms, int32_t* launches) {
  if (!s) return 1;
  LAPS_ENTER(s);
  return last_step_ms_body(s, ms, launches);
}

int laps_get_profile(laps_handle s, char* names, float* ms, int32_t cap, int32_t* count) {
  if (!s) return 1;
  LAPS_ENTER(s);
  return get_profile_body(s, names, ms, cap, count);
}

int laps_set_external_force(laps_handle s, const double* force_local) {
  if (!s) return 1;
  LAPS_ENTER(s);
  return set_external_force_body(s, force_local);
}

int laps_get_state(laps_handle s, double* uu_local, double* uu_prim_local) {
  if (!s) return 1;
  LAPS_ENTER(s);
  return get_state_body(s, uu_local, uu_prim_local);
}

int laps_get_output(laps_handle s, double* out_local, int32_t primitive) {
  if (!s) return 1;
  LAPS_ENTER(s);
  return get_output_body(s, out_local, primitive);
}

int laps_get_output_async(laps_handle s, double* out_local, int32_t primitive) {
  if (!s) return 1;
  LAPS_ENTER(s);
  return get_output_async_body(s, out_local, primitive);
}

int laps_output_wait(laps_handle s) {
  if (!s) return 1;
  LAPS_ENTER(s);
  return output_wait_body(s);
}

int laps_get_spectral(laps_handle s, double* out) {
  if (!s) return 1;
  LAPS_ENTER(s);
  return get_spectral_body(s, out);
}

}  // extern "C"
