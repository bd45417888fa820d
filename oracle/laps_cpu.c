/* CPU restatement (C + OpenMP) of the LAPS 3D compressible Hall-MHD + expanding-box RK step.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY: nothing in the product package links or loads this file.  It is the
 * "port" that bench.py times as the CPU baseline (cpu_baseline, --impl reference) because the reference's own
 * Fortran + MPI + FFTW build cannot be produced in this image.  tests/test_cpu_port.py checks it against the
 * NumPy oracle (oracle/laps_oracle.py), which is pinned by golden vectors of the reference's own source executed
 * through oracle/fortran_exec.py and by analytic known answers.
 *
 * It keeps the reference's STRUCTURE (file:line relative to /root/reference/src_compressible/): module-level
 * arrays (mhdinit.f90:141-178), one field at a time through line-at-a-time 1-D transforms with strided
 * gather/scatter (fftw.f90:42-103,136-222; mhdrhs.f90:128-172), separate pointwise sweeps for calc_flux
 * (mhdrhs.f90:21-124), calc_rhs (:174-279), rkt (rktmod.f90:34-62), dealias (dealiasing.f90:70-112),
 * update_uu_prim_from_uu (mhdrhs.f90:282-294), vardt (mhd.f90:328-429).  OpenMP threads stand in for the MPI
 * ranks (the transposes of parallel.f90 are then plain strided access).  FFTW is replaced by transforms written
 * here (power-of-two sizes only): batched Stockham radix-4 transforms, LAPS_BL lines per SIMD batch, real lines
 * through a half-length complex transform as FFTW's r2c/c2r plans do (the per-line radix-2/4 version this file
 * started with stays selectable, cpu_set_fft(0), so that the gain can be stated).
 *
 * Arrays are C order a[v][iz][iy][ix] = Fortran a(ix,iy,iz,v); spectra [v][kz][ky][kx], kx = 0..nx/2.
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

typedef struct {
  int nx, ny, nz;
  double Lx, Ly, Lz, gamma;
  int if_resis, if_resis_exp; double eta;
  int if_visc, if_visc_exp; double nu;
  int if_conserve_background; double cfl;
  int dealias_option; double afx, afy, afz;
  int if_AEB, if_corotating; double radius0, Ur0, corotating_angle;
  int if_hall; double di;
} cpu_params;

typedef struct {
  cpu_params p;
  int nxh; size_t nr, nc;                 /* points per real field, modes per spectral field */
  double *uu, *prim, *flux, *J;           /* [8], [4], [19], [3] real fields */
  cplx *uf, *ff, *fnl, *fnl_rk, *jf;      /* [8], [19], [8], [8], [3] spectra */
  double *wnx, *wny, *wnz, *ksq;          /* wave numbers, k_square[kz][ky][kx] */
  double *filtx, *filty, *filtz;
  cplx *twx, *twy, *twz;                  /* exp(-2 pi i m / n) */
  double Ur, radius, tau, cosa, sina, time, dt;
  double cc1[3], dd1[3], tstep[3];
} cpu_state;

static const double kPi = 3.141592653589793;   /* mhdinit.f90:7 */

/* ------------------------------------------------------------------ 1-D transforms (stand-in for FFTW) */
/* In-place complex transform of m points, m a power of two: bit reversal, then the radix-2 passes taken two at a
 * time (one radix-4 style sweep over the line per pair).  tw = exp(-2 pi i q / ntab), q < ntab, ntab a multiple of m.
 * dir = -1 forward, +1 backward, unnormalised. */
static void fft_inplace(cplx* a, int m, const cplx* tw, int ntab, int dir) {
  for (int i = 1, j = 0; i < m; ++i) {             /* bit reversal */
    int bit = m >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { cplx t = a[i]; a[i] = a[j]; a[j] = t; }
  }
  int h = 1, lg = 0;
  while ((1 << lg) < m) ++lg;
  if (lg & 1) {                                    /* odd number of passes: the first one (twiddle 1) on its own */
    for (int i = 0; i < m; i += 2) { const cplx u = a[i], v = a[i + 1]; a[i] = u + v; a[i + 1] = u - v; }
    h = 2;
  }
  const cplx rot = dir < 0 ? -I : I;               /* exp(dir * i pi / 2) */
  for (; h < m; h <<= 2) {                         /* passes of half-length h and 2h together */
    const int s1 = ntab / (2 * h), s2 = ntab / (4 * h);
    for (int i = 0; i < m; i += 4 * h)
      for (int k = 0; k < h; ++k) {
        cplx w1 = tw[k * s1], w2 = tw[k * s2];
        if (dir > 0) { w1 = conj(w1); w2 = conj(w2); }
        const cplx w3 = w2 * rot;
        cplx* q = a + i + k;
        const cplx t1 = q[h] * w1, t3 = q[3 * h] * w1;
        const cplx a0 = q[0] + t1, a1 = q[0] - t1, a2 = q[2 * h] + t3, a3 = q[2 * h] - t3;
        const cplx u2 = a2 * w2, u3 = a3 * w3;
        q[0] = a0 + u2; q[2 * h] = a0 - u2; q[h] = a1 + u3; q[3 * h] = a1 - u3;
      }
  }
}

/* r2c of n real points through one complex transform of n/2 points (what FFTW's r2c plans do as well):
 * out[k], k = 0..n/2, unnormalised.  work: n/2 complex. */
static void rfft_line(const double* x, int n, const cplx* tw, cplx* work, cplx* out) {
  const int m = n / 2;
  for (int j = 0; j < m; ++j) work[j] = x[2 * j] + I * x[2 * j + 1];
  fft_inplace(work, m, tw, n, -1);
  out[0] = creal(work[0]) + cimag(work[0]);
  out[m] = creal(work[0]) - cimag(work[0]);
  for (int k = 1; k < m; ++k) {
    const cplx zk = work[k], zc = conj(work[m - k]);
    out[k] = 0.5 * ((zk + zc) - I * tw[k] * (zk - zc));
  }
}

/* c2r (unnormalised backward) of the half spectrum in[0..n/2] to n real points; like FFTW the imaginary parts of
 * the DC and Nyquist bins are ignored.  work: n/2 complex. */
static void irfft_line(const cplx* in, int n, const cplx* tw, cplx* work, double* x) {
  const int m = n / 2;
  const double x0 = creal(in[0]), xm = creal(in[m]);
  work[0] = (x0 + xm) + I * (x0 - xm);
  for (int k = 1; k < m; ++k) {
    const cplx xk = in[k], xc = conj(in[m - k]);
    work[k] = (xk + xc) + I * conj(tw[k]) * (xk - xc);
  }
  fft_inplace(work, m, tw, n, +1);
  for (int j = 0; j < m; ++j) { x[2 * j] = creal(work[j]); x[2 * j + 1] = cimag(work[j]); }
}

static cplx* twiddles(int n) {
  cplx* t = (cplx*)malloc(sizeof(cplx) * n);
  for (int m = 0; m < n; ++m) t[m] = cexp(-2.0 * kPi * I * m / n);
  return t;
}

static int maxdim(const cpu_state* s) {
  const int nx = s->p.nx, ny = s->p.ny, nz = s->p.nz;
  return nx > ny ? (nx > nz ? nx : nz) : (ny > nz ? ny : nz);
}

/* ------------------------------------------------------------------ batched transforms
 * The 3-D transforms below run LAPS_BL lines at a time through one Stockham autosort transform whose innermost,
 * contiguous index is the line (split real / imaginary planes, so every butterfly is straight SIMD arithmetic over the
 * batch: AVX-512 / AVX2 clones are picked at load time), and gather LAPS_BL neighbouring kx columns per cache line in
 * the y and z passes instead of one strided element per line.  Same arithmetic as a per-line plan; this is what makes
 * the CPU baseline a fair stand-in for an FFTW build (measured against the per-line version it replaces: see
 * DESIGN.md, CPU baseline). */
#define LAPS_BL 8
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define LAPS_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define LAPS_CLONES
#endif

/* radix-4 Stockham passes (+ one radix-2 pass when log2 m is odd) over [point][LAPS_BL]; returns 0 if the result is
 * in (xr, xi), 1 if it is in (yr, yi).  tw = exp(-2 pi i q / ntab), ntab a multiple of m. */
LAPS_CLONES static int fft_batch(double* restrict xr, double* restrict xi, double* restrict yr, double* restrict yi,
                                 int m, const cplx* tw, int ntab, int dir) {
  int n = m, st = 1, flip = 0;
  const double sg = dir < 0 ? 1.0 : -1.0;        /* forward: W = exp(-i..), j(b-d) enters with the signs below */
  while (n >= 4) {
    const int n1 = n / 4, step = ntab / n;
    for (int p = 0; p < n1; ++p) {
      cplx w1 = tw[p * step], w2 = tw[2 * p * step], w3 = tw[3 * p * step];
      if (dir > 0) { w1 = conj(w1); w2 = conj(w2); w3 = conj(w3); }
      const double w1r = creal(w1), w1i = cimag(w1), w2r = creal(w2), w2i = cimag(w2), w3r = creal(w3), w3i = cimag(w3);
      for (int q = 0; q < st; ++q) {
        const size_t ia = (size_t)(q + st * p) * LAPS_BL, ib = ia + (size_t)st * n1 * LAPS_BL;
        const size_t ic = ib + (size_t)st * n1 * LAPS_BL, id = ic + (size_t)st * n1 * LAPS_BL;
        const size_t o0 = (size_t)(q + st * 4 * p) * LAPS_BL, o1 = o0 + (size_t)st * LAPS_BL, o2 = o1 + (size_t)st * LAPS_BL, o3 = o2 + (size_t)st * LAPS_BL;
#pragma omp simd
        for (int b = 0; b < LAPS_BL; ++b) {
          const double ar = xr[ia + b], ai = xi[ia + b], br = xr[ib + b], bi = xi[ib + b];
          const double cr = xr[ic + b], ci = xi[ic + b], dr = xr[id + b], di = xi[id + b];
          const double apcr = ar + cr, apci = ai + ci, amcr = ar - cr, amci = ai - ci;
          const double bpdr = br + dr, bpdi = bi + di;
          /* j (b - d): real part -(bi - di), imaginary part (br - dr); forward uses amc - j(b-d) for output 1 */
          const double jr = -sg * (bi - di), ji = sg * (br - dr);
          yr[o0 + b] = apcr + bpdr; yi[o0 + b] = apci + bpdi;
          const double t1r = amcr - jr, t1i = amci - ji;
          yr[o1 + b] = w1r * t1r - w1i * t1i; yi[o1 + b] = w1r * t1i + w1i * t1r;
          const double t2r = apcr - bpdr, t2i = apci - bpdi;
          yr[o2 + b] = w2r * t2r - w2i * t2i; yi[o2 + b] = w2r * t2i + w2i * t2r;
          const double t3r = amcr + jr, t3i = amci + ji;
          yr[o3 + b] = w3r * t3r - w3i * t3i; yi[o3 + b] = w3r * t3i + w3i * t3r;
        }
      }
    }
    { double* t = xr; xr = yr; yr = t; t = xi; xi = yi; yi = t; }
    flip ^= 1; n = n1; st *= 4;
  }
  if (n == 2) {
    for (int q = 0; q < st; ++q) {
      const size_t ia = (size_t)q * LAPS_BL, ib = ia + (size_t)st * LAPS_BL;
#pragma omp simd
      for (int b = 0; b < LAPS_BL; ++b) {
        const double ar = xr[ia + b], ai = xi[ia + b], br = xr[ib + b], bi = xi[ib + b];
        yr[ia + b] = ar + br; yi[ia + b] = ai + bi; yr[ib + b] = ar - br; yi[ib + b] = ai - bi;
      }
    }
    flip ^= 1;
  }
  return flip;
}

typedef struct { double *xr, *xi, *yr, *yi; } batch_buf;
static batch_buf batch_alloc(int m) {
  batch_buf b;
  const size_t n = (size_t)m * LAPS_BL;
  b.xr = (double*)aligned_alloc(64, 4 * n * sizeof(double));
  b.xi = b.xr + n; b.yr = b.xi + n; b.yi = b.yr + n;
  return b;
}

/* c2c along a strided axis for nb (<= LAPS_BL) neighbouring columns: base[i * stride + b], i < m; scaled by `scale` */
LAPS_CLONES static void c2c_columns(cplx* base, size_t stride, int m, int nb, const cplx* tw, int dir, double scale, batch_buf B) {
  for (int i = 0; i < m; ++i) {
    const double* src = (const double*)(base + (size_t)i * stride);
    for (int b = 0; b < nb; ++b) { B.xr[(size_t)i * LAPS_BL + b] = src[2 * b]; B.xi[(size_t)i * LAPS_BL + b] = src[2 * b + 1]; }
    for (int b = nb; b < LAPS_BL; ++b) { B.xr[(size_t)i * LAPS_BL + b] = 0.0; B.xi[(size_t)i * LAPS_BL + b] = 0.0; }
  }
  const int f = fft_batch(B.xr, B.xi, B.yr, B.yi, m, tw, m, dir);
  const double* rr = f ? B.yr : B.xr; const double* ri = f ? B.yi : B.xi;
  for (int i = 0; i < m; ++i) {
    double* dst = (double*)(base + (size_t)i * stride);
    for (int b = 0; b < nb; ++b) { dst[2 * b] = rr[(size_t)i * LAPS_BL + b] * scale; dst[2 * b + 1] = ri[(size_t)i * LAPS_BL + b] * scale; }
  }
}

/* r2c of nb (<= LAPS_BL) real lines a[b * lstride + j] through one batched complex transform of n/2 points;
 * out[b * ostride + k] = X_b[k] * scale, k = 0..n/2 */
LAPS_CLONES static void r2c_lines(const double* a, size_t lstride, int n, int nb, const cplx* tw, cplx* out, size_t ostride, double scale, batch_buf B) {
  const int m = n / 2;
  for (int b = 0; b < LAPS_BL; ++b) {
    if (b < nb) { const double* x = a + (size_t)b * lstride; for (int j = 0; j < m; ++j) { B.xr[(size_t)j * LAPS_BL + b] = x[2 * j]; B.xi[(size_t)j * LAPS_BL + b] = x[2 * j + 1]; } }
    else for (int j = 0; j < m; ++j) { B.xr[(size_t)j * LAPS_BL + b] = 0.0; B.xi[(size_t)j * LAPS_BL + b] = 0.0; }
  }
  const int f = fft_batch(B.xr, B.xi, B.yr, B.yi, m, tw, n, -1);
  const double* zr = f ? B.yr : B.xr; const double* zi = f ? B.yi : B.xi;
  for (int b = 0; b < nb; ++b) {
    cplx* o = out + (size_t)b * ostride;
    o[0] = (zr[b] + zi[b]) * scale;
    o[m] = (zr[b] - zi[b]) * scale;
  }
  for (int k = 1; k < m; ++k) {
    const double wr = creal(tw[k]), wi = cimag(tw[k]);
    for (int b = 0; b < nb; ++b) {
      const double ar = zr[(size_t)k * LAPS_BL + b], ai = zi[(size_t)k * LAPS_BL + b];
      const double cr = zr[(size_t)(m - k) * LAPS_BL + b], ci = -zi[(size_t)(m - k) * LAPS_BL + b];   /* conj(Z[m-k]) */
      const double sr = ar + cr, si = ai + ci, dr = ar - cr, di = ai - ci;
      /* 0.5 * (s - i w d) */
      const double tr = wr * dr - wi * di, ti = wr * di + wi * dr;     /* w d */
      out[(size_t)b * ostride + k] = (0.5 * scale) * ((sr + ti) + I * (si - tr));
    }
  }
}

/* c2r (unnormalised backward) of nb half spectra in[b * istride + k] to real lines a[b * lstride + j]; the imaginary
 * parts of the DC and Nyquist bins are ignored, as FFTW's c2r does */
LAPS_CLONES static void c2r_lines(const cplx* in, size_t istride, int n, int nb, const cplx* tw, double* a, size_t lstride, batch_buf B) {
  const int m = n / 2;
  for (int b = 0; b < LAPS_BL; ++b) {
    if (b >= nb) { for (int k = 0; k < m; ++k) { B.xr[(size_t)k * LAPS_BL + b] = 0.0; B.xi[(size_t)k * LAPS_BL + b] = 0.0; } continue; }
    const cplx* x = in + (size_t)b * istride;
    const double x0 = creal(x[0]), xm = creal(x[m]);
    B.xr[b] = x0 + xm; B.xi[b] = x0 - xm;
    for (int k = 1; k < m; ++k) {
      const cplx xk = x[k], xc = conj(x[m - k]);
      const cplx v = (xk + xc) + I * conj(tw[k]) * (xk - xc);
      B.xr[(size_t)k * LAPS_BL + b] = creal(v); B.xi[(size_t)k * LAPS_BL + b] = cimag(v);
    }
  }
  const int f = fft_batch(B.xr, B.xi, B.yr, B.yi, m, tw, n, +1);
  const double* zr = f ? B.yr : B.xr; const double* zi = f ? B.yi : B.xi;
  for (int b = 0; b < nb; ++b) {
    double* x = a + (size_t)b * lstride;
    for (int j = 0; j < m; ++j) { x[2 * j] = zr[(size_t)j * LAPS_BL + b]; x[2 * j + 1] = zi[(size_t)j * LAPS_BL + b]; }
  }
}

static int g_per_line_fft = 0;   /* cpu_set_fft(0): the per-line transforms above (the version this file started with), for comparison */

/* fftw.f90:42-71 + 136-180: r2c along x (/nx), c2c along y (/ny), c2c along z (/nz); one field */
static void forward3d_per_line(const cpu_state* s, const double* a, cplx* w);
static void inverse3d_per_line(const cpu_state* s, cplx* w, double* a);

static void forward3d(const cpu_state* s, const double* a, cplx* w) {
  if (g_per_line_fft) { forward3d_per_line(s, a, w); return; }
  const int nx = s->p.nx, ny = s->p.ny, nz = s->p.nz, nxh = s->nxh;
  const int nbx = (nxh + LAPS_BL - 1) / LAPS_BL, nby = (ny + LAPS_BL - 1) / LAPS_BL;
#pragma omp parallel
  {
    batch_buf B = batch_alloc(maxdim(s));
#pragma omp for collapse(2) schedule(static)
    for (int iz = 0; iz < nz; ++iz)
      for (int jb = 0; jb < nby; ++jb) {
        const int iy = jb * LAPS_BL, nb = ny - iy < LAPS_BL ? ny - iy : LAPS_BL;
        r2c_lines(a + ((size_t)iz * ny + iy) * nx, (size_t)nx, nx, nb, s->twx, w + ((size_t)iz * ny + iy) * nxh, (size_t)nxh, 1.0 / nx, B);
      }
#pragma omp for collapse(2) schedule(static)
    for (int iz = 0; iz < nz; ++iz)
      for (int kb = 0; kb < nbx; ++kb) {
        const int kx = kb * LAPS_BL, nb = nxh - kx < LAPS_BL ? nxh - kx : LAPS_BL;
        c2c_columns(w + (size_t)iz * ny * nxh + kx, (size_t)nxh, ny, nb, s->twy, -1, 1.0 / ny, B);
      }
#pragma omp for collapse(2) schedule(static)
    for (int iy = 0; iy < ny; ++iy)
      for (int kb = 0; kb < nbx; ++kb) {
        const int kx = kb * LAPS_BL, nb = nxh - kx < LAPS_BL ? nxh - kx : LAPS_BL;
        c2c_columns(w + (size_t)iy * nxh + kx, (size_t)ny * nxh, nz, nb, s->twz, -1, 1.0 / nz, B);
      }
    free(B.xr);
  }
}

/* fftw.f90:73-103 + 182-222: unnormalised backward c2c along z, then y, then c2r along x (w is overwritten) */
static void inverse3d(const cpu_state* s, cplx* w, double* a) {
  if (g_per_line_fft) { inverse3d_per_line(s, w, a); return; }
  const int nx = s->p.nx, ny = s->p.ny, nz = s->p.nz, nxh = s->nxh;
  const int nbx = (nxh + LAPS_BL - 1) / LAPS_BL, nby = (ny + LAPS_BL - 1) / LAPS_BL;
#pragma omp parallel
  {
    batch_buf B = batch_alloc(maxdim(s));
#pragma omp for collapse(2) schedule(static)
    for (int iy = 0; iy < ny; ++iy)
      for (int kb = 0; kb < nbx; ++kb) {
        const int kx = kb * LAPS_BL, nb = nxh - kx < LAPS_BL ? nxh - kx : LAPS_BL;
        c2c_columns(w + (size_t)iy * nxh + kx, (size_t)ny * nxh, nz, nb, s->twz, +1, 1.0, B);
      }
#pragma omp for collapse(2) schedule(static)
    for (int iz = 0; iz < nz; ++iz)
      for (int kb = 0; kb < nbx; ++kb) {
        const int kx = kb * LAPS_BL, nb = nxh - kx < LAPS_BL ? nxh - kx : LAPS_BL;
        c2c_columns(w + (size_t)iz * ny * nxh + kx, (size_t)nxh, ny, nb, s->twy, +1, 1.0, B);
      }
#pragma omp for collapse(2) schedule(static)
    for (int iz = 0; iz < nz; ++iz)
      for (int jb = 0; jb < nby; ++jb) {
        const int iy = jb * LAPS_BL, nb = ny - iy < LAPS_BL ? ny - iy : LAPS_BL;
        c2r_lines(w + ((size_t)iz * ny + iy) * nxh, (size_t)nxh, nx, nb, s->twx, a + ((size_t)iz * ny + iy) * nx, (size_t)nx, B);
      }
    free(B.xr);
  }
}

/* ---- the per-line version (cpu_set_fft(0)) ---- */
/* fftw.f90:42-71 + 136-180: r2c along x (/nx), c2c along y (/ny), c2c along z (/nz); one field */
static void forward3d_per_line(const cpu_state* s, const double* a, cplx* w) {
  const int nx = s->p.nx, ny = s->p.ny, nz = s->p.nz, nxh = s->nxh;
#pragma omp parallel
  {
    cplx* line = (cplx*)malloc(sizeof(cplx) * (size_t)(maxdim(s) + nxh));
    cplx* half = line + maxdim(s);
#pragma omp for collapse(2) schedule(static)
    for (int iz = 0; iz < nz; ++iz)
      for (int iy = 0; iy < ny; ++iy) {
        rfft_line(a + ((size_t)iz * ny + iy) * nx, nx, s->twx, line, half);
        cplx* dst = w + ((size_t)iz * ny + iy) * nxh;
        for (int k = 0; k < nxh; ++k) dst[k] = half[k] / nx;
      }
#pragma omp for collapse(2) schedule(static)
    for (int iz = 0; iz < nz; ++iz)
      for (int kx = 0; kx < nxh; ++kx) {
        cplx* base = w + (size_t)iz * ny * nxh + kx;
        for (int i = 0; i < ny; ++i) line[i] = base[(size_t)i * nxh];
        fft_inplace(line, ny, s->twy, ny, -1);
        for (int i = 0; i < ny; ++i) base[(size_t)i * nxh] = line[i] / ny;
      }
#pragma omp for collapse(2) schedule(static)
    for (int iy = 0; iy < ny; ++iy)
      for (int kx = 0; kx < nxh; ++kx) {
        cplx* base = w + (size_t)iy * nxh + kx;
        for (int i = 0; i < nz; ++i) line[i] = base[(size_t)i * ny * nxh];
        fft_inplace(line, nz, s->twz, nz, -1);
        for (int i = 0; i < nz; ++i) base[(size_t)i * ny * nxh] = line[i] / nz;
      }
    free(line);
  }
}

/* fftw.f90:73-103 + 182-222: unnormalised backward c2c along z, then y, then c2r along x (w is overwritten) */
static void inverse3d_per_line(const cpu_state* s, cplx* w, double* a) {
  const int nx = s->p.nx, ny = s->p.ny, nz = s->p.nz, nxh = s->nxh;
#pragma omp parallel
  {
    cplx* line = (cplx*)malloc(sizeof(cplx) * (size_t)maxdim(s));
#pragma omp for collapse(2) schedule(static)
    for (int iy = 0; iy < ny; ++iy)
      for (int kx = 0; kx < nxh; ++kx) {
        cplx* base = w + (size_t)iy * nxh + kx;
        for (int i = 0; i < nz; ++i) line[i] = base[(size_t)i * ny * nxh];
        fft_inplace(line, nz, s->twz, nz, +1);
        for (int i = 0; i < nz; ++i) base[(size_t)i * ny * nxh] = line[i];
      }
#pragma omp for collapse(2) schedule(static)
    for (int iz = 0; iz < nz; ++iz)
      for (int kx = 0; kx < nxh; ++kx) {
        cplx* base = w + (size_t)iz * ny * nxh + kx;
        for (int i = 0; i < ny; ++i) line[i] = base[(size_t)i * nxh];
        fft_inplace(line, ny, s->twy, ny, +1);
        for (int i = 0; i < ny; ++i) base[(size_t)i * nxh] = line[i];
      }
#pragma omp for collapse(2) schedule(static)
    for (int iz = 0; iz < nz; ++iz)
      for (int iy = 0; iy < ny; ++iy)
        irfft_line(w + ((size_t)iz * ny + iy) * nxh, nx, s->twx, line, a + ((size_t)iz * ny + iy) * nx);
    free(line);
  }
}

/* ------------------------------------------------------------------ set-up */
static double* wave_numbers(int n, double L) {   /* mhdinit.f90:79-110 (Nyquist kept positive) */
  double* k = (double*)malloc(sizeof(double) * n);
  for (int i = 1; i <= n; ++i) k[i - 1] = (i <= n / 2 + 1) ? 2 * kPi * (i - 1) / L : 2 * kPi * (i - 1 - n) / L;
  return k;
}

static double filter_1d(double k, double L, int n, double af) {   /* dealiasing.f90:36-43 */
  const double aj = (5.0 + 6.0 * af) / 8.0, bj = (1.0 + 2.0 * af) / 2.0, cj = -(1.0 - 2 * af) / 8.0, w = k * L / n;
  return (aj + bj * cos(w) + cj * cos(2 * w)) / (1 + 2 * af * cos(w));
}

static void update_ksquare(cpu_state* s, int initial) {   /* mhdinit.f90:114-122, AEBmod.f90:87-124 */
  const int ny = s->p.ny, nz = s->p.nz, nxh = s->nxh;
  const double r0 = s->p.radius0, r = s->radius, c = s->cosa, sn = s->sina;
#pragma omp parallel for collapse(2) schedule(static)
  for (int iz = 0; iz < nz; ++iz)
    for (int iy = 0; iy < ny; ++iy)
      for (int ix = 0; ix < nxh; ++ix) {
        const double kx = s->wnx[ix], ky = s->wny[iy], kz = s->wnz[iz];
        double v;
        if (initial) v = kx * kx + ky * ky + kz * kz;
        else if (s->p.if_corotating)
          v = kx * kx * (c * c + (sn * r0 / r) * (sn * r0 / r)) + ky * ky * (sn * sn + (c * r0 / r) * (c * r0 / r)) +
              kx * ky * 2 * c * sn * (1 - (r0 / r) * (r0 / r)) + (kz * r0 / r) * (kz * r0 / r);
        else v = kx * kx + (ky * r0 / r) * (ky * r0 / r) + (kz * r0 / r) * (kz * r0 / r);
        s->ksq[((size_t)iz * ny + iy) * nxh + ix] = v;
      }
}

void* cpu_create(const cpu_params* p) {
  cpu_state* s = (cpu_state*)calloc(1, sizeof(cpu_state));
  s->p = *p;
  s->nxh = p->nx / 2 + 1;
  s->nr = (size_t)p->nx * p->ny * p->nz;
  s->nc = (size_t)s->nxh * p->ny * p->nz;
  s->uu = (double*)calloc(8 * s->nr, sizeof(double));
  s->prim = (double*)calloc(4 * s->nr, sizeof(double));
  s->flux = (double*)calloc(19 * s->nr, sizeof(double));
  s->J = (double*)calloc(3 * s->nr, sizeof(double));
  s->uf = (cplx*)calloc(8 * s->nc, sizeof(cplx));
  s->ff = (cplx*)calloc(19 * s->nc, sizeof(cplx));
  s->fnl = (cplx*)calloc(8 * s->nc, sizeof(cplx));
  s->fnl_rk = (cplx*)calloc(8 * s->nc, sizeof(cplx));
  s->jf = (cplx*)calloc(3 * s->nc, sizeof(cplx));
  s->ksq = (double*)calloc(s->nc, sizeof(double));
  if (!s->uu || !s->prim || !s->flux || !s->J || !s->uf || !s->ff || !s->fnl || !s->fnl_rk || !s->jf || !s->ksq) return NULL;
  s->wnx = wave_numbers(p->nx, p->Lx); s->wny = wave_numbers(p->ny, p->Ly); s->wnz = wave_numbers(p->nz, p->Lz);
  s->twx = twiddles(p->nx); s->twy = twiddles(p->ny); s->twz = twiddles(p->nz);
  s->filtx = (double*)malloc(sizeof(double) * s->nxh); s->filty = (double*)malloc(sizeof(double) * p->ny); s->filtz = (double*)malloc(sizeof(double) * p->nz);
  for (int i = 0; i < s->nxh; ++i) s->filtx[i] = filter_1d(s->wnx[i], p->Lx, p->nx, p->afx);
  for (int i = 0; i < p->ny; ++i) s->filty[i] = filter_1d(s->wny[i], p->Ly, p->ny, p->afy);
  for (int i = 0; i < p->nz; ++i) s->filtz[i] = filter_1d(s->wnz[i], p->Lz, p->nz, p->afz);
  s->Ur = p->if_AEB ? p->Ur0 : 0.0;           /* mhd.f90:88-90 */
  s->radius = p->radius0;
  s->tau = s->radius / s->Ur;
  const double ang = p->if_corotating ? p->corotating_angle : 0.0;
  s->cosa = cos(ang); s->sina = sin(ang);
  update_ksquare(s, 1);
  return s;
}

void cpu_destroy(void* h) {
  cpu_state* s = (cpu_state*)h;
  if (!s) return;
  free(s->uu); free(s->prim); free(s->flux); free(s->J); free(s->uf); free(s->ff); free(s->fnl); free(s->fnl_rk); free(s->jf);
  free(s->ksq); free(s->wnx); free(s->wny); free(s->wnz); free(s->twx); free(s->twy); free(s->twz);
  free(s->filtx); free(s->filty); free(s->filtz); free(s);
}

/* 1: batched SIMD transforms (default), 0: the per-line transforms */
void cpu_set_fft(int batched) { g_per_line_fft = !batched; }

/* all host threads, whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1) */
void cpu_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int cpu_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------ pointwise pieces */
static void update_prim(cpu_state* s) {   /* mhdrhs.f90:282-294 */
  const size_t n = s->nr; const double gm1 = s->p.gamma - 1.0;
  double* u = s->uu; double* q = s->prim;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) {
    q[i] = u[n + i] / u[i]; q[n + i] = u[2 * n + i] / u[i]; q[2 * n + i] = u[3 * n + i] / u[i];
    q[3 * n + i] = (u[7 * n + i] - 0.5 * (u[n + i] * q[i] + u[2 * n + i] * q[n + i] + u[3 * n + i] * q[2 * n + i] +
                                         u[4 * n + i] * u[4 * n + i] + u[5 * n + i] * u[5 * n + i] + u[6 * n + i] * u[6 * n + i])) * gm1;
  }
}

/* derivative vectors (imaginary parts), mhdrhs.f90:191-204 */
static inline void kvec(const cpu_state* s, int ix, int iy, int iz, double* kx, double* ky, double* kz) {
  const double r0 = s->p.radius0, r = s->radius;
  *kz = s->wnz[iz] * r0 / r;
  *ky = s->wny[iy] * r0 / r;
  *kx = s->wnx[ix];
  if (s->p.if_AEB && s->p.if_corotating) {
    *kx = s->wnx[ix] * s->cosa + s->wny[iy] * s->sina;
    *ky = (-s->wnx[ix] * s->sina + s->wny[iy] * s->cosa) * r0 / r;
  }
}

static void calc_current(cpu_state* s) {   /* mhdrhs.f90:296-362 */
  const int ny = s->p.ny, nz = s->p.nz, nxh = s->nxh; const size_t nc = s->nc;
#pragma omp parallel for collapse(2) schedule(static)
  for (int iz = 0; iz < nz; ++iz)
    for (int iy = 0; iy < ny; ++iy)
      for (int ix = 0; ix < nxh; ++ix) {
        double kx, ky, kz; kvec(s, ix, iy, iz, &kx, &ky, &kz);
        const size_t m = ((size_t)iz * ny + iy) * nxh + ix;
        const cplx bx = s->uf[4 * nc + m], by = s->uf[5 * nc + m], bz = s->uf[6 * nc + m];
        s->jf[m] = I * ky * bz - I * kz * by;
        s->jf[nc + m] = I * kz * bx - I * kx * bz;
        s->jf[2 * nc + m] = I * kx * by - I * ky * bx;
      }
  for (int v = 0; v < 3; ++v) inverse3d(s, s->jf + v * nc, s->J + v * s->nr);
}

static void calc_flux(cpu_state* s) {   /* mhdrhs.f90:21-124 */
  const size_t n = s->nr; const cpu_params* p = &s->p;
  if (p->if_hall) calc_current(s);
  const double* u = s->uu; const double* q = s->prim; double* f = s->flux;
  const double gam = p->gamma, tau = s->tau;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) {
    const double rho = u[i], mx = u[n + i], my = u[2 * n + i], mz = u[3 * n + i];
    const double Bx = u[4 * n + i], By = u[5 * n + i], Bz = u[6 * n + i], en = u[7 * n + i];
    const double ux = q[i], uy = q[n + i], uz = q[2 * n + i], P = q[3 * n + i];
    const double ptot = P + 0.5 * (Bx * Bx + By * By + Bz * Bz);
    const double udotb = ux * Bx + uy * By + uz * Bz;
    f[i] = mx; f[n + i] = my; f[2 * n + i] = mz;
    f[3 * n + i] = mx * ux - Bx * Bx + ptot; f[4 * n + i] = my * ux - By * Bx; f[5 * n + i] = mz * ux - Bz * Bx;
    f[6 * n + i] = mx * uy - Bx * By; f[7 * n + i] = my * uy - By * By + ptot; f[8 * n + i] = mz * uy - Bz * By;
    f[9 * n + i] = mx * uz - Bx * Bz; f[10 * n + i] = my * uz - By * Bz; f[11 * n + i] = mz * uz - Bz * Bz + ptot;
    double Ex = uz * By - uy * Bz, Ey = ux * Bz - uz * Bx, Ez = uy * Bx - ux * By;
    if (p->if_hall) {
      const double Jx = s->J[i], Jy = s->J[n + i], Jz = s->J[2 * n + i], dr = p->di / rho;
      Ex = Ex + dr * (Jy * Bz - Jz * By); Ey = Ey + dr * (Jz * Bx - Jx * Bz); Ez = Ez + dr * (Jx * By - Jy * Bx);
    }
    f[12 * n + i] = Ex; f[13 * n + i] = Ey; f[14 * n + i] = Ez;
    f[15 * n + i] = (en + ptot) * ux - udotb * Bx; f[16 * n + i] = (en + ptot) * uy - udotb * By; f[17 * n + i] = (en + ptot) * uz - udotb * Bz;
    if (p->if_AEB)
      f[18 * n + i] = -2 * gam / (gam - 1) * P / tau - (2.0 * Bx * Bx + By * By + Bz * Bz) / tau - (mx * ux + 2 * my * uy + 2 * mz * uz) / tau;
  }
}

static void calc_rhs(cpu_state* s) {   /* mhdrhs.f90:174-279 */
  const int ny = s->p.ny, nz = s->p.nz, nxh = s->nxh; const size_t nc = s->nc; const cpu_params* p = &s->p;
  static const double aebc[7] = {2.0, 2.0, 3.0, 3.0, 2.0, 1.0, 1.0};
#pragma omp parallel for collapse(2) schedule(static)
  for (int iz = 0; iz < nz; ++iz)
    for (int iy = 0; iy < ny; ++iy)
      for (int ix = 0; ix < nxh; ++ix) {
        double kxr, kyr, kzr; kvec(s, ix, iy, iz, &kxr, &kyr, &kzr);
        const cplx kx = I * kxr, ky = I * kyr, kz = I * kzr;
        const size_t m = ((size_t)iz * ny + iy) * nxh + ix;
        const cplx* ff = s->ff + m; cplx* fnl = s->fnl + m; const cplx* uf = s->uf + m;
#define FF(j) ff[(size_t)(j) * nc]
        fnl[0] = -(kx * FF(0) + ky * FF(1) + kz * FF(2));
        fnl[nc] = -(kx * FF(3) + ky * FF(4) + kz * FF(5));
        fnl[2 * nc] = -(kx * FF(6) + ky * FF(7) + kz * FF(8));
        fnl[3 * nc] = -(kx * FF(9) + ky * FF(10) + kz * FF(11));
        fnl[4 * nc] = kz * FF(13) - ky * FF(14);
        fnl[5 * nc] = kx * FF(14) - kz * FF(12);
        fnl[6 * nc] = ky * FF(12) - kx * FF(13);
        fnl[7 * nc] = -(kx * FF(15) + ky * FF(16) + kz * FF(17));
        if (p->if_AEB) {
          for (int v = 0; v < 7; ++v) fnl[(size_t)v * nc] -= aebc[v] * uf[(size_t)v * nc] / s->tau;
          fnl[7 * nc] += FF(18);
        }
#undef FF
        const double k2 = s->ksq[m];
        if (p->if_visc && p->if_visc_exp) for (int v = 1; v <= 3; ++v) fnl[(size_t)v * nc] -= p->nu * uf[(size_t)v * nc] * k2;
        if (p->if_resis && p->if_resis_exp && !(p->if_conserve_background && ix == 0 && iz == 0))
          for (int v = 4; v <= 6; ++v) fnl[(size_t)v * nc] -= p->eta * uf[(size_t)v * nc] * k2;
      }
}

static void rkt_and_dealias(cpu_state* s, int irk) {   /* rktmod.f90:34-62, dealiasing.f90:70-112 */
  const int ny = s->p.ny, nz = s->p.nz, nxh = s->nxh; const size_t nc = s->nc; const cpu_params* p = &s->p;
  const double cc = s->cc1[irk], dd = s->dd1[irk], ts = s->tstep[irk];
#pragma omp parallel for collapse(2) schedule(static)
  for (int iz = 0; iz < nz; ++iz)
    for (int iy = 0; iy < ny; ++iy)
      for (int ix = 0; ix < nxh; ++ix) {
        const size_t m = ((size_t)iz * ny + iy) * nxh + ix;
        const double k2 = s->ksq[m];
        int masked = 0;
        if (p->dealias_option == 1) {
          const double tx = s->wnx[ix] * p->Lx / (2 * kPi * p->nx), ty = s->wny[iy] * p->Ly / (2 * kPi * p->ny), tz = s->wnz[iz] * p->Lz / (2 * kPi * p->nz);
          masked = sqrt(tx * tx + ty * ty + tz * tz) > (1.0 / 3.0);
        }
        for (int v = 0; v < 8; ++v) {
          const size_t j = (size_t)v * nc + m;
          cplx un = cc * s->fnl[j] + dd * s->fnl_rk[j] + s->uf[j];
          s->fnl_rk[j] = s->fnl[j];
          if (v >= 1 && v <= 3 && p->if_visc && !p->if_visc_exp) un = un / (ts * k2 * p->nu + 1.0);
          if (v >= 4 && v <= 6 && p->if_resis && !p->if_resis_exp) un = un / (ts * k2 * p->eta + 1.0);
          if (masked) un = 0.0;
          if (p->dealias_option == 2) un = un * s->filtx[ix] * s->filty[iy] * s->filtz[iz];
          s->uf[j] = un;
        }
      }
}

/* ------------------------------------------------------------------ driver-level calls */
void cpu_set_primitive(void* h, const double* prim8) {   /* mhdinit.f90:1038-1056 + fftw.f90:42-71 */
  cpu_state* s = (cpu_state*)h; const size_t n = s->nr; const double gam = s->p.gamma;
  memcpy(s->uu, prim8, 8 * n * sizeof(double));
  double* u = s->uu; double* q = s->prim;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) {
    q[i] = u[n + i]; q[n + i] = u[2 * n + i]; q[2 * n + i] = u[3 * n + i]; q[3 * n + i] = u[7 * n + i];
    u[n + i] = u[i] * q[i]; u[2 * n + i] = u[i] * q[n + i]; u[3 * n + i] = u[i] * q[2 * n + i];
    u[7 * n + i] = q[3 * n + i] / (gam - 1) + 0.5 * (u[i] * (q[i] * q[i] + q[n + i] * q[n + i] + q[2 * n + i] * q[2 * n + i]) +
                                                     u[4 * n + i] * u[4 * n + i] + u[5 * n + i] * u[5 * n + i] + u[6 * n + i] * u[6 * n + i]);
  }
  for (int v = 0; v < 8; ++v) forward3d(s, s->uu + v * n, s->uf + v * s->nc);
  s->time = 0.0; s->dt = 0.0;
}

static void rkt_init(cpu_state* s, double dt) {   /* rktmod.f90:15-32 */
  memset(s->fnl_rk, 0, 8 * s->nc * sizeof(cplx));
  s->cc1[0] = 8.0 / 15.0 * dt; s->cc1[1] = 5.0 / 12.0 * dt; s->cc1[2] = 0.75 * dt;
  s->dd1[0] = 0.0; s->dd1[1] = -17.0 / 60.0 * dt; s->dd1[2] = -5.0 / 12.0 * dt;
  s->tstep[0] = 8.0 / 15.0 * dt; s->tstep[1] = 2.0 / 15.0 * dt; s->tstep[2] = 1.0 / 3.0 * dt;
}

double cpu_vardt(void* h) {   /* mhd.f90:328-429 */
  cpu_state* s = (cpu_state*)h; const size_t n = s->nr; const cpu_params* p = &s->p;
  const double dx = p->Lx / p->nx, dy = p->Ly / p->ny, dz = p->Lz / p->nz, s2 = sqrt(2.0), rr = s->radius / p->radius0;
  const double dmin = fmin(fmin(dx, dy), dz);
  const double* u = s->uu; const double* q = s->prim;
  double dtmin = INFINITY;
#pragma omp parallel for schedule(static) reduction(min : dtmin)
  for (size_t i = 0; i < n; ++i) {
    const double rho = u[i], cs2 = p->gamma * q[3 * n + i] / rho, sr = sqrt(rho);
    const double ca[3] = {u[4 * n + i] / sr, u[5 * n + i] / sr, u[6 * n + i] / sr};
    const double cms2 = cs2 + (ca[0] * ca[0] + ca[1] * ca[1] + ca[2] * ca[2]);
    double chall = 0.0;
    if (p->if_hall) chall = p->di / rho * fmax(fmax(u[4 * n + i], u[5 * n + i]), u[6 * n + i]) / dmin;   /* signed max, mhd.f90:396-398 */
    double cm[3];
    for (int d = 0; d < 3; ++d) {
      const double cns = sqrt(fmax(cms2 * cms2 - 4 * cs2 * ca[d] * ca[d], 0.0));
      const double cf = sqrt(cms2 + cns) / s2, csl = sqrt(fmax(cms2 - cns, 0.0)) / s2, v = q[(size_t)d * n + i];
      double c = fabs(v + cf);
      c = fmax(c, fabs(v + csl)); c = fmax(c, fabs(v + ca[d])); c = fmax(c, fabs(v - cf));
      c = fmax(c, fabs(v - csl)); c = fmax(c, fabs(v - ca[d])); c = fmax(c, fabs(v));
      if (p->if_hall) c = fmax(c, chall);
      cm[d] = c;
    }
    const double t = fmin(fmin(dx / cm[0], dy / cm[1] * rr), dz / cm[2] * rr);
    dtmin = fmin(dtmin, t);
  }
  dtmin = dtmin * p->cfl;
  if (s->dt < 0.98 * dtmin || s->dt > 1.02 * dtmin) s->dt = dtmin;
  rkt_init(s, s->dt);
  return s->dt;
}

static void evolve_radius(cpu_state* s, double t) {   /* AEBmod.f90:56-73 */
  s->radius = s->p.radius0 + s->Ur * t;
  s->tau = s->radius / s->Ur;
  update_ksquare(s, 0);
}

void cpu_evolve(void* h) {   /* mhd.f90:298-326 */
  cpu_state* s = (cpu_state*)h; const int nf = s->p.if_AEB ? 19 : 18;
  for (int irk = 0; irk < 3; ++irk) {
    calc_flux(s);
    for (int j = 0; j < nf; ++j) forward3d(s, s->flux + (size_t)j * s->nr, s->ff + (size_t)j * s->nc);   /* mhdrhs.f90:128-172 */
    calc_rhs(s);
    rkt_and_dealias(s, irk);
    for (int v = 0; v < 8; ++v) {   /* fftw.f90:73-103 (the transform overwrites its input: work on a copy, fnl is free now) */
      memcpy(s->fnl + (size_t)v * s->nc, s->uf + (size_t)v * s->nc, s->nc * sizeof(cplx));
      inverse3d(s, s->fnl + (size_t)v * s->nc, s->uu + (size_t)v * s->nr);
    }
    update_prim(s);
  }
}

double cpu_step(void* h) {   /* mhd.f90:244-248,285 */
  cpu_state* s = (cpu_state*)h;
  cpu_evolve(s);
  s->time = s->time + s->dt;
  evolve_radius(s, s->time);
  return cpu_vardt(s);
}

void cpu_get_state(void* h, double* uu8, double* prim4, double* time, double* dt) {
  cpu_state* s = (cpu_state*)h;
  if (uu8) memcpy(uu8, s->uu, 8 * s->nr * sizeof(double));
  if (prim4) memcpy(prim4, s->prim, 4 * s->nr * sizeof(double));
  if (time) *time = s->time;
  if (dt) *dt = s->dt;
}
