// Real-space pointwise kernels and reductions (memory-bound, FP64):
//   k_flux           calc_flux incl. Hall E and EBM energy source      (mhdrhs.f90:48-122)
//   k_prim_to_cons   initial_calc_conserve_variable                    (mhdinit.f90:1038-1056)
//   k_cons_to_prim   update_uu_prim_from_uu                            (mhdrhs.f90:282-294)
//   k_cfl            per-point CFL limit + min reduction (vardt)       (mhd.f90:352-416)
//   k_moments1/2     calc_rms sums                                     (mhdrms.f90:73-125)
//   k_divb           calc_max_divB                                     (mhd.f90:541-568)
// Primitive velocity/pressure are recomputed in registers from uu wherever they are needed, in
// the reference's expression order, so the uu_prim array of the reference never exists on the GPU.
#pragma once
#include "compat.h"

namespace laps {

struct Prim { double ux, uy, uz, p; };

// mhdrhs.f90:285-293
LAPS_D Prim prim_of(double rho, double mx, double my, double mz, double bx, double by, double bz,
                    double e, double gm1) {
  Prim q;
  q.ux = mx / rho;
  q.uy = my / rho;
  q.uz = mz / rho;
  q.p = (e - 0.5 * (mx * q.ux + my * q.uy + mz * q.uz + bx * bx + by * by + bz * bz)) * gm1;
  return q;
}

// ------------------------------------------------------------------ block reductions
template <class Op>
LAPS_D double block_reduce(double v, Op op, double* scratch /* >= 32 doubles */) {
  LAPS_UNROLL
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  double r = scratch[0];
  for (int i = 1; i < nw; ++i) r = op(r, scratch[i]);
  return r;
}
struct OpSum { LAPS_D double operator()(double a, double b) const { return a + b; } };
struct OpMin { LAPS_D double operator()(double a, double b) const { return a < b ? a : b; } };
struct OpMax { LAPS_D double operator()(double a, double b) const { return a > b ? a : b; } };

struct CflParams {
  const double* uu; size_t npts;
  double gamma, di;
  double dmin;            // min(dx,dy,dz) (mhd.f90:398); min(dx,dy) in the 2D tree (2D/mhd.f90:369)
  double floor_x, floor_y; // resistivity/dx, resistivity/dy of the 2D tree's explicit-resistivity limit (2D/mhd.f90:361-364), else 0
  int hall;
  int screen;       // 1: evaluate the FP64 signal speeds only where a cheap FP32 bound says they can raise a maximum (bit-identical)
  double* partial;  // [3][row_stride]: CTA b of this launch writes entry boff + b of every row
  int row_stride, boff;   // (the sweep may be split into several launches over z chunks: together they fill [0, row_stride))
};

// Signal speeds of one point along x, y, z, folded into best[] (mhd.f90:352-416; see k_cfl for why the
// slow-mode candidates are absent and why maxima are reduced instead of minima of dx/c).
LAPS_D void cfl_point(const CflParams& P, double rho, double Bx, double By, double Bz, const Prim& q, double (&best)[3]) {
  const double s2 = sqrt(2.0);
  const double cs2 = P.gamma * q.p / rho;
  const double sr = sqrt(rho);
  const double ca[3] = {Bx / sr, By / sr, Bz / sr};
  const double uvel[3] = {q.ux, q.uy, q.uz};
  const double ca2 = ca[0] * ca[0] + ca[1] * ca[1] + ca[2] * ca[2];
  const double cms2 = cs2 + ca2;
  double chall = 0.0;
  if (P.hall) chall = P.di / rho * fmax(fmax(Bx, By), Bz) / P.dmin;   // signed max, as mhd.f90:396-398
  LAPS_UNROLL
  for (int d = 0; d < 3; ++d) {
    const double cns = sqrt(fmax(cms2 * cms2 - 4 * cs2 * ca[d] * ca[d], 0.0));
    const double cf = sqrt(cms2 + cns) / s2;
    const double uu_ = uvel[d];
    double c = fabs(uu_ + cf);
    c = fmax(c, fabs(uu_ + ca[d]));
    c = fmax(c, fabs(uu_ - cf));
    c = fmax(c, fabs(uu_ - ca[d]));
    c = fmax(c, fabs(uu_));
    if (d == 0) c = fmax(c, P.floor_x);
    if (d == 1) c = fmax(c, P.floor_y);
    if (P.hall) c = fmax(c, chall);
    best[d] = fmax(best[d], c);
  }
}

// A cheap FP32 upper bound of the three signal speeds of one point, so that the FP64 sqrt/div chain of cfl_point runs only
// where it can raise a maximum.  With cf <= sqrt(cs^2 + ca^2) and |ca_d| <= |ca|:  c_d <= max(|u_d| + sqrt(cs^2 + ca^2),
// floor_d, c_hall).  The bound is evaluated in FP32 from the conserved variables (one reciprocal, one square root) and
// inflated by a slack that covers the FP32 rounding, including the cancellation in p = (gamma - 1)(e - ...): the error of
// sqrt(cs^2 + ca^2) is below 5e-4 (|u| + sqrt(cs^2 + ca^2)), the slack is 4e-3 of that scale.  A point is skipped only
// if its bound is below the running maxima `sb` (the CTA's, which hold exact values of other points), so the reduced
// maxima are bit-identical to evaluating every point.  NaNs fail the comparison and take the exact path.
LAPS_D bool cfl_may_raise(const CflParams& P, double rho, double mx, double my, double mz, double Bx, double By, double Bz,
                          double en, const double (&sb)[3]) {
  const float r = 1.0f / (float)rho;
  const float ax = fabsf((float)mx) * r, ay = fabsf((float)my) * r, az = fabsf((float)mz) * r;
  const float bx = (float)Bx, by = (float)By, bz = (float)Bz;
  const float b2 = bx * bx + by * by + bz * bz;
  const float m2r = ((float)mx * (float)mx + (float)my * (float)my + (float)mz * (float)mz) * r;
  const float p = fmaxf(((float)en - 0.5f * (m2r + b2)) * (float)(P.gamma - 1.0), 0.0f);
  const float c = sqrtf(((float)P.gamma * p + b2) * r);
  const float slack = 4e-3f * (ax + ay + az + c);
  float ch = 0.0f;
  if (P.hall) ch = (float)P.di * r * fmaxf(fmaxf(fabsf(bx), fabsf(by)), fabsf(bz)) / (float)P.dmin * 1.001f;
  const float ux_ = fmaxf(fmaxf(ax + c + slack, (float)P.floor_x * 1.001f), ch);
  const float uy_ = fmaxf(fmaxf(ay + c + slack, (float)P.floor_y * 1.001f), ch);
  const float uz_ = fmaxf(az + c + slack, ch);
  return !((double)ux_ <= sb[0] && (double)uy_ <= sb[1] && (double)uz_ <= sb[2]);
}

// The CTA's running maxima (bit patterns of non-negative doubles order like the values): read by every thread before a
// point, raised after an exact evaluation.  Stale reads only make the screen more conservative.
LAPS_D void cfl_shared_read(const unsigned long long* s_best, double (&sb)[3]) {
  LAPS_UNROLL
  for (int d = 0; d < 3; ++d) sb[d] = __longlong_as_double((long long)*reinterpret_cast<const volatile unsigned long long*>(s_best + d));
}
LAPS_D void cfl_shared_raise(unsigned long long* s_best, const double (&best)[3], const double (&sb)[3]) {
  LAPS_UNROLL
  for (int d = 0; d < 3; ++d)
    if (best[d] > sb[d]) atomicMax(s_best + d, (unsigned long long)__double_as_longlong(best[d]));
}

// The exact evaluation as an out-of-line call: it runs for a small fraction of the points once the screen is active, and
// keeping its live values (a dozen FP64 temporaries around the sqrt / div sequences) out of the calling loop leaves
// calc_flux at three CTAs per SM.
__device__ __noinline__ void cfl_point_call(double gamma, double di, double dmin, double floor_x, double floor_y, int hall,
                                            double rho, double Bx, double By, double Bz, double ux, double uy, double uz, double p,
                                            double& b0, double& b1, double& b2) {
  CflParams P;
  P.gamma = gamma; P.di = di; P.dmin = dmin; P.floor_x = floor_x; P.floor_y = floor_y; P.hall = hall;
  Prim q; q.ux = ux; q.uy = uy; q.uz = uz; q.p = p;
  double b[3] = {b0, b1, b2};
  cfl_point(P, rho, Bx, By, Bz, q, b);
  b0 = b[0]; b1 = b[1]; b2 = b[2];
}

struct FluxParams {
  const double* uu;     // [8][npts]
  const double* J;      // [3][npts] (Hall) or null
  double* F;            // [nf][fstride]: 0-2 mass, 3-11 momentum tensor, 12-14 E, 15-17 energy, 18 EBM source
  size_t npts;          // field stride of uu and J
  size_t in_off, count; // points [in_off, in_off + count) of the slab are processed (z-chunked launches) ...
  size_t fstride;       // ... into F[slot * fstride + (point - in_off)]
  int hall, aeb;
  double gamma, di, tau;
  int z_radial;         // 2D tree, radial direction along z (2D/mhdrhs.f90:96-100)
  int slot[19];         // field slot of each flux in F, < 0: not needed (the z fluxes of the 2D tree)
  CflParams cfl;        // k_flux<true>: the CFL maxima of vardt are taken in the same sweep (same points, same primitives)
};

// 3 CTAs per SM (<= 85 registers): 11 loads + 19 stores per point want the occupancy (measured: 2 CTAs/SM cost 17 %);
// the CFL variant reaches that through the out-of-line cfl_point_call
template <bool CFL>
__global__ void __launch_bounds__(256, 3) k_flux(const FluxParams P) {
  const size_t n = P.npts;
  const double gm1 = P.gamma - 1.0;
  double best[3] = {0.0, 0.0, 0.0};
  __shared__ unsigned long long s_best[3];
  if (CFL) {
    if (threadIdx.x < 3) s_best[threadIdx.x] = 0ull;
    __syncthreads();
  }
  for (size_t ii = blockIdx.x * (size_t)blockDim.x + threadIdx.x; ii < P.count; ii += (size_t)gridDim.x * blockDim.x) {
    const size_t i = P.in_off + ii;
    const double rho = P.uu[i], mx = P.uu[n + i], my = P.uu[2 * n + i], mz = P.uu[3 * n + i];
    const double Bx = P.uu[4 * n + i], By = P.uu[5 * n + i], Bz = P.uu[6 * n + i], en = P.uu[7 * n + i];
    const Prim q = prim_of(rho, mx, my, mz, Bx, By, Bz, en, gm1);
    if (CFL) {   // the signal speeds of vardt, only where they can raise a maximum (cfl_may_raise)
      double sb[3];
      cfl_shared_read(s_best, sb);
      if (!P.cfl.screen || cfl_may_raise(P.cfl, rho, mx, my, mz, Bx, By, Bz, en, sb)) {
        cfl_point_call(P.cfl.gamma, P.cfl.di, P.cfl.dmin, P.cfl.floor_x, P.cfl.floor_y, P.cfl.hall, rho, Bx, By, Bz, q.ux, q.uy, q.uz, q.p,
                       best[0], best[1], best[2]);
        cfl_shared_raise(s_best, best, sb);
      }
    }
    const double ux = q.ux, uy = q.uy, uz = q.uz, p = q.p;
    const double ptot = p + 0.5 * (Bx * Bx + By * By + Bz * Bz);
    const double udotb = ux * Bx + uy * By + uz * Bz;
    double* F = P.F + ii;
    // every flux is stored as soon as it is formed (few live values: the kernel wants 3 CTAs per SM)
#define LAPS_PUT(j, val) do { if (P.slot[j] >= 0) F[(size_t)P.slot[j] * P.fstride] = (val); } while (0)
    LAPS_PUT(0, mx); LAPS_PUT(1, my); LAPS_PUT(2, mz);
    LAPS_PUT(3, mx * ux - Bx * Bx + ptot);
    LAPS_PUT(4, my * ux - By * Bx);
    LAPS_PUT(5, mz * ux - Bz * Bx);
    LAPS_PUT(6, mx * uy - Bx * By);
    LAPS_PUT(7, my * uy - By * By + ptot);
    LAPS_PUT(8, mz * uy - Bz * By);
    LAPS_PUT(9, mx * uz - Bx * Bz);
    LAPS_PUT(10, my * uz - By * Bz);
    LAPS_PUT(11, mz * uz - Bz * Bz + ptot);
    double Ex = uz * By - uy * Bz;
    double Ey = ux * Bz - uz * Bx;
    double Ez = uy * Bx - ux * By;
    if (P.hall) {
      const double Jx = P.J[i], Jy = P.J[n + i], Jz = P.J[2 * n + i];
      const double dr = P.di / rho;
      Ex = Ex + dr * (Jy * Bz - Jz * By);
      Ey = Ey + dr * (Jz * Bx - Jx * Bz);
      Ez = Ez + dr * (Jx * By - Jy * Bx);
    }
    LAPS_PUT(12, Ex); LAPS_PUT(13, Ey); LAPS_PUT(14, Ez);
    const double h = en + ptot;
    LAPS_PUT(15, h * ux - udotb * Bx);
    LAPS_PUT(16, h * uy - udotb * By);
    LAPS_PUT(17, h * uz - udotb * Bz);
    if (P.aeb) {
      double src;
      if (P.z_radial)
        src = -2 * P.gamma / gm1 * p / P.tau - (Bx * Bx + By * By + 2.0 * Bz * Bz) / P.tau -
              (2 * mx * ux + 2 * my * uy + mz * uz) / P.tau;
      else
        src = -2 * P.gamma / gm1 * p / P.tau - (2.0 * Bx * Bx + By * By + Bz * Bz) / P.tau -
              (mx * ux + 2 * my * uy + 2 * mz * uz) / P.tau;
      LAPS_PUT(18, src);
    }
#undef LAPS_PUT
  }
  if (CFL) {
    __shared__ double scratch[32];
    LAPS_UNROLL
    for (int d = 0; d < 3; ++d) {
      const double r = block_reduce(best[d], OpMax(), scratch);
      if (threadIdx.x == 0) P.cfl.partial[(size_t)d * P.cfl.row_stride + P.cfl.boff + blockIdx.x] = r;
    }
  }
}

// in place: [rho,ux,uy,uz,bx,by,bz,p] -> [rho,rho u,B,e]
// (incompressible tree, src_incompressible/mhdinit.f90:1031-1042: uu(8) stays the pressure)
__global__ void __launch_bounds__(256) k_prim_to_cons(double* uu, size_t n, double gamma, int incomp) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double rho = uu[i], ux = uu[n + i], uy = uu[2 * n + i], uz = uu[3 * n + i];
    const double bx = uu[4 * n + i], by = uu[5 * n + i], bz = uu[6 * n + i], p = uu[7 * n + i];
    uu[n + i] = rho * ux;
    uu[2 * n + i] = rho * uy;
    uu[3 * n + i] = rho * uz;
    if (!incomp) uu[7 * n + i] = p / (gamma - 1) + 0.5 * (rho * (ux * ux + uy * uy + uz * uz) + bx * bx + by * by + bz * bz);
  }
}

// prim[4][n] = (ux,uy,uz,p) from conserved uu
// (incompressible tree: uu_prim has the velocity only, src_incompressible/mhdrhs.f90:235-242; the 4th
// row returns the pressure uu(8))
__global__ void __launch_bounds__(256) k_cons_to_prim(const double* uu, double* prim, size_t n, double gamma, int incomp) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const Prim q = prim_of(uu[i], uu[n + i], uu[2 * n + i], uu[3 * n + i], uu[4 * n + i], uu[5 * n + i],
                           uu[6 * n + i], uu[7 * n + i], gamma - 1.0);
    prim[i] = q.ux; prim[n + i] = q.uy; prim[2 * n + i] = q.uz; prim[3 * n + i] = incomp ? uu[7 * n + i] : q.p;
  }
}

// The 8-field array output_uu writes (mhdoutput.f90:95-123) into a snapshot buffer: primitive != 0 -> rho, u, B, p, else uu
__global__ void __launch_bounds__(256) k_output_pack(const double* uu, double* out, size_t n, double gamma, int incomp, int primitive) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double rho = uu[i], mx = uu[n + i], my = uu[2 * n + i], mz = uu[3 * n + i];
    const double bx = uu[4 * n + i], by = uu[5 * n + i], bz = uu[6 * n + i], en = uu[7 * n + i];
    out[i] = rho; out[4 * n + i] = bx; out[5 * n + i] = by; out[6 * n + i] = bz;
    if (primitive) {
      const Prim q = prim_of(rho, mx, my, mz, bx, by, bz, en, gamma - 1.0);
      out[n + i] = q.ux; out[2 * n + i] = q.uy; out[3 * n + i] = q.uz; out[7 * n + i] = incomp ? en : q.p;
    } else {
      out[n + i] = mx; out[2 * n + i] = my; out[3 * n + i] = mz; out[7 * n + i] = en;
    }
  }
}

// mhd.f90:352-416.  partial[d][b] = max over the block's points of the signal speed along d.
// The reference takes min over points of dx/cmax_x, dy/cmax_y*(R/R0), dz/cmax_z*(R/R0); a correctly
// rounded division (and the product with R/R0) is monotonic, so min_i fl(dx/c_i) == fl(dx/max_i c_i)
// bit for bit: reduce the three maxima and divide once on the host (laps_vardt).  The slow-mode
// candidates |u +- csl| are dropped, also exactly: csl <= cf holds in floating point (sqrt and the
// division by sqrt(2) are monotonic), so they can never exceed max(|u + cf|, |u - cf|).
__global__ void __launch_bounds__(256) k_cfl(const CflParams P) {
  __shared__ double scratch[32];
  __shared__ unsigned long long s_best[3];
  if (threadIdx.x < 3) s_best[threadIdx.x] = 0ull;
  __syncthreads();
  const size_t n = P.npts;
  double best[3] = {0.0, 0.0, 0.0};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double rho = P.uu[i], mx = P.uu[n + i], my = P.uu[2 * n + i], mz = P.uu[3 * n + i];
    const double Bx = P.uu[4 * n + i], By = P.uu[5 * n + i], Bz = P.uu[6 * n + i], en = P.uu[7 * n + i];
    double sb[3];
    cfl_shared_read(s_best, sb);
    if (P.screen && !cfl_may_raise(P, rho, mx, my, mz, Bx, By, Bz, en, sb)) continue;
    const Prim q = prim_of(rho, mx, my, mz, Bx, By, Bz, en, P.gamma - 1.0);
    cfl_point(P, rho, Bx, By, Bz, q, best);
    cfl_shared_raise(s_best, best, sb);
  }
  LAPS_UNROLL
  for (int d = 0; d < 3; ++d) {
    const double r = block_reduce(best[d], OpMax(), scratch);
    if (threadIdx.x == 0) P.partial[(size_t)d * P.row_stride + P.boff + blockIdx.x] = r;
  }
}

// final reduction of per-block partials: out[j] = op over b of partial[j*nb + b]; one block per j
template <class Op>
__global__ void __launch_bounds__(256) k_reduce_final(const double* partial, int nb, double* out, double init) {
  __shared__ double scratch[32];
  double v = init;
  Op op;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) v = op(v, partial[(size_t)blockIdx.x * nb + b]);
  const double r = block_reduce(v, op, scratch);
  if (threadIdx.x == 0) out[blockIdx.x] = r;
}

// mhdrms.f90:73-93: sums and sums of squares of (rho,u,B,p), plus sum(e) and sum(u.B) (invariants).
// partial layout [18][gridDim.x]
// Incompressible tree: entry 8 is the pressure uu(8) (the reference reads the non-existent uu_prim(:,:,:,4),
// src_incompressible/mhdrms.f90:73,83) and the energy invariant is (rho u^2 + B^2)/2.
__global__ void __launch_bounds__(256) k_moments1(const double* uu, size_t n, double gamma, double* partial, int incomp) {
  __shared__ double scratch[32];
  double s[18];
  LAPS_UNROLL
  for (int j = 0; j < 18; ++j) s[j] = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double rho = uu[i], mx = uu[n + i], my = uu[2 * n + i], mz = uu[3 * n + i];
    const double Bx = uu[4 * n + i], By = uu[5 * n + i], Bz = uu[6 * n + i], en = uu[7 * n + i];
    const Prim q = prim_of(rho, mx, my, mz, Bx, By, Bz, en, gamma - 1.0);
    const double f[8] = {rho, q.ux, q.uy, q.uz, Bx, By, Bz, incomp ? en : q.p};
    LAPS_UNROLL
    for (int j = 0; j < 8; ++j) { s[j] += f[j]; s[8 + j] += f[j] * f[j]; }
    s[16] += incomp ? 0.5 * (mx * q.ux + my * q.uy + mz * q.uz + Bx * Bx + By * By + Bz * Bz) : en;
    s[17] += q.ux * Bx + q.uy * By + q.uz * Bz;
  }
  LAPS_UNROLL
  for (int j = 0; j < 18; ++j) {
    const double r = block_reduce(s[j], OpSum(), scratch);
    if (threadIdx.x == 0) partial[(size_t)j * gridDim.x + blockIdx.x] = r;
  }
}

// mhdrms.f90:109-120: sum rho (u_i - ubar_i)^2 ; partial layout [3][gridDim.x]
__global__ void __launch_bounds__(256) k_moments2(const double* uu, size_t n, double ubx, double uby, double ubz,
                                                  double* partial) {
  __shared__ double scratch[32];
  double s0 = 0, s1 = 0, s2 = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double rho = uu[i];
    const double ux = uu[n + i] / rho, uy = uu[2 * n + i] / rho, uz = uu[3 * n + i] / rho;
    s0 += rho * (ux - ubx) * (ux - ubx);
    s1 += rho * (uy - uby) * (uy - uby);
    s2 += rho * (uz - ubz) * (uz - ubz);
  }
  double r = block_reduce(s0, OpSum(), scratch);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
  r = block_reduce(s1, OpSum(), scratch);
  if (threadIdx.x == 0) partial[gridDim.x + blockIdx.x] = r;
  r = block_reduce(s2, OpSum(), scratch);
  if (threadIdx.x == 0) partial[2 * (size_t)gridDim.x + blockIdx.x] = r;
}

// mhd.f90:541-568: max | kx Bx^ + ky By^ + kz Bz^ | over the local modes; u = [8][ncol][nz]
struct DivbParams {
  const cplx* u; size_t fstride; int ncol, nz, nyl, yoff, ystride;   // ky = yoff + kyl * ystride
  const double* kxr; const double* kyr; const double* kze;
  double radius0, radius, cosa, sina; int corot_k;
  int mode2d, z_radial;   // 2D tree: the line axis carries ky, kz = 0 (2D/mhd.f90:527-550)
  const double* kzr;      // 2D tree with if_corotating: raw line wave numbers (both components of k vary along the line)
  int v0;                 // first component of the vector: 4 = B (calc_max_divB), 1 = rho u (calc_max_divV, src_incompressible/mhd.f90:620-668)
  double* partial;
};
__global__ void __launch_bounds__(256) k_divb(const DivbParams P) {
  __shared__ double scratch[32];
  double best = 0.0;
  const size_t total = (size_t)P.ncol * P.nz;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int col = (int)(i / P.nz), kz = (int)(i % P.nz);
    const int kx = col / P.nyl, ky = P.yoff + (col % P.nyl) * P.ystride;
    const double kxr = P.kxr[kx], kyr = P.kyr[ky];
    double kxe = kxr, kye = __ddiv_rn(__dmul_rn(kyr, P.radius0), P.radius);
    if (P.corot_k) {
      kxe = __dadd_rn(__dmul_rn(kxr, P.cosa), __dmul_rn(kyr, P.sina));
      kye = __ddiv_rn(__dmul_rn(__dadd_rn(__dmul_rn(-kxr, P.sina), __dmul_rn(kyr, P.cosa)), P.radius0), P.radius);
    }
    if (P.z_radial) kxe = __ddiv_rn(__dmul_rn(kxr, P.radius0), P.radius);
    const double kzz = P.kze[kz];
    const cplx bx = P.u[(size_t)P.v0 * P.fstride + i], by = P.u[(size_t)(P.v0 + 1) * P.fstride + i], bz = P.u[(size_t)(P.v0 + 2) * P.fstride + i];
    double rx = kxe, ry = kzz;
    if (P.mode2d && P.corot_k) {   // 2D/mhd.f90:539-545
      const double kyl = P.kzr[kz];
      rx = __dadd_rn(__dmul_rn(kxr, P.cosa), __dmul_rn(kyl, P.sina));
      ry = __ddiv_rn(__dmul_rn(__dadd_rn(__dmul_rn(-kxr, P.sina), __dmul_rn(kyl, P.cosa)), P.radius0), P.radius);
    }
    const cplx d = P.mode2d ? cadd(cadd(cmul_i(bx, rx), cmul_i(by, ry)), cmul_i(bz, 0.0))
                            : cadd(cadd(cmul_i(bx, kxe), cmul_i(by, kye)), cmul_i(bz, kzz));
    best = fmax(best, sqrt(d.x * d.x + d.y * d.y));
  }
  const double r = block_reduce(best, OpMax(), scratch);
  if (threadIdx.x == 0) P.partial[blockIdx.x] = r;
}

// ------------------------------------------------------------------ incompressible tree (src_incompressible/)
// calc_flux_for_pressure (mhdrhs.f90:393-437) and calc_flux (mhdrhs.f90:25-84) in one sweep: both are
// functions of the same real fields.  F slots 0-2: Fp = -(rho u . grad) u + J x B ; 3-5: E = -u x B (+ Hall).
struct FluxIncParams {
  const double* uu;     // [8][npts]  rho, rho u, B, p
  const double* J;      // [3][npts]
  const double* G;      // [9][npts]  d u_b / d x_a at slot 3*b + a
  double* F;            // [6][npts]
  size_t npts;
  int hall;
  double di;
};

__global__ void __launch_bounds__(256) k_flux_incomp(const FluxIncParams P) {
  const size_t n = P.npts;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double rho = P.uu[i], mx = P.uu[n + i], my = P.uu[2 * n + i], mz = P.uu[3 * n + i];
    const double Bx = P.uu[4 * n + i], By = P.uu[5 * n + i], Bz = P.uu[6 * n + i];
    const double Jx = P.J[i], Jy = P.J[n + i], Jz = P.J[2 * n + i];
    double g[9];
    LAPS_UNROLL
    for (int j = 0; j < 9; ++j) g[j] = P.G[(size_t)j * n + i];
    double* F = P.F + i;
    F[0] = -mx * g[0] - my * g[1] - mz * g[2] + Jy * Bz - Jz * By;
    F[n] = -mx * g[3] - my * g[4] - mz * g[5] + Jz * Bx - Jx * Bz;
    F[2 * n] = -mx * g[6] - my * g[7] - mz * g[8] + Jx * By - Jy * Bx;
    const double ux = mx / rho, uy = my / rho, uz = mz / rho;   // uu_prim (mhdrhs.f90:235-242)
    double Ex = uz * By - uy * Bz;
    double Ey = ux * Bz - uz * Bx;
    double Ez = uy * Bx - ux * By;
    if (P.hall) {
      const double dr = P.di / rho;
      Ex = Ex + dr * (Jy * Bz - Jz * By);
      Ey = Ey + dr * (Jz * Bx - Jx * Bz);
      Ez = Ez + dr * (Jx * By - Jy * Bx);
    }
    F[3 * n] = Ex; F[4 * n] = Ey; F[5 * n] = Ez;
  }
}

// vardt of the incompressible tree (src_incompressible/mhd.f90:369-476): Alfven and flow speeds only.
// Same reduction as k_cfl: the three maxima of the signal speeds (division is monotonic).
__global__ void __launch_bounds__(256) k_cfl_incomp(const CflParams P) {
  __shared__ double scratch[32];
  const size_t n = P.npts;
  double best[3] = {0.0, 0.0, 0.0};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double rho = P.uu[i];
    const double sr = sqrt(rho);
    double chall = 0.0;
    if (P.hall) chall = P.di / rho * fmax(fmax(P.uu[4 * n + i], P.uu[5 * n + i]), P.uu[6 * n + i]) / P.dmin;
    LAPS_UNROLL
    for (int d = 0; d < 3; ++d) {
      const double ca = P.uu[(size_t)(4 + d) * n + i] / sr;
      const double u = P.uu[(size_t)(1 + d) * n + i] / rho;
      double c = fmax(fmax(fabs(u + ca), fabs(u - ca)), fabs(u));
      if (d == 0) c = fmax(c, P.floor_x);   // explicit-resistivity limit of the 2D tree (src_incompressible/2D/mhd.f90:436-439)
      if (d == 1) c = fmax(c, P.floor_y);
      if (P.hall) c = fmax(c, chall);
      best[d] = fmax(best[d], c);
    }
  }
  LAPS_UNROLL
  for (int d = 0; d < 3; ++d) {
    const double r = block_reduce(best[d], OpMax(), scratch);
    if (threadIdx.x == 0) P.partial[(size_t)d * P.row_stride + P.boff + blockIdx.x] = r;
  }
}

// max |f| over nfields real fields (calc_max_divB_real / calc_max_divV_real, src_incompressible/mhd.f90:672-732);
// partial layout [nfields][gridDim.x]
__global__ void __launch_bounds__(256) k_absmax(const double* f, size_t n, int nfields, double* partial) {
  __shared__ double scratch[32];
  for (int j = 0; j < nfields; ++j) {
    double best = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
      best = fmax(best, fabs(f[(size_t)j * n + i]));
    const double r = block_reduce(best, OpMax(), scratch);
    if (threadIdx.x == 0) partial[(size_t)j * gridDim.x + blockIdx.x] = r;
  }
}

// checkNan (2D/mhd.f90:563-591): 1.0 where any of the n values is a NaN; partial layout [gridDim.x]
__global__ void __launch_bounds__(256) k_nan_flag(const double* f, size_t n, double* partial) {
  __shared__ double scratch[32];
  double bad = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = f[i];
    if (v != v) bad = 1.0;
  }
  const double r = block_reduce(bad, OpMax(), scratch);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

// Sparse fill of the spectral buffer for laps_set_primitive_modes: u[v][idx[e]] = val[v][e]
__global__ void __launch_bounds__(256) k_scatter_modes(cplx* u, size_t fstride, const long long* idx, const cplx* val, int nent, int nfields) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nent * nfields) return;
  const int v = t / nent, e = t % nent;
  u[(size_t)v * fstride + (size_t)idx[e]] = val[(size_t)v * nent + e];
}

}  // namespace laps
