// Incompressible tree (src_incompressible/): the z pass of one RK stage for the variables coupled by
// the pressure projection — rho u (3 components), the pressure, and rho — for one (kx,ky) column
// at a time.  It replaces, without ever materialising flux_pressure_fourier, fnl or k_square:
//   fftw.f90:173-179 (forward z lines of the three Fp fields, "/nz")
//   mhdrhs.f90:468-518 calc_pressure_fourier     p^ = -(k . Fp^)/k^2        (0 where k^2 < 1e-10)
//   mhdrhs.f90:117-232 calc_rhs                  fnl(2:4) = Fp^ + ((k . Fp^)/k^2) k, fnl(1) = fnl(8) = 0,
//                                                expanding-box terms (fnl(8) uses the NEW p^), explicit viscosity
//   rktmod.f90:34-62 rkt, dealiasing.f90:70-112 dealias, fftw.f90:195-201 (inverse z lines),
//   parallel.f90:300-324 transpose_zy (stores go to the owner of each z)
// The magnetic field rows (fnl(5:7) = curl E) are ordinary kZRhs tasks of k_spec_z / k_rhs_z.
//
// Per column: three forward transforms (the scaled spectra are parked in thread-private stash
// slots), the projection in registers, then five update + inverse-transform rounds (p, rho u x 3, rho).
#pragma once
#include "spectral_z.cuh"

namespace laps {

template <int N, int CG>
struct ITile {
  typedef Geom<N> G;
  static constexpr int PITCH = G::pitch(1);
  static constexpr int NTHREADS = CG * G::NT;
  static constexpr int COLSTRIDE = PITCH + 3 * N;
  static constexpr size_t SMEM = (size_t)CG * COLSTRIDE * sizeof(cplx);
  static constexpr int BY_SMEM = (int)((227 * 1024) / (SMEM + 1024));
  static constexpr int WTHREADS = (NTHREADS + 31) / 32 * 32;   // registers and thread slots are handed out per warp
  static constexpr int BY_REGS = 65536 / (WTHREADS * 96);
  static constexpr int BY_THREADS = 2048 / WTHREADS;
  static constexpr int M0 = BY_SMEM < BY_REGS ? BY_SMEM : BY_REGS;
  static constexpr int M1 = M0 < BY_THREADS ? M0 : BY_THREADS;
  static constexpr int MINB = M1 < 1 ? 1 : (M1 > 32 ? 32 : M1);
};

template <int N, int CG>
__global__ void __launch_bounds__(ITile<N, CG>::NTHREADS, ITile<N, CG>::MINB)
k_incomp_z(const ZParams P) {
  typedef Geom<N> G;
  typedef Fft<N, -1> FF;
  typedef Fft<N, +1> FI;
  typedef ITile<N, CG> T;
  LAPS_DYN_SMEM(cplx, sm);
  const int tid = threadIdx.x;
  const int l = tid / G::NT, u = tid % G::NT;
  const int colm = z_column(P, blockIdx.x * CG + l);
  const bool live = colm >= 0;
  const int col = live ? colm : 0;
  const int kx = col / P.nyl;
  const int ky = P.yoff + (col % P.nyl) * P.ystride;
  const size_t coff = (size_t)col * N;
  cplx* W = sm + l * T::COLSTRIDE;   // padded work line of the transforms
  cplx* S0 = W + T::PITCH;           // thread-private slots e*NT + u (forward OUTPUT order)
  cplx* S1 = S0 + N;
  cplx* S2 = S1 + N;

  // derivative vectors (imaginary parts), mhdrhs.f90:134-150
  const double kxr = __ldg(P.kxr + kx), kyr = __ldg(P.kyr + ky);
  double kxe = kxr, kye = __ddiv_rn(__dmul_rn(kyr, P.radius0), P.radius);
  if (P.corot_k) {
    kxe = __dadd_rn(__dmul_rn(kxr, P.cosa), __dmul_rn(kyr, P.sina));
    kye = __ddiv_rn(__dmul_rn(__dadd_rn(__dmul_rn(-kxr, P.sina), __dmul_rn(kyr, P.cosa)), P.radius0), P.radius);
  }
  const double ksq_xy = ksq_xy_eval(P, kxr, kyr, kx, ky);
  const double dxy = (P.dealias_option == 1 || P.dealias_option == 3) ? __dadd_rn(__ldg(P.dax + kx), __ldg(P.day + ky))
                                             : ((P.dealias_option == 2) ? __ldg(P.dax + kx) : 0.0);
  const double dfy = (P.dealias_option == 2) ? __ldg(P.day + ky) : 0.0;

  cplx r[8];
  // ---------------- forward z of Fp1, Fp2 (parked) and Fp3 (registers) ----------------
  LAPS_UNROLL
  for (int j = 0; j < 3; ++j) {
    const cplx* s = P.W2 + (size_t)j * P.fstride + coff;
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) r[e] = live ? s[u + e * G::NT] : mk(0.0, 0.0);
    FF::first(r, u, W, P.tw);
    FF::finish(r, u, W, P.tw);
    if (j < 2) {
      cplx* S = j == 0 ? S0 : S1;
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) S[e * G::NT + u] = cscale(r[e], P.scale);
    }
    __syncthreads();  // every last-stage read of the work line is done before it is refilled
  }
  // ---------------- projection (calc_pressure_fourier + the momentum rows of calc_rhs) ----------------
  LAPS_UNROLL
  for (int e = 0; e < 8; ++e) {
    const int kz = FF::kout(u, e);
    // 2D tree (src_incompressible/2D/mhdrhs.f90:189): the line axis carries the reference's ky, kz = 0
    const double kl = __ldg(P.kze + kz);
    double kxx = kxe, kyy = P.mode2d ? kl : kye;
    const double kzz = P.mode2d ? 0.0 : kl;
    double k2 = __dadd_rn(ksq_xy, __ldg(P.ksq_z + kz));
    if (P.corot2d) {   // 2D tree with if_corotating (2D/mhdrhs.f90:196-201,593-598): both components vary along the line
      const double kyl = __ldg(P.kzr + kz);
      corot2d_k(P, kxr, kyl, kxx, kyy);
      if (P.corot_ksq) k2 = __dadd_rn(k2, corot2d_cross(P, kxr, kyl));
    }
    const cplx f1 = S0[e * G::NT + u], f2 = S1[e * G::NT + u], f3 = cscale(r[e], P.scale);
    const cplx sum = cadd(cadd(cmul_i(f1, kxx), cmul_i(f2, kyy)), cmul_i(f3, kzz));
    if (k2 < 1e-10) {   // "background field, not important in Fourier space" (mhdrhs.f90:155-159,505-508)
      S0[e * G::NT + u] = mk(0.0, 0.0); S1[e * G::NT + u] = mk(0.0, 0.0); S2[e * G::NT + u] = mk(0.0, 0.0);
      r[e] = mk(0.0, 0.0);
    } else {
      const cplx kd = mk(__ddiv_rn(sum.x, k2), __ddiv_rn(sum.y, k2));
      S0[e * G::NT + u] = cadd(f1, cmul_i(kd, kxx));
      S1[e * G::NT + u] = cadd(f2, cmul_i(kd, kyy));
      S2[e * G::NT + u] = cadd(f3, cmul_i(kd, kzz));
      r[e] = mk(-kd.x, -kd.y);   // p^
    }
  }
  // ---------------- five rounds: update one variable, inverse z, store ----------------
  // round 0: p (v = 7, r holds p^);  1-3: rho u (v = 1..3, fnl from the stash);  4: rho (v = 0, fnl = 0)
  LAPS_UNROLL
  for (int round = 0; round < 5; ++round) {
    const int v = round == 0 ? 7 : (round == 4 ? 0 : round);
    const size_t voff = (size_t)v * P.fstride + coff;
    const cplx* S = round == 1 ? S0 : (round == 2 ? S1 : S2);
    // expanding box (mhdrhs.f90:187-204): 2, 2, 3, 3 for rho, rho u; 2*gamma for p
    const double cab = v == 7 ? P.aeb_p : ((v == 2 || v == 3) ? 3.0 : 2.0);
    const double ca = P.aeb ? cab / P.tau : 0.0;
    const bool mom = v >= 1 && v <= 3;
    const double ce = (mom && P.visc_exp) ? P.nu : 0.0;    // mhdrhs.f90:206-216
    const double ci = (mom && P.visc_imp) ? P.nu : 0.0;    // rktmod.f90:47-52
    const bool need_ksq = (ce != 0.0) || (ci != 0.0);
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) {
      const int kz = FF::kout(u, e);
      cplx fnl = mom ? S[e * G::NT + u] : mk(0.0, 0.0);
      // the pressure row reads the value calc_pressure_fourier has just written (mhd.f90:338 precedes calc_rhs, :350)
      const bool keep = live && !z_mode_dead(P, dxy, kz);   // masked modes: zero in memory, not touched
      const cplx uo = round == 0 ? r[e] : (keep ? P.u_in[voff + kz] : mk(0.0, 0.0));
      fnl.x -= ca * uo.x;
      fnl.y -= ca * uo.y;
      double ksq = 0.0;
      if (need_ksq) {
        ksq = __dadd_rn(ksq_xy, __ldg(P.ksq_z + kz));
        if (P.corot2d && P.corot_ksq) ksq = __dadd_rn(ksq, corot2d_cross(P, kxr, __ldg(P.kzr + kz)));
        fnl.x -= (ce * uo.x) * ksq;
        fnl.y -= (ce * uo.y) * ksq;
      }
      cplx un;   // rkt (rktmod.f90:40-42)
      if (P.read_rk) {
        const cplx fr = keep ? P.fnl_rk[voff + kz] : mk(0.0, 0.0);
        un = mk((P.cc * fnl.x + P.dd * fr.x) + uo.x, (P.cc * fnl.y + P.dd * fr.y) + uo.y);
      } else {
        un = mk(P.cc * fnl.x + uo.x, P.cc * fnl.y + uo.y);
      }
      if (P.write_rk && keep) P.fnl_rk[voff + kz] = fnl;
      if (need_ksq) {
        const double inv = __drcp_rn(__dadd_rn(__dmul_rn(__dmul_rn(P.dt_irk, ksq), ci), 1.0));
        un.x *= inv;
        un.y *= inv;
      }
      if (P.dealias_option == 1) {   // dealiasing.f90:87-110
        if (__dadd_rn(dxy, __ldg(P.daz + kz)) >= P.da_thresh) un = mk(0.0, 0.0);
      } else if (P.dealias_option == 2) {
        const double fz = __ldg(P.daz + kz);
        un = mk(__dmul_rn(__dmul_rn(__dmul_rn(un.x, dxy), dfy), fz), __dmul_rn(__dmul_rn(__dmul_rn(un.y, dxy), dfy), fz));
      } else if (P.dealias_option == 3) {   // square truncation (2D/dealiasing.f90:99-114): per-axis flags
        if (dxy != 0.0 || __ldg(P.daz + kz) != 0.0) un = mk(0.0, 0.0);
      }
      if (keep) P.u_out[voff + kz] = un;
      r[e] = un;
    }
    // re-shape the register contents into the stage-0 input pattern of the inverse transform
    if constexpr (G::RLAST != 8) {
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) W[G::pad(FF::kout(u, e))] = r[e];
      __syncthreads();
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) r[e] = W[FI::in_pos(u, G::pad(u), e)];
      __syncthreads();
    }
    FI::first(r, u, W, P.tw);
    FI::finish(r, u, W, P.tw);
    if (live) {
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) {
        const int z = FI::kout(u, e);
        const int p = P.V1.owner(z);
        cplx* dst = P.V1.base[p] + (((size_t)v * P.nxh + kx) * P.ny + ky) * P.V1.len[p] + (z - P.V1.off[p]);
        *dst = r[e];
      }
    }
    __syncthreads();  // the work line is refilled by the next round
  }
}

}  // namespace laps
