// Fused z-pass: forward z-FFT of the flux fields + spectral right-hand side + Runge-Kutta stage
// update + implicit diffusion + dealiasing + inverse z-FFT of the updated state, one launch per
// RK stage.  It replaces, for one (kx,ky) column at a time and without ever materialising
// flux_fourier/fnl/k_square:
//   fftw.f90:173-179 (forward z lines, "/nz")      mhdrhs.f90:174-279 (calc_rhs)
//   rktmod.f90:34-62 (rkt)                         dealiasing.f90:70-112 (dealias)
//   fftw.f90:195-201 (inverse z lines)             AEBmod.f90:87-124 (k_square, on the fly)
// A second small launch (kind kZCurrent) forms J^ = i k x B^ (mhdrhs.f90:313-339) from the
// updated field and inverse-transforms it.
//
// Linearity is used to save transforms: kx and ky are constant along a z-line, so
//   -(kx F^a + ky F^b + kz F^c) = -( FFTz[kx fa + ky fb] + kz FFTz[fc] )
// needs two z-transforms instead of three (results differ from the reference's order of
// operations by round-off only).
#pragma once
#include "fft_passes.cuh"

namespace laps {

enum ZKind { kZRhs = 0, kZForwardOnly = 1, kZInverseOnly = 2, kZCurrent = 3,
             // incompressible tree (src_incompressible/mhdrhs.f90):
             kZGrad = 4,   // i k_a u^[v] / cx, a = jcomp   (calc_gradient_velocity_real, :312-391)
             kZDiv = 5,
             // fnl(1) = -i k . (rho u)^ taken from the spectral state itself: calc_flux sets flux(1:3) = uu(2:4)
             // (mhdrhs.f90:58-60), whose transform (mhdrhs.f90:128-172) is uu_fourier(2:4) again — no transform needed
             kZMass = 6 };  // i (kx u^[v] + ky u^[v+1] + kz u^[v+2]) / cx   (calc_divB_real / calc_divV_real, :532-648)

// One row of work (blockIdx.y).  For kZRhs:
//   G = ca*(i kx)*W2[fa] + cb*(i ky)*W2[fb] + cx*W2[fx]      (missing terms have index < 0)
//   fnl = sg*FFTz[G] + sc*(i kz)*FFTz[cf1*W2[fc] + cf2*W2[fc2]]
struct ZTask {
  int kind;
  int v;       // state component updated (0..7) / read (kZInverseOnly)
  int gout;    // slot of the inverse-z output in V1, < 0: none
  int fa, fb, fx, fc;
  double ca, cb, cx, sg, sc;
  // second field of the (i k_line) term: fnl = sg*FFTz[G] + sc*(i k_line)*FFTz[cf1*W2[fc] + cf2*W2[fc2]].  Used by the 2D tree
  // with if_corotating, where both components of the rotated wave vector vary along the line (2D/mhdrhs.f90:282-288);
  // everywhere else fc2 < 0 and cf1 = 1.
  int fc2;
  double cf1, cf2;
  double aeb_c;   // 2,2,3,3,2,1,1,0 (mhdrhs.f90:235-247)
  int diff;       // 0 none, 1 viscosity (v=1..3), 2 resistivity (v=4..6)
  int jcomp;      // kZCurrent: 0,1,2
};

struct ZParams {
  int nxh, ny, nyl, yoff, ystride, nz, ncol;   // global ky of local row kyl: yoff + kyl * ystride (slabs: stride 1)
  const cplx* W2;        // [f][col][z]
  size_t fstride;        // ncol * nz
  const cplx* u_in;      // [v][col][kz]
  const cplx* u_old;     // kZMass: the state at the start of the stage (u_in then points at the updated one)
  cplx* u_out;
  cplx* fnl_rk;
  PeerTable V1;          // [g][kx][ky][zl] on the owner of z
  const cplx* tw;
  // 1-D tables (host-built in the reference's operation order, tables.cpp)
  const double* kxr;     // wave_number_x(1:nx/2+1)
  const double* kyr;     // wave_number_y
  const double* kze;     // wave_number_z * radius0 / radius
  const double* ksq_x;   // terms of k_square (see tables.cpp)
  const double* ksq_y;
  const double* ksq_z;
  const double* dax;     // dealias: option 1 -> squared radius terms, option 2 -> filters
  const double* day;
  const double* daz;
  double radius0, radius, cosa, sina, tau;
  double ksq_c1, ksq_c2, ksq_c3;   // corotating k_square coefficients (AEBmod.f90:103-110)
  int corot_k;           // if_AEB .and. if_corotating (derivative vectors)
  int corot_ksq;         // corotating k_square formula active
  int aeb;
  double cc, dd, dt_irk;
  int visc_imp, visc_exp, resis_imp, resis_exp, conserve_bg;
  double nu, eta;
  int dealias_option;
  int read_rk, write_rk;
  double scale;          // 1/nz
  // Columns the kernels visit: compact index c in [0, ncolc) -> kx = c / nkyl, and the local ky index
  // kyl = (r < nA ? a0 + r : b0 + r - nA), r = c % nkyl: the locally owned ky rows that survive the
  // dealiasing mask form at most two runs (ky <= kymax and ky >= ny - kymax).  Without pruning
  // ncolc = ncol, nkyl = nA = nyl, a0 = 0.
  int ncolc, nkyl, nA, a0, b0;
  // With the spherical mask (option 1) the surviving columns are those inside a circle in (kx, ky): colmap[c] is
  // then the memory column (kx * nyl + kyl) of compact index c and the formula above is not used.
  const int* colmap;
  const int* colkx;      // kx of colmap[c] (saves the division by nyl where a kernel walks many columns)
  // kzprune: modes the mask removes are neither loaded nor stored (they are zero in the state arrays, and
  // whatever fnl_rk holds there is never used): mask(kz) = dax + day + daz[kz] >= da_thresh (option 1) or a
  // non-zero per-axis flag (option 3).
  int kzprune;
  int mode2d;            // 2D tree: the line axis is the reference's y, d/dz = 0 (src_compressible/2D/mhdrhs.f90:272)
  // 2D tree with if_corotating (2D/mhdrhs.f90:282-288, 2D/AEBmod.f90:101-106): kx_eff = kx cos + ky sin and
  // ky_eff = (-kx sin + ky cos) R0/R both vary along the line (which carries ky).  The column constants kxe, kye below are
  // then their kx parts (the 3D formulas with the column's ky = 0); the ky parts enter through ZTask::fc/fc2, and where a
  // kernel needs the full vector per mode it uses kzr, the RAW line wave numbers.
  int corot2d;
  const double* kzr;
  double ksq_cross;      // corot2d: 1 - (R0/R)^2 of the cross term of k_square (2D/AEBmod.f90:104-105)
  int z_radial;          // 2D/mhdrhs.f90:278-280: kx is stretched too
  int bg_all_kz;         // 2D/mhdrhs.f90:372-374: if_conserve_background skips every mode with ix == 1
  double da_thresh;      // dealias option 1: smallest s with sqrt(s) > 1./3. (dealiasing.f90:94)
  double aeb_p;          // incompressible tree: 2*adiabatic_index, the expanding-box coefficient of the pressure row
  int tune;              // bit 0: L2-prefetch state/history lines, bit 1: L2-prefetch the G inputs
  ZTask task[12];
};

// Shared memory per column: one padded work line (the FFT exchanges) and one unpadded stash line
// holding the (i kz)-term between the two forward transforms, so that a thread never keeps more
// than one set of 8 points in registers (occupancy: the kernel is latency-bound otherwise).
template <int N, int CG>
struct ZTile {
  typedef Geom<N> G;
  static constexpr int PITCH = G::pitch(1);
  static constexpr int NTHREADS = CG * G::NT;
  static constexpr int COLSTRIDE = PITCH + N;
  static constexpr size_t SMEM = (size_t)CG * COLSTRIDE * sizeof(cplx);
  // resident CTAs per SM to compile for: what shared memory allows, but never below 80 registers
  static constexpr int BY_SMEM = (int)((227 * 1024) / (SMEM + 1024));
  static constexpr int WTHREADS = (NTHREADS + 31) / 32 * 32;   // registers and thread slots are handed out per warp
  static constexpr int BY_REGS = 65536 / (WTHREADS * 80);
  static constexpr int BY_THREADS = 2048 / WTHREADS;
  static constexpr int M0 = BY_SMEM < BY_REGS ? BY_SMEM : BY_REGS;
  static constexpr int M1 = M0 < BY_THREADS ? M0 : BY_THREADS;
  static constexpr int MINB = M1 < 1 ? 1 : (M1 > 32 ? 32 : M1);
};

// memory column (kx * nyl + kyl) of compact column index cc; -1 past the end
LAPS_D int z_column(const ZParams& P, int cc) {
  if (cc >= P.ncolc) return -1;
  if (P.colmap) return __ldg(P.colmap + cc);
  const int kr = cc % P.nkyl;
  return (cc / P.nkyl) * P.nyl + (kr < P.nA ? P.a0 + kr : P.b0 + kr - P.nA);
}

// the same, also returning kx of the column (kyl = column - kx * nyl)
LAPS_D int z_column_kx(const ZParams& P, int cc, int& kx) {
  kx = 0;
  if (cc >= P.ncolc) return -1;
  if (P.colmap) { kx = __ldg(P.colkx + cc); return __ldg(P.colmap + cc); }
  const int q = cc / P.nkyl, kr = cc - q * P.nkyl;
  kx = q;
  return q * P.nyl + (kr < P.nA ? P.a0 + kr : P.b0 + kr - P.nA);
}

// true where the dealiasing mask removes mode (kx, ky, kz): dxy = dax + day of the column (options 1, 3)
LAPS_D bool z_mode_dead(const ZParams& P, double dxy, int kz) {
  if (!P.kzprune) return false;
  const double dz = __ldg(P.daz + kz);
  return P.dealias_option == 1 ? (__dadd_rn(dxy, dz) >= P.da_thresh) : (dxy != 0.0 || dz != 0.0);
}

LAPS_D void prefetch_l2(const void* p) {
#ifndef LAPS_EMU_BUILD
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

// k_square(ix,iy,iz) exactly as update_ksquare / grid_initialize evaluate it: the terms that do not
// depend on iz are summed first there too, so  k_square = ksq_xy(ix,iy) + ksq_z(iz)  bit for bit.
LAPS_D double ksq_xy_of(const ZParams& P, double kxr, double kyr, double ksqx, double ksqy) {
  if (P.corot_ksq) {
    const double t1 = __dmul_rn(ksqx, P.ksq_c1);
    const double t2 = __dmul_rn(ksqy, P.ksq_c2);
    double t3 = __dmul_rn(kxr, kyr);
    t3 = __dmul_rn(t3, 2.0);
    t3 = __dmul_rn(t3, P.cosa);
    t3 = __dmul_rn(t3, P.sina);
    t3 = __dmul_rn(t3, P.ksq_c3);
    return __dadd_rn(__dadd_rn(t1, t2), t3);
  }
  return __dadd_rn(ksqx, ksqy);
}
LAPS_D double ksq_xy_eval(const ZParams& P, double kxr, double kyr, int kx, int ky) {
  return ksq_xy_of(P, kxr, kyr, __ldg(P.ksq_x + kx), __ldg(P.ksq_y + ky));
}

// 2D tree with if_corotating: the rotated wave vector of mode (kx, ky = line index) in the reference's operation order
// (2D/mhdrhs.f90:282-288)
LAPS_D void corot2d_k(const ZParams& P, double kxr, double kyl, double& kx_eff, double& ky_eff) {
  kx_eff = __dadd_rn(__dmul_rn(kxr, P.cosa), __dmul_rn(kyl, P.sina));
  ky_eff = __ddiv_rn(__dmul_rn(__dadd_rn(__dmul_rn(-kxr, P.sina), __dmul_rn(kyl, P.cosa)), P.radius0), P.radius);
}
// ... and the cross term kx*ky*2*cos*sin*(1 - (R0/R)^2) of its k_square (2D/AEBmod.f90:104-105)
LAPS_D double corot2d_cross(const ZParams& P, double kxr, double kyl) {
  double t = __dmul_rn(kxr, kyl);
  t = __dmul_rn(t, 2.0);
  t = __dmul_rn(t, P.cosa);
  t = __dmul_rn(t, P.sina);
  return __dmul_rn(t, P.ksq_cross);
}

// one column group (CG columns) of one task row
template <int N, int CG>
LAPS_D void spec_z_group(const ZParams& P, const ZTask& K, const int group, cplx* sm) {
  typedef Geom<N> G;
  typedef Fft<N, -1> FF;
  typedef Fft<N, +1> FI;
  typedef ZTile<N, CG> T;
  const int tid = threadIdx.x;
  const int l = tid / G::NT, u = tid % G::NT;
  const int colm = z_column(P, group * CG + l);
  const bool live = colm >= 0;
  const int col = live ? colm : 0;
  const int kx = col / P.nyl;
  const int ky = P.yoff + (col % P.nyl) * P.ystride;
  cplx* lineG = sm + l * T::COLSTRIDE;
  cplx* stash = lineG + T::PITCH;          // thread-private slots e*NT + u
  const size_t coff = (size_t)col * N;

  // derivative vectors (imaginary parts), mhdrhs.f90:191-204
  const double kxr = __ldg(P.kxr + kx), kyr = __ldg(P.kyr + ky);
  double kxe = kxr, kye = __ddiv_rn(__dmul_rn(kyr, P.radius0), P.radius);
  if (P.z_radial) kxe = __ddiv_rn(__dmul_rn(kxr, P.radius0), P.radius);
  if (P.corot_k) {
    kxe = __dadd_rn(__dmul_rn(kxr, P.cosa), __dmul_rn(kyr, P.sina));
    kye = __ddiv_rn(__dmul_rn(__dadd_rn(__dmul_rn(-kxr, P.sina), __dmul_rn(kyr, P.cosa)), P.radius0), P.radius);
  }

  cplx r[8];
  if (K.kind == kZRhs || K.kind == kZForwardOnly) {
    const bool hasC = K.fc >= 0;
    const size_t voff = (size_t)K.v * P.fstride + coff;
    if (K.kind == kZRhs && live) {
      // Lines needed by the later phases of this CTA are pulled into L2 now (no registers held), so
      // that their DRAM latency overlaps the first transform instead of being exposed phase by phase.
      const size_t po = coff + (size_t)u * (N / G::NT);
      if (P.tune & 1) {
        prefetch_l2(P.u_in + (size_t)K.v * P.fstride + po);
        if (P.read_rk) prefetch_l2(P.fnl_rk + (size_t)K.v * P.fstride + po);
      }
      if ((P.tune & 2) && hasC) {
        if (K.fa >= 0) prefetch_l2(P.W2 + (size_t)K.fa * P.fstride + po);
        if (K.fb >= 0) prefetch_l2(P.W2 + (size_t)K.fb * P.fstride + po);
        if (K.fx >= 0) prefetch_l2(P.W2 + (size_t)K.fx * P.fstride + po);
      }
    }
    // ---------------- forward z of the (i kz) term, kept in the stash ----------------
    if (hasC) {
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) r[e] = mk(0.0, 0.0);
      if (live) {
        const cplx* s = P.W2 + (size_t)K.fc * P.fstride + coff;
        LAPS_UNROLL
        for (int e = 0; e < 8; ++e) r[e] = s[u + e * G::NT];
        if (K.fc2 >= 0) {   // 2D tree with if_corotating: two fields share the (i k_line) factor
          const cplx* s2 = P.W2 + (size_t)K.fc2 * P.fstride + coff;
          LAPS_UNROLL
          for (int e = 0; e < 8; ++e) r[e] = cadd(cscale(r[e], K.cf1), cscale(s2[u + e * G::NT], K.cf2));
        } else if (K.cf1 != 1.0) {
          LAPS_UNROLL
          for (int e = 0; e < 8; ++e) r[e] = cscale(r[e], K.cf1);
        }
      }
      FF::first(r, u, lineG, P.tw);
      FF::finish(r, u, lineG, P.tw);
      const double cs = K.sc * P.scale;
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) stash[e * G::NT + u] = cmul_i(r[e], cs * __ldg(P.kze + FF::kout(u, e)));
      __syncthreads();  // every last-stage read of the work line is done before it is refilled
    }
    // ---------------- forward z of G ----------------
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) r[e] = mk(0.0, 0.0);
    if (live) {
      if (K.fa >= 0) {
        const cplx* s = P.W2 + (size_t)K.fa * P.fstride + coff;
        const double c = (K.kind == kZRhs) ? K.ca * kxe : 0.0;
        LAPS_UNROLL
        for (int e = 0; e < 8; ++e) {
          const cplx a = s[u + e * G::NT];
          r[e] = (K.kind == kZRhs) ? cmul_i(a, c) : a;
        }
      }
      if (K.fb >= 0) {
        const cplx* s = P.W2 + (size_t)K.fb * P.fstride + coff;
        const double c = K.cb * kye;
        LAPS_UNROLL
        for (int e = 0; e < 8; ++e) r[e] = cadd(r[e], cmul_i(s[u + e * G::NT], c));
      }
      if (K.fx >= 0) {
        const cplx* s = P.W2 + (size_t)K.fx * P.fstride + coff;
        LAPS_UNROLL
        for (int e = 0; e < 8; ++e) r[e] = cadd(r[e], cscale(s[u + e * G::NT], K.cx));
      }
    }
    FF::first(r, u, lineG, P.tw);
    FF::finish(r, u, lineG, P.tw);

    // ---------------- spectral update on the 8 modes this thread holds ----------------
    if (K.kind == kZForwardOnly) {
      if (live) {
        LAPS_UNROLL
        for (int e = 0; e < 8; ++e) P.u_out[voff + FF::kout(u, e)] = cscale(r[e], P.scale);
      }
      return;
    }
    // Task-uniform coefficients first, so that the per-mode code below is straight-line arithmetic
    // the compiler can interleave across the 8 modes (a zero coefficient switches a term off).
    // Divisions by a common divisor become multiplications by its reciprocal, and the mask test
    // sqrt(s) > 1/3 becomes s >= P.da_thresh with the host-computed smallest s that passes it
    // (bit-identical decisions, see upload_tables).
    const double sgs = K.sg * P.scale;
    const double ca = (P.aeb && K.aeb_c != 0.0) ? K.aeb_c / P.tau : 0.0;                       // mhdrhs.f90:235-247
    const double ce = (K.diff == 1 && P.visc_exp) ? P.nu : ((K.diff == 2 && P.resis_exp) ? P.eta : 0.0);   // :253-275
    const double ci = (K.diff == 1 && P.visc_imp) ? P.nu : ((K.diff == 2 && P.resis_imp) ? P.eta : 0.0);   // rktmod.f90:47-60
    const bool need_ksq = (ce != 0.0) || (ci != 0.0);
    const double ksq_xy = need_ksq ? ksq_xy_eval(P, kxr, kyr, kx, ky) : 0.0;
    const bool keep_bg = K.diff == 2 && P.conserve_bg && kx == 0;   // "ix==1 .and. iz==1" skip of mhdrhs.f90:262-270
    const double dxy = (P.dealias_option == 1 || P.dealias_option == 3) ? __dadd_rn(__ldg(P.dax + kx), __ldg(P.day + ky))
                                               : ((P.dealias_option == 2) ? __ldg(P.dax + kx) : 0.0);
    const double dfy = (P.dealias_option == 2) ? __ldg(P.day + ky) : 0.0;
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) {
      const int kz = FF::kout(u, e);
      cplx fnl = cscale(r[e], sgs);
      if (hasC) fnl = cadd(fnl, stash[e * G::NT + u]);
      const bool keep = live && !z_mode_dead(P, dxy, kz);   // masked modes: zero in memory, not touched
      const cplx uo = keep ? P.u_in[voff + kz] : mk(0.0, 0.0);
      fnl.x -= ca * uo.x;
      fnl.y -= ca * uo.y;
      double ksq = 0.0;
      if (need_ksq) {
        ksq = __dadd_rn(ksq_xy, __ldg(P.ksq_z + kz));
        if (P.corot2d && P.corot_ksq) ksq = __dadd_rn(ksq, corot2d_cross(P, kxr, __ldg(P.kzr + kz)));
        const double cee = (keep_bg && (P.bg_all_kz || kz == 0)) ? 0.0 : ce;
        fnl.x -= (cee * uo.x) * ksq;
        fnl.y -= (cee * uo.y) * ksq;
      }
      // rkt (rktmod.f90:40-42): u = cc*fnl + dd*fnl_rk + u ; fnl_rk = fnl
      cplx un;
      if (P.read_rk) {
        const cplx fr = keep ? P.fnl_rk[voff + kz] : mk(0.0, 0.0);
        un = mk((P.cc * fnl.x + P.dd * fr.x) + uo.x, (P.cc * fnl.y + P.dd * fr.y) + uo.y);
      } else {
        un = mk(P.cc * fnl.x + uo.x, P.cc * fnl.y + uo.y);
      }
      if (P.write_rk && keep) P.fnl_rk[voff + kz] = fnl;
      if (need_ksq) {  // implicit diffusion (rktmod.f90:47-60); ci == 0 gives exactly 1
        const double inv = __drcp_rn(__dadd_rn(__dmul_rn(__dmul_rn(P.dt_irk, ksq), ci), 1.0));
        un.x *= inv;
        un.y *= inv;
      }
      // dealias (dealiasing.f90:87-110)
      if (P.dealias_option == 1) {
        if (__dadd_rn(dxy, __ldg(P.daz + kz)) >= P.da_thresh) un = mk(0.0, 0.0);
      } else if (P.dealias_option == 2) {
        const double fz = __ldg(P.daz + kz);
        un = mk(__dmul_rn(__dmul_rn(__dmul_rn(un.x, dxy), dfy), fz), __dmul_rn(__dmul_rn(__dmul_rn(un.y, dxy), dfy), fz));
      } else if (P.dealias_option == 3) {   // square truncation (2D/dealiasing.f90:102-117): per-axis flags
        if (dxy != 0.0 || __ldg(P.daz + kz) != 0.0) un = mk(0.0, 0.0);
      }
      if (keep) P.u_out[voff + kz] = un;
      r[e] = un;
    }
    if (K.kind == kZForwardOnly || K.gout < 0) return;
    // re-shape the register contents into the stage-0 input pattern of the inverse transform
    __syncthreads();  // all last-stage reads of lineG are done
    if constexpr (G::RLAST != 8) {
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) lineG[G::pad(FF::kout(u, e))] = r[e];
      __syncthreads();
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) r[e] = lineG[FI::in_pos(u, G::pad(u), e)];
      __syncthreads();
    }
  } else if (K.kind == kZInverseOnly) {
    const size_t voff = (size_t)K.v * P.fstride + coff;
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) r[e] = live ? P.u_in[voff + u + e * G::NT] : mk(0.0, 0.0);
  } else if (K.kind == kZMass) {
    // continuity row (mhdrhs.f90:207-209,235-237), rkt, dealias — in the INPUT order of the inverse transform
    const cplx* M = P.u_old + P.fstride + coff;
    const size_t voff = coff;   // v = 0
    const double ca = P.aeb ? K.aeb_c / P.tau : 0.0;
    const double dxy = (P.dealias_option == 1 || P.dealias_option == 3) ? __dadd_rn(__ldg(P.dax + kx), __ldg(P.day + ky))
                                               : ((P.dealias_option == 2) ? __ldg(P.dax + kx) : 0.0);
    const double dfy = (P.dealias_option == 2) ? __ldg(P.day + ky) : 0.0;
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) {
      const int kz = u + e * G::NT;
      const double kzz = __ldg(P.kze + kz);
      double rx = kxe, ry = P.mode2d ? kzz : kye;
      const double rz = P.mode2d ? 0.0 : kzz;
      if (P.corot2d) corot2d_k(P, kxr, __ldg(P.kzr + kz), rx, ry);
      const bool keep = live && !z_mode_dead(P, dxy, kz);   // masked modes: zero in memory, not touched
      const cplx m1 = keep ? M[kz] : mk(0.0, 0.0);
      const cplx m2 = keep ? M[P.fstride + kz] : mk(0.0, 0.0);
      const cplx m3 = keep ? M[2 * P.fstride + kz] : mk(0.0, 0.0);
      const cplx sum = cadd(cadd(cmul_i(m1, rx), cmul_i(m2, ry)), cmul_i(m3, rz));
      cplx fnl = mk(-sum.x, -sum.y);
      const cplx uo = keep ? P.u_old[voff + kz] : mk(0.0, 0.0);
      fnl.x -= ca * uo.x;
      fnl.y -= ca * uo.y;
      cplx un;
      if (P.read_rk) {
        const cplx fr = keep ? P.fnl_rk[voff + kz] : mk(0.0, 0.0);
        un = mk((P.cc * fnl.x + P.dd * fr.x) + uo.x, (P.cc * fnl.y + P.dd * fr.y) + uo.y);
      } else {
        un = mk(P.cc * fnl.x + uo.x, P.cc * fnl.y + uo.y);
      }
      if (P.write_rk && keep) P.fnl_rk[voff + kz] = fnl;
      if (P.dealias_option == 1) {
        if (__dadd_rn(dxy, __ldg(P.daz + kz)) >= P.da_thresh) un = mk(0.0, 0.0);
      } else if (P.dealias_option == 2) {
        const double fz = __ldg(P.daz + kz);
        un = mk(__dmul_rn(__dmul_rn(__dmul_rn(un.x, dxy), dfy), fz), __dmul_rn(__dmul_rn(__dmul_rn(un.y, dxy), dfy), fz));
      } else if (P.dealias_option == 3) {
        if (dxy != 0.0 || __ldg(P.daz + kz) != 0.0) un = mk(0.0, 0.0);
      }
      if (keep) P.u_out[voff + kz] = un;
      r[e] = un;
    }
  } else if (K.kind == kZGrad) {
    const cplx* U = P.u_in + (size_t)K.v * P.fstride + coff;
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) {
      const int kz = u + e * G::NT;
      // 2D trees: the line axis carries the reference's ky and kz = 0 (src_incompressible/2D/mhdrhs.f90:396)
      const double kzz = __ldg(P.kze + kz);
      double rx = kxe, ry = P.mode2d ? kzz : kye;
      if (P.corot2d) corot2d_k(P, kxr, __ldg(P.kzr + kz), rx, ry);
      const double ka = K.jcomp == 0 ? rx : (K.jcomp == 1 ? ry : (P.mode2d ? 0.0 : kzz));
      const cplx t = cmul_i(live ? U[kz] : mk(0.0, 0.0), ka);
      r[e] = mk(__ddiv_rn(t.x, K.cx), __ddiv_rn(t.y, K.cx));
    }
  } else if (K.kind == kZDiv) {
    const cplx* U = P.u_in + (size_t)K.v * P.fstride + coff;
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) {
      const int kz = u + e * G::NT;
      const cplx a = live ? U[kz] : mk(0.0, 0.0);
      const cplx b = live ? U[P.fstride + kz] : mk(0.0, 0.0);
      const cplx c = live ? U[2 * P.fstride + kz] : mk(0.0, 0.0);
      const double kzz = __ldg(P.kze + kz);
      double rx = kxe, ry = P.mode2d ? kzz : kye;
      if (P.corot2d) corot2d_k(P, kxr, __ldg(P.kzr + kz), rx, ry);
      const cplx t = cadd(cadd(cmul_i(a, rx), cmul_i(b, ry)), cmul_i(c, P.mode2d ? 0.0 : kzz));
      r[e] = mk(__ddiv_rn(t.x, K.cx), __ddiv_rn(t.y, K.cx));
    }
  } else {  // kZCurrent: J^ = i k x B^ (mhdrhs.f90:329-336) from the updated state
    const int j = K.jcomp;
    const double dxy_col = P.kzprune ? __dadd_rn(__ldg(P.dax + kx), __ldg(P.day + ky)) : 0.0;
    const cplx* B1 = P.u_in + (size_t)(4 + (j + 1) % 3) * P.fstride + coff;  // B_{j+1}
    const cplx* B2 = P.u_in + (size_t)(4 + (j + 2) % 3) * P.fstride + coff;  // B_{j+2}
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) {
      const int kz = u + e * G::NT;
      const double kzz = __ldg(P.kze + kz);
      // k_{j+1} B_{j+2} - k_{j+2} B_{j+1}   with (k0,k1,k2) = (kx,ky,kz) of the REFERENCE axes; in the
      // 2D tree the line axis carries the reference's ky and kz = 0 (2D/mhdrhs.f90:412-440)
      double rx = kxe, ry = P.mode2d ? kzz : kye;
      const double rz = P.mode2d ? 0.0 : kzz;
      if (P.corot2d) corot2d_k(P, kxr, __ldg(P.kzr + kz), rx, ry);
      const double k1 = (j == 0) ? ry : (j == 1 ? rz : rx);
      const double k2 = (j == 0) ? rz : (j == 1 ? rx : ry);
      const bool keep = live && !z_mode_dead(P, dxy_col, kz);   // masked modes of the state are zero
      const cplx b1 = keep ? B1[kz] : mk(0.0, 0.0);
      const cplx b2 = keep ? B2[kz] : mk(0.0, 0.0);
      r[e] = csub(cmul_i(b2, k1), cmul_i(b1, k2));
    }
  }

  // ---------------- inverse z, stored on the owner of each z (transpose_zy fused) ----------------
  FI::first(r, u, lineG, P.tw);
  FI::finish(r, u, lineG, P.tw);
  if (live) {
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) {
      const int z = FI::kout(u, e);
      const int p = P.V1.owner(z);
      cplx* dst = P.V1.base[p] + (((size_t)K.gout * P.nxh + kx) * P.ny + ky) * P.V1.len[p] + (z - P.V1.off[p]);
      *dst = r[e];
    }
  }
}

// grid.x = column groups (or fewer: grid-stride loop, see k_fwd_y), grid.y = task rows
template <int N, int CG>
__global__ void __launch_bounds__(ZTile<N, CG>::NTHREADS, ZTile<N, CG>::MINB)
k_spec_z(const ZParams P, const int ngroups) {
  LAPS_DYN_SMEM(cplx, sm);
  const ZTask& K = P.task[blockIdx.y];
  for (int group = blockIdx.x; group < ngroups; group += gridDim.x) {
    spec_z_group<N, CG>(P, K, group, sm);
    if (group + (int)gridDim.x < ngroups) __syncthreads();   // the next group's first stage refills the lines
  }
}

}  // namespace laps
