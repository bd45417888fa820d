// calc_flux (mhdrhs.f90:21-124) fused into the forward x pass (mhdrhs.f90:143-149, fftw.f90:58-64):
// the 18 (+1) real flux fields of a stage are never written to memory.  k_flux + k_fwd_x move
// (8 + 3 + nf) R + nf R + nf C per stage; this kernel moves (8 + 3) R + nf C.
//
// One CTA owns one pair of adjacent real x lines (zl, y, y+1) — one complex transform per flux — for
// ALL fluxes of the stage:
//   phase 0  the 8 conserved fields of the two lines are staged in shared memory (coalesced), the
//            primitive velocity and pressure are formed once per point in the reference's expression
//            order (mhdrhs.f90:285-293) and parked next to them;
//   phase 1  GR thread groups of N/8 threads each take one flux per round: form it at the 16 points
//            the group's transform needs (8 per line), transform, split the two half spectra and
//            store the (y, y+1) pair of every kx as one 32-byte access into W1 [f][kx][zl][y].
//            A group synchronises only with itself (named barriers), so groups overlap each other's
//            transform, split and store phases.
// Adjacent CTAs (y fastest) complete each 128-byte line of W1 within microseconds, so the L2 merges
// the sectors before they are written back.
#pragma once
#include "fft_passes.cuh"
#include "pointwise.cuh"

namespace laps {

struct FusedFluxParams {
  const double* uu;     // [8][npts]
  const double* J;      // [3][npts] (Hall) or null
  cplx* W1;             // [nf][nxh][nzl][ny]
  size_t npts;
  int nzl, ny, nkx;
  const cplx* tw;
  double scale;         // 1/nx
  int hall, aeb;
  double gamma, di, tau;
  int nflux;            // fluxes to form and transform
  int id[19];           // flux id (0-based F1..F18, 18 = expanding-box source) of each W1 slot, in slot order
};

template <int N, int GR>
struct FTile {
  typedef Geom<N> G;
  static constexpr int PITCH = G::pitch(1);
  static constexpr int NTHREADS = GR * G::NT;
  static constexpr int NIN = 15;   // rho | di/rho, mx, my, mz, Bx, By, Bz, e, ux, uy, uz, p, Jx, Jy, Jz
  static constexpr size_t SMEM_IN = (size_t)NIN * 2 * N * sizeof(double);
  static constexpr size_t SMEM = SMEM_IN + (size_t)GR * PITCH * sizeof(cplx);
};

template <int N, int GR>
__global__ void __launch_bounds__(FTile<N, GR>::NTHREADS, 1)
k_flux_fwd_x(const FusedFluxParams P) {
  typedef Geom<N> G;
  typedef Fft<N, -1> F;
  typedef FTile<N, GR> T;
  LAPS_DYN_SMEM(double, smd);
  double* in = smd;                                           // in[v * 2N + i], i = line * N + x
  cplx* work = reinterpret_cast<cplx*>(smd + T::NIN * 2 * N);
  const int tid = threadIdx.x;
  const int ypairs = P.ny / 2;
  const int zl = blockIdx.x / ypairs;
  const int y0 = (blockIdx.x % ypairs) * 2;
  const size_t row = ((size_t)zl * P.ny + y0) * N;            // first element of the line pair in a real field
  constexpr int L2N = 2 * N;

  // ---------------- phase 0: stage the conserved fields, then the primitives ----------------
  for (int i = tid * 2; i < 8 * L2N; i += T::NTHREADS * 2) {  // 16-byte accesses, coalesced
    const int v = i / L2N, o = i % L2N;
    const double2 a = *reinterpret_cast<const double2*>(P.uu + (size_t)v * P.npts + row + o);
    *reinterpret_cast<double2*>(in + v * L2N + o) = a;
  }
  if (P.hall)
    for (int i = tid * 2; i < 3 * L2N; i += T::NTHREADS * 2) {
      const int v = i / L2N, o = i % L2N;
      const double2 a = *reinterpret_cast<const double2*>(P.J + (size_t)v * P.npts + row + o);
      *reinterpret_cast<double2*>(in + (12 + v) * L2N + o) = a;
    }
  __syncthreads();
  const double gm1 = P.gamma - 1.0;
  for (int i = tid; i < L2N; i += T::NTHREADS) {
    const double rho = in[i], mx = in[L2N + i], my = in[2 * L2N + i], mz = in[3 * L2N + i];
    const Prim q = prim_of(rho, mx, my, mz, in[4 * L2N + i], in[5 * L2N + i], in[6 * L2N + i], in[7 * L2N + i], gm1);
    in[8 * L2N + i] = q.ux; in[9 * L2N + i] = q.uy; in[10 * L2N + i] = q.uz; in[11 * L2N + i] = q.p;
    if (P.hall) in[i] = P.di / rho;                           // the only later use of rho (mhdrhs.f90:108-121)
  }
  __syncthreads();

  // ---------------- phase 1: one flux per group per round ----------------
  const int g = tid / G::NT, u = tid % G::NT;
  cplx* line = work + g * T::PITCH;
  const int nxh = N / 2 + 1;
  const int rounds = (P.nflux + GR - 1) / GR;
  for (int rd = 0; rd < rounds; ++rd) {
    const int f = rd * GR + g;
    const bool live = f < P.nflux;
    const int j = live ? P.id[f] : 0;
    cplx r[8];
    if (!live) {
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) r[e] = mk(0.0, 0.0);
    } else if (j < 3) {                       // mass flux rho u (mhdrhs.f90:58-60)
      const double* m = in + (1 + j) * L2N;
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) { const int x = u + e * G::NT; r[e] = mk(m[x], m[N + x]); }
    } else if (j < 12) {                      // momentum tensor rho u_c u_r - B_r B_c + ptot delta_rc (:63-79)
      const int c = (j - 3) / 3, rr = (j - 3) % 3;
      const double* m = in + (1 + rr) * L2N;
      const double* uc = in + (8 + c) * L2N;
      const double* br = in + (4 + rr) * L2N;
      const double* bc = in + (4 + c) * L2N;
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) {
        double o[2];
        LAPS_UNROLL
        for (int h = 0; h < 2; ++h) {
          const int x = h * N + u + e * G::NT;
          double v = m[x] * uc[x] - br[x] * bc[x];
          if (c == rr) {
            const double Bx = in[4 * L2N + x], By = in[5 * L2N + x], Bz = in[6 * L2N + x];
            v = v + (in[11 * L2N + x] + 0.5 * (Bx * Bx + By * By + Bz * Bz));
          }
          o[h] = v;
        }
        r[e] = mk(o[0], o[1]);
      }
    } else if (j < 15) {                      // E = -u x B (+ Hall) (:82-84, 108-121)
      const int c = j - 12, c1 = (c + 1) % 3, c2 = (c + 2) % 3;
      const double* u1 = in + (8 + c1) * L2N; const double* u2 = in + (8 + c2) * L2N;
      const double* b1 = in + (4 + c1) * L2N; const double* b2 = in + (4 + c2) * L2N;
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) {
        double o[2];
        LAPS_UNROLL
        for (int h = 0; h < 2; ++h) {
          const int x = h * N + u + e * G::NT;
          double v = u2[x] * b1[x] - u1[x] * b2[x];
          if (P.hall) {
            const double j1 = in[(12 + c1) * L2N + x], j2 = in[(12 + c2) * L2N + x];
            v = v + in[x] * (j1 * b2[x] - j2 * b1[x]);
          }
          o[h] = v;
        }
        r[e] = mk(o[0], o[1]);
      }
    } else if (j < 18) {                      // energy flux (e + ptot) u_c - (u.B) B_c (:87-89)
      const int c = j - 15;
      const double* uc = in + (8 + c) * L2N;
      const double* bc = in + (4 + c) * L2N;
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) {
        double o[2];
        LAPS_UNROLL
        for (int h = 0; h < 2; ++h) {
          const int x = h * N + u + e * G::NT;
          const double Bx = in[4 * L2N + x], By = in[5 * L2N + x], Bz = in[6 * L2N + x];
          const double ptot = in[11 * L2N + x] + 0.5 * (Bx * Bx + By * By + Bz * Bz);
          const double udotb = in[8 * L2N + x] * Bx + in[9 * L2N + x] * By + in[10 * L2N + x] * Bz;
          o[h] = (in[7 * L2N + x] + ptot) * uc[x] - udotb * bc[x];
        }
        r[e] = mk(o[0], o[1]);
      }
    } else {                                  // expanding-box energy source (:93-103)
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) {
        double o[2];
        LAPS_UNROLL
        for (int h = 0; h < 2; ++h) {
          const int x = h * N + u + e * G::NT;
          const double Bx = in[4 * L2N + x], By = in[5 * L2N + x], Bz = in[6 * L2N + x];
          o[h] = -2 * P.gamma / gm1 * in[11 * L2N + x] / P.tau - (2.0 * Bx * Bx + By * By + Bz * Bz) / P.tau -
                 (in[L2N + x] * in[8 * L2N + x] + 2 * in[2 * L2N + x] * in[9 * L2N + x] + 2 * in[3 * L2N + x] * in[10 * L2N + x]) / P.tau;
        }
        r[e] = mk(o[0], o[1]);
      }
    }
    // from here on only the group's own threads touch its work line: group barriers, the groups drift freely
    F::first(r, u, line, P.tw);
    F::template finish_g<GR>(r, u, line, P.tw, 1 + g);
    group_barrier<G::NT, GR>(1 + g);  // everyone has consumed its last-stage slots
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) line[G::pad(F::kout(u, e))] = r[e];
    group_barrier<G::NT, GR>(1 + g);
    if (live) {       // split Z = A + iB into the half spectra of the two lines; (a, b) = W1[..][y0], W1[..][y0+1]
      const double hs = 0.5 * P.scale;
      for (int k = u; k < P.nkx; k += G::NT) {
        const cplx zk = line[G::pad(k)];
        const cplx zn = line[G::pad((N - k) & (N - 1))];
        const cplx a = mk((zk.x + zn.x) * hs, (zk.y - zn.y) * hs);
        const cplx b = mk((zk.y + zn.y) * hs, (zn.x - zk.x) * hs);
        st256(P.W1 + (((size_t)f * nxh + k) * P.nzl + zl) * P.ny + y0, a, b);
      }
    }
    group_barrier<G::NT, GR>(1 + g);  // the work line is refilled by the next round
  }
}

}  // namespace laps
