// The RK-stage z pass as a persistent, software-pipelined kernel (the hot kernel of the step).
//
// Same arithmetic as the kZRhs branch of k_spec_z (spectral_z.cuh) — forward z transforms of the
// flux combinations, calc_rhs (mhdrhs.f90:174-279), rkt (rktmod.f90:34-62), dealias
// (dealiasing.f90:70-112), inverse z transform stored on the owner of each z (transpose_zy,
// parallel.f90:300-324) — but every input line (up to four flux spectra, the state, the RK history)
// reaches the CTA through asynchronous global->shared copies into THREAD-PRIVATE landing slots,
// issued one phase ahead, so that no register is held across a DRAM round trip and the copies of
// one phase overlap the transform of the previous one:
//
//   item i:   [fc -> S was issued during item i-1]      issue fa -> Q0, fb -> Q1
//             wait fc : FFT(fc), (i kz).result -> S (in place, same slots)
//             wait fa : r  = (i kx ca) Q0                 issue fx -> Q0
//             wait fb : r += (i ky cb) Q1                 issue u  -> Q1
//             wait fx : r += cx Q0                        issue rk -> Q0
//             FFT(r)
//             wait u, rk : spectral update (S, Q1, Q0)    issue next item's fc -> S
//             inverse FFT, stores
//
// Shared memory per column: padded work line + S + Q0 + Q1.  CTAs are persistent (one wave,
// items strided by the grid) so that the fc prefetch crosses item boundaries.
#pragma once
#include "spectral_z.cuh"

namespace laps {

// NQ = landing lines next to S: 2 (Q0, Q1: every input line is in flight one phase ahead) or 1 (Q0 only: the second
// flux line of a combination, the expanding-box source and the RK history are fetched when they are needed — from L2, where
// the look-ahead hints of the previous item have put them — in exchange for a third more resident columns per SM).
template <int N, int CG, int NQ = 2>
struct RTile {
  typedef Geom<N> G;
  static constexpr int PITCH = G::pitch(1);
  static constexpr int NTHREADS = CG * G::NT;
  static constexpr int COLSTRIDE = PITCH + (1 + NQ) * N;
  static constexpr size_t SMEM = (size_t)CG * COLSTRIDE * sizeof(cplx);
  static constexpr int BY_SMEM = (int)((227 * 1024) / (SMEM + 1024));
  static constexpr int BY_REGS = 65536 / (NTHREADS * (NQ == 2 ? 80 : 112));
  static constexpr int BY_THREADS = 2048 / NTHREADS;
  static constexpr int M0 = BY_SMEM < BY_REGS ? BY_SMEM : BY_REGS;
  static constexpr int M1 = M0 < BY_THREADS ? M0 : BY_THREADS;
  static constexpr int MINB = M1 < 1 ? 1 : (M1 > 32 ? 32 : M1);
};

template <int N, int CG, int NQ = 2>
__global__ void __launch_bounds__(RTile<N, CG, NQ>::NTHREADS, RTile<N, CG, NQ>::MINB)
k_rhs_z(const ZParams P, const int ntasks, const int ngroups) {
  typedef Geom<N> G;
  typedef Fft<N, -1> FF;
  typedef Fft<N, +1> FI;
  typedef RTile<N, CG, NQ> T;
  LAPS_DYN_SMEM(cplx, sm);
  const int tid = threadIdx.x;
  const int l = tid / G::NT, u = tid % G::NT;
  cplx* W = sm + l * T::COLSTRIDE;      // padded work line
  cplx* S = W + T::PITCH;               // fc landing, then the (i kz) term
  cplx* Q0 = S + N;
  cplx* Q1 = NQ == 2 ? Q0 + N : Q0;   // NQ == 1: one landing line, used in turn
  // thread-constant base twiddles of every stage (forward; the inverse uses the conjugates)
  const cplx tw0 = FF::template tw_stage<0>(P.tw, u);
  const cplx tw1 = G::NSTAGE >= 3 ? FF::template tw_stage<1>(P.tw, u) : mk(1.0, 0.0);
  const cplx tw2 = G::NSTAGE >= 4 ? FF::template tw_stage<2>(P.tw, u) : mk(1.0, 0.0);

  // slot e of this thread <- element idx(e) of a line (8 copies, one commit group; an absent line
  // still commits an (empty) group so that the group count is the same in every thread)
  auto land_in = [&](cplx* buf, const cplx* line, bool on) {
    if (on) {
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) cp_async16(buf + e * G::NT + u, line + u + e * G::NT);
    }
    cp_async_commit();
  };
  // indexed by the OUTPUT order of the forward FFT; modes in `dead` (bit e) are masked: zero in memory, not fetched
  auto land_out = [&](cplx* buf, const cplx* line, bool on, unsigned dead) {
    if (on) {
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e)
        if (!((dead >> e) & 1u)) cp_async16(buf + e * G::NT + u, line + FF::kout(u, e));
    }
    cp_async_commit();
  };

  // Items = (column group, task), task fastest.  The walk over them carries (task, group) instead of dividing the item
  // number, and the column of the NEXT item (memory column + kx: two table loads whose results are addresses) is fetched
  // at the top of the current one, so that no item starts with a chain of dependent global loads.
  const int st_t = (int)(gridDim.x % (unsigned)ntasks), st_g = (int)(gridDim.x / (unsigned)ntasks);
  int task = (int)(blockIdx.x % (unsigned)ntasks), group = (int)(blockIdx.x / (unsigned)ntasks);
  int colm = -1, kxc = 0;
  if (group < ngroups) {  // prologue: the first item's column and its fc
    colm = z_column_kx(P, group * CG + l, kxc);
    const ZTask& K = P.task[task];
    land_in(S, P.W2 + (size_t)(K.fc >= 0 ? K.fc : 0) * P.fstride + (size_t)(colm < 0 ? 0 : colm) * N, K.fc >= 0 && colm >= 0);
  }
  while (group < ngroups) {
    const ZTask& K = P.task[task];
    int ntask = task + st_t, ngroup = group + st_g;
    if (ntask >= ntasks) { ntask -= ntasks; ++ngroup; }
    const bool more = ngroup < ngroups;
    int coln = -1, kxn = 0;
    if (more) coln = z_column_kx(P, ngroup * CG + l, kxn);   // consumed after the first transform and at the end of the item
    const bool live = colm >= 0;
    const int col = live ? colm : 0;
    const int kx = live ? kxc : 0;
    const int ky = P.yoff + (col - kx * P.nyl) * P.ystride;
    const size_t coff = (size_t)col * N;
    const size_t voff = (size_t)K.v * P.fstride + coff;
    const bool hasC = K.fc >= 0;

    land_in(Q0, P.W2 + (size_t)(K.fa >= 0 ? K.fa : 0) * P.fstride + coff, live && K.fa >= 0);
    if constexpr (NQ == 2) land_in(Q1, P.W2 + (size_t)(K.fb >= 0 ? K.fb : 0) * P.fstride + coff, live && K.fb >= 0);

    // per-column table entries: issued here, first used after the transform below
    const double kxr = __ldg(P.kxr + kx), kyr = __ldg(P.kyr + ky);
    const double ksqx = __ldg(P.ksq_x + kx), ksqy = __ldg(P.ksq_y + ky);
    const double dax = P.dealias_option ? __ldg(P.dax + kx) : 0.0, day = P.dealias_option ? __ldg(P.day + ky) : 0.0;

    cplx r[8];
    // ---------------- forward z of the (i kz) term, result kept in S ----------------
    if constexpr (NQ == 2) cp_async_wait<2>();   // fc has landed (fa, fb may still be in flight)
    else cp_async_wait<1>();
    if (hasC) {
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) r[e] = live ? S[e * G::NT + u] : mk(0.0, 0.0);
      FF::first_w(r, u, W, tw0);
      FF::finish_w(r, u, W, tw1, tw2);
      const double cs = K.sc * P.scale;
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) S[e * G::NT + u] = cmul_i(r[e], cs * __ldg(P.kze + FF::kout(u, e)));
      __syncthreads();  // every last-stage read of the work line is done before it is refilled
    }
    // derivative vectors (imaginary parts), mhdrhs.f90:191-204
    double kxe = kxr, kye = __ddiv_rn(__dmul_rn(kyr, P.radius0), P.radius);
    if (P.z_radial) kxe = __ddiv_rn(__dmul_rn(kxr, P.radius0), P.radius);
    if (P.corot_k) {
      kxe = __dadd_rn(__dmul_rn(kxr, P.cosa), __dmul_rn(kyr, P.sina));
      kye = __ddiv_rn(__dmul_rn(__dadd_rn(__dmul_rn(-kxr, P.sina), __dmul_rn(kyr, P.cosa)), P.radius0), P.radius);
    }
    const double dxy = (P.dealias_option == 1 || P.dealias_option == 3) ? __dadd_rn(dax, day) : dax;
    unsigned dead = 0;   // bit e: the mask removes mode kout(u, e) of this column
    if (P.kzprune) {
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) dead |= z_mode_dead(P, dxy, FF::kout(u, e)) ? (1u << e) : 0u;
    }
    if ((P.tune & 4) && more && coln >= 0) {
      // one item ahead: pull every line of the NEXT item into L2 (no registers, no shared memory),
      // so that its asynchronous copies run at L2 latency instead of DRAM latency
      const ZTask& Kn = P.task[ntask];
      const size_t po = (size_t)coln * N + (size_t)u * (N / G::NT);
      if (Kn.fa >= 0) prefetch_l2(P.W2 + (size_t)Kn.fa * P.fstride + po);
      if (Kn.fb >= 0) prefetch_l2(P.W2 + (size_t)Kn.fb * P.fstride + po);
      if (Kn.fx >= 0) prefetch_l2(P.W2 + (size_t)Kn.fx * P.fstride + po);
      // (state lines: not the 128-byte groups that lie entirely inside the masked kz interval)
      bool want = true;
      if (P.kzprune) {
        const int kyn = P.yoff + (coln - kxn * P.nyl) * P.ystride;
        const double dn = __dadd_rn(__ldg(P.dax + kxn), __ldg(P.day + kyn));
        const int k0 = u * (N / G::NT);
        want = !(z_mode_dead(P, dn, k0) && z_mode_dead(P, dn, k0 + N / G::NT - 1));
      }
      if (want) {
        prefetch_l2(P.u_in + (size_t)Kn.v * P.fstride + po);
        if (P.read_rk) prefetch_l2(P.fnl_rk + (size_t)Kn.v * P.fstride + po);
      }
    }
    // ---------------- G = (i kx ca) fa + (i ky cb) fb + cx fx ----------------
    if constexpr (NQ == 2) cp_async_wait<1>();   // fa
    else cp_async_wait<0>();
    {
      const double c = K.ca * kxe;
      const bool on = live && K.fa >= 0;
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) r[e] = on ? cmul_i(Q0[e * G::NT + u], c) : mk(0.0, 0.0);
    }
    if constexpr (NQ == 2) {
      land_in(Q0, P.W2 + (size_t)(K.fx >= 0 ? K.fx : 0) * P.fstride + coff, live && K.fx >= 0);
      cp_async_wait<1>();   // fb
      if (live && K.fb >= 0) {
        const double c = K.cb * kye;
        LAPS_UNROLL
        for (int e = 0; e < 8; ++e) r[e] = cadd(r[e], cmul_i(Q1[e * G::NT + u], c));
      }
      land_out(Q1, P.u_in + voff, live, dead);
      cp_async_wait<1>();   // fx
      if (live && K.fx >= 0) {
        LAPS_UNROLL
        for (int e = 0; e < 8; ++e) r[e] = cadd(r[e], cscale(Q0[e * G::NT + u], K.cx));
      }
      land_out(Q0, P.fnl_rk + voff, live && P.read_rk, dead);
    } else {
      // one landing line: fb and fx come straight from L2 (hinted one item ahead) into registers, then u takes the line
      if (live && K.fb >= 0) {
        const cplx* sfb = P.W2 + (size_t)K.fb * P.fstride + coff;
        const double c = K.cb * kye;
        cplx t[8];
        LAPS_UNROLL
        for (int e = 0; e < 8; ++e) t[e] = sfb[u + e * G::NT];
        LAPS_UNROLL
        for (int e = 0; e < 8; ++e) r[e] = cadd(r[e], cmul_i(t[e], c));
      }
      if (live && K.fx >= 0) {
        const cplx* sfx = P.W2 + (size_t)K.fx * P.fstride + coff;
        cplx t[8];
        LAPS_UNROLL
        for (int e = 0; e < 8; ++e) t[e] = sfx[u + e * G::NT];
        LAPS_UNROLL
        for (int e = 0; e < 8; ++e) r[e] = cadd(r[e], cscale(t[e], K.cx));
      }
      land_out(Q0, P.u_in + voff, live, dead);
    }
    FF::first_w(r, u, W, tw0);
    FF::finish_w(r, u, W, tw1, tw2);

    // per-mode table entries of the update, fetched together ahead of their use
    double t_ksqz[8], t_daz[8];
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) {
      const int kz = FF::kout(u, e);
      t_ksqz[e] = __ldg(P.ksq_z + kz);
      t_daz[e] = P.dealias_option ? __ldg(P.daz + kz) : 0.0;
    }
    // ---------------- spectral update on the 8 modes this thread holds ----------------
    const double sgs = K.sg * P.scale;
    const double ca = (P.aeb && K.aeb_c != 0.0) ? K.aeb_c / P.tau : 0.0;                       // mhdrhs.f90:235-247
    const double ce = (K.diff == 1 && P.visc_exp) ? P.nu : ((K.diff == 2 && P.resis_exp) ? P.eta : 0.0);   // :253-275
    const double ci = (K.diff == 1 && P.visc_imp) ? P.nu : ((K.diff == 2 && P.resis_imp) ? P.eta : 0.0);   // rktmod.f90:47-60
    const bool need_ksq = (ce != 0.0) || (ci != 0.0);
    const double ksq_xy = ksq_xy_of(P, kxr, kyr, ksqx, ksqy);
    const bool keep_bg = K.diff == 2 && P.conserve_bg && kx == 0;   // "ix==1 .and. iz==1" skip of mhdrhs.f90:262-270
    const double dfy = day;
    cplx frd[NQ == 2 ? 1 : 8];   // NQ == 1: the RK history straight from L2 into registers
    if constexpr (NQ == 1) {
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e)
        frd[e] = (P.read_rk && live && !((dead >> e) & 1u)) ? P.fnl_rk[voff + FF::kout(u, e)] : mk(0.0, 0.0);
    }
    cp_async_wait<0>();   // u, rk
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) {
      const int kz = FF::kout(u, e);
      cplx fnl = cscale(r[e], sgs);
      if (hasC) fnl = cadd(fnl, S[e * G::NT + u]);
      const bool keep = live && !((dead >> e) & 1u);
      const cplx uo = keep ? Q1[e * G::NT + u] : mk(0.0, 0.0);
      fnl.x -= ca * uo.x;
      fnl.y -= ca * uo.y;
      double ksq = 0.0;
      if (need_ksq) {
        ksq = __dadd_rn(ksq_xy, t_ksqz[e]);
        const double cee = (keep_bg && (P.bg_all_kz || kz == 0)) ? 0.0 : ce;
        fnl.x -= (cee * uo.x) * ksq;
        fnl.y -= (cee * uo.y) * ksq;
      }
      // rkt (rktmod.f90:40-42): u = cc*fnl + dd*fnl_rk + u ; fnl_rk = fnl
      cplx un;
      if (P.read_rk) {
        cplx fr;
        if constexpr (NQ == 2) fr = keep ? Q0[e * G::NT + u] : mk(0.0, 0.0);
        else fr = frd[e];
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
        if (__dadd_rn(dxy, t_daz[e]) >= P.da_thresh) un = mk(0.0, 0.0);
      } else if (P.dealias_option == 2) {
        const double fz = t_daz[e];
        un = mk(__dmul_rn(__dmul_rn(__dmul_rn(un.x, dxy), dfy), fz), __dmul_rn(__dmul_rn(__dmul_rn(un.y, dxy), dfy), fz));
      } else if (P.dealias_option == 3) {   // square truncation (2D/dealiasing.f90:102-117): per-axis flags
        if (dxy != 0.0 || t_daz[e] != 0.0) un = mk(0.0, 0.0);
      }
      if (keep) P.u_out[voff + kz] = un;
      r[e] = un;
    }
    {  // next item's fc -> S (this thread has finished with its S slots)
      if (more) {
        const ZTask& Kn = P.task[ntask];
        land_in(S, P.W2 + (size_t)(Kn.fc >= 0 ? Kn.fc : 0) * P.fstride + (size_t)(coln < 0 ? 0 : coln) * N, Kn.fc >= 0 && coln >= 0);
      } else {
        cp_async_commit();
      }
    }
    // re-shape the register contents into the stage-0 input pattern of the inverse transform
    __syncthreads();  // all last-stage reads of W are done
    if constexpr (G::RLAST != 8) {
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) W[G::pad(FF::kout(u, e))] = r[e];
      __syncthreads();
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) r[e] = W[G::pad(u) + G::pad(e * G::NT)];
      __syncthreads();
    }
    // ---------------- inverse z, stored on the owner of each z (transpose_zy fused) ----------------
    FI::first_w(r, u, W, cconj(tw0));
    FI::finish_w(r, u, W, cconj(tw1), cconj(tw2));
    if (live) {
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) {
        const int z = FI::kout(u, e);
        const int p = P.V1.owner(z);
        cplx* dst = P.V1.base[p] + (((size_t)K.gout * P.nxh + kx) * P.ny + ky) * P.V1.len[p] + (z - P.V1.off[p]);
        *dst = r[e];
      }
    }
    __syncthreads();  // W is refilled by the next item's first stage
    task = ntask; group = ngroup; colm = coln; kxc = kxn;
  }
  cp_async_wait<0>();
}

}  // namespace laps
