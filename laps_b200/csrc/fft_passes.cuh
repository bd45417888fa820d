// Batched x- and y-passes of the slab-decomposed 3D FFT (forward: fftw.f90:42-71,136-164;
// inverse: fftw.f90:73-103,203-222 of the reference), one launch per pass over ALL fields.
//
// Work-array layouts (complex128 unless noted; the right-most index is contiguous).  They rotate
// so that every pass reads or writes whole contiguous lines on one side and TL*16-byte chunks on
// the other, and so that the z-pass (spectral_z.cuh) sees contiguous z-lines:
//   real   R [f][zl][y][x]      zl = local z of this rank's slab (z in Zj(rank), parallel.f90:102)
//   W1     [f][kx][zl][y]       after the x-pass      (kx = 0..nx/2)
//   W2     [f][kx][kyl][z]      after the y-pass, ON THE RANK THAT OWNS ky (kyl local, z global):
//                               the y-pass stores straight into the owner's buffer, which is the
//                               reference's transpose_yz (parallel.f90:273-297) fused into the pass
//   V1     [g][kx][ky][zl]      inverse z-pass output, on the rank that owns z (transpose_zy fused)
//   V2     [g][kx][zl][y]       after the inverse y-pass
// The reference's transpose_xy/yx are identity in slab mode (iproc=1, parallel.f90:56-58).
#pragma once
#include "fft_core.cuh"

namespace laps {

constexpr int kMaxPeers = 8;

// Where the slab-exchange side of a pass lives: one base pointer per peer plus the
// decompose_1d tables (parallel.f90:326-349) of the exchanged axis.
struct PeerTable {
  cplx* base[kMaxPeers];
  int off[kMaxPeers];   // first global index owned by peer p
  int len[kMaxPeers];   // number owned
  int nparts;
  int quot;             // n / nparts (decompose_1d: every part has quot entries, the last one the remainder too)
  // cyclic = 0: contiguous slabs, the reference's decompose_1d (parallel.f90:326-349).  cyclic = 1 (ky axis only,
  // LAPS_TUNE_CYCLIC): index i lives on rank i % nparts at local position i / nparts, which spreads the rows the
  // dealiasing mask keeps (low |ky|) evenly over the ranks.
  int cyclic;
  // (no integer divisions on the device: the passes call these once per stored element)
  LAPS_HD int owner(int i) const {
    if (nparts == 1) return 0;
    if (cyclic) return (nparts & (nparts - 1)) == 0 ? (i & (nparts - 1)) : i % nparts;
    int p = 0;
    for (int q = 1; q < kMaxPeers; ++q) p += (q < nparts && i >= off[q]) ? 1 : 0;   // slabs are contiguous and ordered
    return p;
  }
  LAPS_HD int local(int i, int p) const {
    if (!cyclic) return i - off[p];
    return (nparts & (nparts - 1)) == 0 ? (i >> shift) : i / nparts;
  }
  int shift;            // log2(nparts) when nparts is a power of two (cyclic ownership)
};

// Shared-memory tile of TL lines.  A quarter warp (8 lanes of a 128-bit access) covers
// min(TL,8) lines x 8/min(TL,8) adjacent positions when lines are walked side by side, so the
// line pitch is chosen congruent to 8/min(TL,8) modulo 8 to keep those accesses conflict free.
template <int N, int TL>
struct Tile {
  typedef Geom<N> G;
  static constexpr int PITCH = G::pitch(TL >= 8 ? 1 : (TL == 4 ? 2 : (TL == 2 ? 4 : 0)));
  static constexpr int NTHREADS = TL * G::NT;
  static constexpr size_t SMEM = (size_t)TL * PITCH * sizeof(cplx);
  // resident CTAs per SM to compile for: what shared memory allows, but never below 64 registers
  static constexpr int BY_SMEM = (int)((227 * 1024) / (SMEM + 1024));
  static constexpr int WTHREADS = (NTHREADS + 31) / 32 * 32;   // registers and thread slots are handed out per warp
  static constexpr int BY_REGS = 65536 / (WTHREADS * 64);
  static constexpr int BY_THREADS = 2048 / WTHREADS;
  static constexpr int M0 = BY_SMEM < BY_REGS ? BY_SMEM : BY_REGS;
  static constexpr int M1 = M0 < BY_THREADS ? M0 : BY_THREADS;
  static constexpr int MINB = M1 < 1 ? 1 : (M1 > 32 ? 32 : M1);   // what an SM can hold: ptxas ignores bounds beyond it
};

// ---------------------------------------------------------------------------------------------
// forward x: real lines -> half spectra, two real lines per complex transform
// (replaces the r2c loops fftw.f90:58-64 / mhdrhs.f90:143-149, including the "/nx").
// grid.x = planes * (ny / (2*TL)), grid.y = number of fields.  `in` holds `planes` z planes per field (the
// whole slab, or one z chunk of it whose first plane is zl0 of the slab); W1 is always the whole slab.
template <int N, int TL>
__global__ void __launch_bounds__(Tile<N, TL>::NTHREADS, Tile<N, TL>::MINB)
k_fwd_x(const double* __restrict__ in, size_t in_fstride, cplx* __restrict__ W1,
        int nzl, int ny, const cplx* __restrict__ tw, double scale, int nkx, int zl0) {
  typedef Geom<N> G;
  typedef Fft<N, -1> F;
  typedef Tile<N, TL> T;
  LAPS_DYN_SMEM(cplx, sm);
  const int tid = threadIdx.x;
  const int l = tid / G::NT, u = tid % G::NT;
  const int ytiles = ny / (2 * TL);
  const int zl = blockIdx.x / ytiles;
  const int y0 = (blockIdx.x % ytiles) * 2 * TL;
  const int f = blockIdx.y;
  const int nxh = N / 2 + 1;

  const double* ra = in + (size_t)f * in_fstride + ((size_t)zl * ny + y0 + 2 * l) * N;
  const double* rb = ra + N;
  cplx r[8];
  LAPS_UNROLL
  for (int e = 0; e < 8; ++e) r[e] = mk(ra[u + e * G::NT], rb[u + e * G::NT]);
  cplx* line = sm + l * T::PITCH;
  F::first(r, u, line, tw);
  F::template finish_g<TL>(r, u, line, tw, 1 + l);   // a line's transform synchronises its own N/8 threads only
  group_barrier<G::NT, TL>(1 + l);          // everyone has consumed its last-stage slots
  if constexpr (G::RLAST == 8) {            // kout(u, e) = u + e * NT: one padded base + constants
    const int pu = G::pad(u);
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) line[F::pad_in(pu, e)] = r[e];
  } else {
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) line[G::pad(F::kout(u, e))] = r[e];
  }
  __syncthreads();
  // split Z = A + iB into the two half spectra; TL line-pairs side by side give 2*TL*16-byte chunks,
  // each thread storing its (a, b) pair as one 256-bit access
  const double hs = 0.5 * scale;
  constexpr int TOT = (N / 2 + 1) * TL;
  constexpr int ITERS = (TOT + T::NTHREADS - 1) / T::NTHREADS;
  LAPS_UNROLL
  for (int i = 0; i < ITERS; ++i) {
    const int it = tid + i * T::NTHREADS;
    if (it < nkx * TL) {   // nkx <= N/2+1: columns beyond it are removed by the dealiasing mask anyway
      const int lp = it % TL, k = it / TL;
      const cplx zk = sm[lp * T::PITCH + G::pad(k)];
      const cplx zn = sm[lp * T::PITCH + G::pad(G::POW2 ? ((N - k) & (N - 1)) : (k ? N - k : 0))];
      const cplx a = mk((zk.x + zn.x) * hs, (zk.y - zn.y) * hs);
      const cplx b = mk((zk.y + zn.y) * hs, (zn.x - zk.x) * hs);
      st256(W1 + (((size_t)f * nxh + k) * nzl + zl0 + zl) * ny + y0 + 2 * lp, a, b);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward y: contiguous y-lines of W1 -> W2 on the owner of ky, z fastest
// (fftw.f90:156-162 incl. "/ny", fused with transpose_yz parallel.f90:273-297).
// grid.x = ceil(nzl/TL) * nxh, grid.y = fields
template <int N, int TL>
__global__ void __launch_bounds__(Tile<N, TL>::NTHREADS, Tile<N, TL>::MINB)
k_fwd_y(const cplx* __restrict__ W1, PeerTable W2, int nzl, int nz, int zoff,
        const cplx* __restrict__ tw, double scale, int nxh, int kymax_all, const int* __restrict__ kymax_x, int ntiles,
        int zbase, int zcount) {
  typedef Geom<N> G;
  typedef Fft<N, -1> F;
  typedef Tile<N, TL> T;
  LAPS_DYN_SMEM(cplx, sm);
  const int tid = threadIdx.x;
  // the launch covers the z planes [zbase, zbase + zcount) of the slab (the whole slab, or one z chunk of the two-stream schedule)
  const int ztiles = (zcount + TL - 1) / TL;
  const int zend = zbase + zcount;
  const int f = blockIdx.y;
  // one tile per CTA, or — when the launch is held to a few CTAs per SM so that an HBM-bound pass of another stream
  // can share the SMs while this one waits on NVLink — a grid-stride loop over the tiles
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int kx = tile / ztiles;
    const int z0 = zbase + (tile % ztiles) * TL;
    // rows kymax < ky < N - kymax of this kx column are removed by the dealiasing mask for every kz (per-kx
    // table: the mask is a sphere, dealiasing.f90:91-94)
    const int kymax = kymax_x ? __ldg(kymax_x + kx) : kymax_all;
    {
      const int l = tid / G::NT, u = tid % G::NT;  // mapping A: coalesced along the line
      cplx r[8];
      if (z0 + l < zend) {
        const cplx* src = W1 + (((size_t)f * nxh + kx) * nzl + z0 + l) * N;
        LAPS_UNROLL
        for (int e = 0; e < 8; ++e) r[e] = src[u + e * G::NT];
      } else {
        LAPS_UNROLL
        for (int e = 0; e < 8; ++e) r[e] = mk(0.0, 0.0);
      }
      F::first(r, u, sm + l * T::PITCH, tw);
    }
    const int l = tid % TL, u = tid / TL;  // mapping B: TL lines side by side
    cplx r[8];
    F::finish(r, u, sm + l * T::PITCH, tw);
    if (z0 + l < zend) {
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) {
        const int ky = F::kout(u, e);
        if (ky > kymax && ky < N - kymax) continue;   // rows the dealiasing mask removes entirely
        const int p = W2.owner(ky);
        cplx* dst = W2.base[p] + (((size_t)f * nxh + kx) * W2.len[p] + W2.local(ky, p)) * nz + zoff + z0 + l;
        *dst = cscale(r[e], scale);
      }
    }
    if (tile + (int)gridDim.x < ntiles) __syncthreads();   // the next tile's first stage refills the lines
  }
}

// ---------------------------------------------------------------------------------------------
// transpose_yz as a separate, light kernel (two-stream schedule): the forward y pass has stored its lines into a LOCAL
// staging buffer laid out like the owners' W2 blocks but with this rank's z slab only ([peer][f][kx][kyl_p][zl]); this
// kernel moves every row (one (f, kx, kyl) of one peer: nzl contiguous values) to its place in the owner's W2
// ([f][kx][kyl_p][z], at this rank's z offset).  It needs no shared memory and few registers, so it runs beside the
// HBM-bound passes of the compute stream without taking their occupancy, and NVLink is busy while they compute — the
// reference's transpose blocks every rank in mpi_sendrecv (parallel.f90:273-297).
// One warp per (f, kx, row of the rectangle); every warp walks the peers in its own rotation, so that at any moment
// the ranks write to different peers.  Rows outside the dealiasing circle of their kx hold nothing and are skipped.
struct PushParams {
  const cplx* src[kMaxPeers];
  cplx* dst[kMaxPeers];
  int len[kMaxPeers];        // rows (ky) owned by peer p
  int yoff[kMaxPeers];       // global ky of its row kyl: yoff + kyl * ystride
  int nA[kMaxPeers], b0[kMaxPeers];   // its rows inside the rectangle |ky| <= kymax: kyl < nA or kyl >= b0
  int nparts, rank, ystride, ny, nxh, nkx, nzl, nz, zoff, f0, nfc, maxrows;
  const int* kymax_x;        // per-kx circle (nullptr: the rectangle)
  int kymax;
};

__global__ void __launch_bounds__(256) k_xchg_push(const PushParams P) {
  const int lane = threadIdx.x & 31;
  const int warp = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  const int nwarps = (int)((gridDim.x * (unsigned)blockDim.x) >> 5);
  const int total = P.nfc * P.nkx * P.maxrows;
  for (int item = warp; item < total; item += nwarps) {
    const int rl = item % P.maxrows;
    const int t = item / P.maxrows;
    const int kx = t % P.nkx, f = P.f0 + t / P.nkx;
    const int kym = P.kymax_x ? __ldg(P.kymax_x + kx) : P.kymax;
    for (int q = 0; q < P.nparts; ++q) {
      int p = P.rank + 1 + q + item;
      p %= P.nparts;
      const int nlive = P.nA[p] + (P.len[p] - P.b0[p]);
      if (rl >= nlive) continue;
      const int kyl = rl < P.nA[p] ? rl : P.b0[p] + rl - P.nA[p];
      const int ky = P.yoff[p] + kyl * P.ystride;
      if (ky > kym && ky < P.ny - kym) continue;
      const size_t row = ((size_t)f * P.nxh + kx) * P.len[p] + kyl;
      const cplx* s = P.src[p] + row * P.nzl;
      cplx* d = P.dst[p] + row * P.nz + P.zoff;
      for (int i = lane; i < P.nzl; i += 32) d[i] = s[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// inverse y: V1 [g][kx][ky][zl] (z fastest) -> contiguous y-lines V2 [g][kx][zl][y]
// (fftw.f90:212-218, unnormalised).  grid.x = ceil(nzl/TL) * nxh, grid.y = fields
template <int N, int TL>
__global__ void __launch_bounds__(Tile<N, TL>::NTHREADS, Tile<N, TL>::MINB)
k_inv_y(const cplx* __restrict__ V1, cplx* __restrict__ V2, int nzl, const cplx* __restrict__ tw, int nxh, int kymax_all,
        const int* __restrict__ kymax_x) {
  typedef Geom<N> G;
  typedef Fft<N, +1> F;
  typedef Tile<N, TL> T;
  LAPS_DYN_SMEM(cplx, sm);
  const int tid = threadIdx.x;
  const int ztiles = (nzl + TL - 1) / TL;
  const int kx = blockIdx.x / ztiles;
  const int z0 = (blockIdx.x % ztiles) * TL;
  const int g = blockIdx.y;
  const int kymax = kymax_x ? __ldg(kymax_x + kx) : kymax_all;
  {
    const int l = tid % TL, u = tid / TL;  // mapping B
    cplx r[8];
    if (z0 + l < nzl) {
      const cplx* src = V1 + (((size_t)g * nxh + kx) * N) * nzl + z0 + l;
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) {
        const int ky = u + e * G::NT;
        r[e] = (ky > kymax && ky < N - kymax) ? mk(0.0, 0.0) : src[(size_t)ky * nzl];   // masked rows are zero
      }
    } else {
      LAPS_UNROLL
      for (int e = 0; e < 8; ++e) r[e] = mk(0.0, 0.0);
    }
    F::first(r, u, sm + l * T::PITCH, tw);
  }
  const int l = tid / G::NT, u = tid % G::NT;  // mapping A
  cplx r[8];
  F::finish(r, u, sm + l * T::PITCH, tw);
  if (z0 + l < nzl) {
    cplx* dst = V2 + (((size_t)g * nxh + kx) * nzl + z0 + l) * N;
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) dst[F::kout(u, e)] = r[e];
  }
}

// ---------------------------------------------------------------------------------------------
// inverse x: V2 [g][kx][zl][y] -> real lines, two per complex transform
// (c2r loops fftw.f90:94-100 / mhdrhs.f90:354-360; like FFTW's c2r the imaginary parts of the
// DC and Nyquist bins are ignored).  Destination pointer per field (uu(:,:,:,v) or J component).
struct RealDst { double* ptr[16]; };

template <int N, int TL>
__global__ void __launch_bounds__(Tile<N, TL>::NTHREADS, Tile<N, TL>::MINB)
k_inv_x(const cplx* __restrict__ V2, const LAPS_GRID_CONSTANT RealDst dst, int nzl, int ny, const cplx* __restrict__ tw, int nkx) {
  typedef Geom<N> G;
  typedef Fft<N, +1> F;
  typedef Tile<N, TL> T;
  LAPS_DYN_SMEM(cplx, sm);
  const int tid = threadIdx.x;
  const int ytiles = ny / (2 * TL);
  const int zl = blockIdx.x / ytiles;
  const int y0 = (blockIdx.x % ytiles) * 2 * TL;
  const int g = blockIdx.y;
  const int nxh = N / 2 + 1;
  {  // all loads of the tile in flight before the first use (256-bit: the pair of lines of one kx)
    constexpr int TOT = (N / 2 + 1) * TL;
    constexpr int ITERS = (TOT + T::NTHREADS - 1) / T::NTHREADS;
    cplx a[ITERS], b[ITERS];
    LAPS_UNROLL
    for (int i = 0; i < ITERS; ++i) {
      const int it = tid + i * T::NTHREADS;
      a[i] = mk(0.0, 0.0); b[i] = mk(0.0, 0.0);
      if (it < nkx * TL) {   // columns beyond nkx are zero (dealiasing mask)
        const int lp = it % TL, k = it / TL;
        ld256(V2 + (((size_t)g * nxh + k) * nzl + zl) * ny + y0 + 2 * lp, a[i], b[i]);
      }
    }
    LAPS_UNROLL
    for (int i = 0; i < ITERS; ++i) {
      const int it = tid + i * T::NTHREADS;
      if (it < TOT) {
        const int lp = it % TL, k = it / TL;
        cplx av = a[i], bv = b[i];
        if (k == 0 || k == N / 2) { av.y = 0.0; bv.y = 0.0; }
        sm[lp * T::PITCH + G::pad(k)] = mk(av.x - bv.y, av.y + bv.x);
        if (k != 0 && k != N / 2) sm[lp * T::PITCH + G::pad(N - k)] = mk(av.x + bv.y, bv.x - av.y);
      }
    }
  }
  __syncthreads();
  const int l = tid / G::NT, u = tid % G::NT;
  cplx* line = sm + l * T::PITCH;
  cplx r[8];
  {
    const int pu = G::pad(u);
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) r[e] = line[F::in_pos(u, pu, e)];
  }
  // (a power-of-two first stage writes the slots it has read; the composite one writes elsewhere in the line)
  if constexpr (!G::POW2) group_barrier<G::NT, TL>(1 + l);
  F::first(r, u, line, tw);
  F::template finish_g<TL>(r, u, line, tw, 1 + l);   // a line's transform synchronises its own N/8 threads only
  double* oa = dst.ptr[g] + ((size_t)zl * ny + y0 + 2 * l) * N;
  double* ob = oa + N;
  LAPS_UNROLL
  for (int e = 0; e < 8; ++e) {
    const int n = F::kout(u, e);
    oa[n] = r[e].x;
    ob[n] = r[e].y;
  }
}

}  // namespace laps
