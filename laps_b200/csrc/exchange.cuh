// Inter-rank synchronisation of the slab exchange, entirely on the device.
//
// The reference's transpose_yz / transpose_zy (parallel.f90:273-324) are P-1 blocking
// mpi_sendrecv calls; here the FFT passes store straight into the peer's buffer over NVLink
// (fft_passes.cuh, spectral_z.cuh), so what is left of the "transpose" is the ordering: a rank may
// read its exchange buffer only after every peer has finished storing into it, and may overwrite a
// peer's buffer only after that peer has finished reading it.  Both are provided by one-CTA
// kernels that run in stream order between the passes:
//   k_xchg_barrier    signal every peer (flag in the peer's memory) then wait for every peer
//   k_xchg_allreduce  the same handshake carrying <= kMailDoubles doubles per rank, combined in
//                     rank order on every rank (mpi_allreduce of mhd.f90:419,567 and
//                     mhdrms.f90:96,98,122: min / max / sum) — bitwise identical on all ranks.
// Flags are monotonically increasing epochs, so no reset (and no reset race) is ever needed; the
// mailboxes are double-buffered by epoch parity.
#pragma once
#include "compat.h"

namespace laps {

constexpr int kXchgPeers = 8;
constexpr int kMailDoubles = 32;

struct XchgBlock {                                     // one per rank, in memory every peer maps
  unsigned long long flag[kXchgPeers];                 // flag[src] = last epoch src has signalled
  double mail[2][kXchgPeers][kMailDoubles];            // [epoch parity][src][j]
};

struct XchgPeers {
  XchgBlock* blk[kXchgPeers];                          // blk[p] = rank p's block (peer-mapped)
  int rank, nranks;
};

LAPS_D void xchg_store_flag(unsigned long long* p, unsigned long long v) {
  *reinterpret_cast<volatile unsigned long long*>(p) = v;
}
LAPS_D unsigned long long xchg_load_flag(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}

// All data stores of earlier kernels in this stream are complete when this kernel starts (stream
// order); the system-scope fences order them against the flag stores for the remote observers.
LAPS_D void xchg_signal_and_wait(const XchgPeers& X, unsigned long long epoch) {
  const int t = threadIdx.x;
  __threadfence_system();
  __syncthreads();
  if (t < X.nranks) xchg_store_flag(&X.blk[t]->flag[X.rank], epoch);
  __syncthreads();
  if (t < X.nranks) {
    const unsigned long long* f = &X.blk[X.rank]->flag[t];
    while (xchg_load_flag(f) < epoch) {
#ifdef LAPS_EMU_BUILD
      emu::spin_pause();
#endif
    }
  }
  __threadfence_system();
  __syncthreads();
}

__global__ void __launch_bounds__(32) k_xchg_barrier(const XchgPeers X, unsigned long long epoch) {
  xchg_signal_and_wait(X, epoch);
}

// op: 0 sum, 1 min, 2 max.  io[0..n) holds this rank's contribution on entry and the combined
// value on exit.
__global__ void __launch_bounds__(32) k_xchg_allreduce(const XchgPeers X, unsigned long long epoch,
                                                       double* io, int n, int op) {
  const int t = threadIdx.x;
  const int par = (int)(epoch & 1ull);
  if (t < n) {
    const double v = io[t];
    for (int p = 0; p < X.nranks; ++p)
      *reinterpret_cast<volatile double*>(&X.blk[p]->mail[par][X.rank][t]) = v;
  }
  xchg_signal_and_wait(X, epoch);
  if (t < n) {
    const XchgBlock* me = X.blk[X.rank];
    double r = *reinterpret_cast<const volatile double*>(&me->mail[par][0][t]);
    for (int p = 1; p < X.nranks; ++p) {
      const double v = *reinterpret_cast<const volatile double*>(&me->mail[par][p][t]);
      r = (op == 0) ? r + v : (op == 1 ? (v < r ? v : r) : (v > r ? v : r));
    }
    io[t] = r;
  }
}

}  // namespace laps
