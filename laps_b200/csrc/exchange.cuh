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
// mailboxes are double-buffered by epoch parity.  Every stream that orders passes between the ranks has
// its own flag set (channel) and epoch counter; the allreduce mailboxes belong to channel 0.
#pragma once
#include "compat.h"

namespace laps {

constexpr int kXchgPeers = 8;
constexpr int kMailDoubles = 32;
constexpr int kXchgChannels = 4;   // independent flag sets: one per stream that orders passes between the ranks

struct XchgBlock {                                     // one per rank, in memory every peer maps
  unsigned long long flag[kXchgChannels][kXchgPeers];  // flag[ch][src] = last epoch src has signalled on channel ch
  unsigned long long abort;                            // != 0: some rank gave up (code = reason << 8 | rank + 1); never reset
  double mail[2][kXchgPeers][kMailDoubles];            // [epoch parity][src][j]
};

// why a rank gave up (XchgBlock::abort, laps_last_error)
constexpr unsigned long long kXchgTimeout = 1, kXchgPeerAbort = 2, kXchgHostFailure = 3;

struct XchgPeers {
  XchgBlock* blk[kXchgPeers];                          // blk[p] = rank p's block (peer-mapped)
  int rank, nranks;
  unsigned long long timeout_ns;                       // budget of one wait (LAPS_XCHG_TIMEOUT_S, default 120 s)
  unsigned long long* host_abort;                      // pinned host word of this rank: the code, once a wait has failed
};

LAPS_D void xchg_store_flag(unsigned long long* p, unsigned long long v) {
  *reinterpret_cast<volatile unsigned long long*>(p) = v;
}
LAPS_D unsigned long long xchg_load_flag(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}
LAPS_D unsigned long long xchg_now_ns() {
#ifdef LAPS_EMU_BUILD
  return emu::now_ns();
#else
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
#endif
}

// Tell every rank (and this rank's host) that the exchange is dead: the peers' waits return at once.
LAPS_D void xchg_raise(const XchgPeers& X, unsigned long long reason) {
  const unsigned long long code = (reason << 8) | (unsigned long long)(X.rank + 1);
  for (int p = 0; p < X.nranks; ++p)
    if (xchg_load_flag(&X.blk[p]->abort) == 0) xchg_store_flag(&X.blk[p]->abort, code);
  if (X.host_abort && xchg_load_flag(X.host_abort) == 0) xchg_store_flag(X.host_abort, code);
  __threadfence_system();
}

// All data stores of earlier kernels in this stream are complete when this kernel starts (stream
// order); the system-scope fences order them against the flag stores for the remote observers.
// The wait is bounded: a peer that died, returned an error mid-stage or called the collectives in a
// different order would otherwise wedge every other rank inside this kernel for good (the reference's
// blocking MPI calls can at least be torn down by the MPI runtime).  A rank whose wait runs out of
// budget, or that sees the abort word set, raises the abort word on every rank and returns; the
// kernels behind it in the stream then run on incomplete data, which is why the host treats a raised
// abort word as fatal for the handle (laps_last_error) and the header as fatal for the whole job.
LAPS_D void xchg_signal_and_wait(const XchgPeers& X, int ch, unsigned long long epoch) {
  const int t = threadIdx.x;
  __threadfence_system();
  __syncthreads();
  if (t < X.nranks) xchg_store_flag(&X.blk[t]->flag[ch][X.rank], epoch);
  __syncthreads();
  if (t < X.nranks) {
    const unsigned long long* f = &X.blk[X.rank]->flag[ch][t];
    const unsigned long long* ab = &X.blk[X.rank]->abort;
    const unsigned long long t0 = xchg_now_ns();
    unsigned spins = 0;
    while (xchg_load_flag(f) < epoch) {
      if ((++spins & 63u) == 0) {
        if (xchg_load_flag(ab) != 0) { xchg_raise(X, kXchgPeerAbort); break; }
        if (xchg_now_ns() - t0 > X.timeout_ns) { xchg_raise(X, kXchgTimeout); break; }
      }
#ifdef LAPS_EMU_BUILD
      emu::spin_pause();
#endif
    }
    // a peer that gave up earlier has still signalled its epochs (its kernels run on, on incomplete data): the wait
    // above is then satisfied at once, so look at the abort word in any case
    if (t == 0 && xchg_load_flag(ab) != 0) xchg_raise(X, kXchgPeerAbort);
  }
  __threadfence_system();
  __syncthreads();
}

__global__ void __launch_bounds__(32) k_xchg_barrier(const XchgPeers X, int ch, unsigned long long epoch) {
  xchg_signal_and_wait(X, ch, epoch);
}

// A rank whose host side failed between two collectives releases its peers (they would time out otherwise).
__global__ void __launch_bounds__(32) k_xchg_abort(const XchgPeers X, unsigned long long reason) {
  if (threadIdx.x == 0) xchg_raise(X, reason);
}

// op: 0 sum, 1 min, 2 max.  io[0..n) holds this rank's contribution on entry and the combined
// value on exit.
__global__ void __launch_bounds__(32) k_xchg_allreduce(const XchgPeers X, int ch, unsigned long long epoch,
                                                       double* io, int n, int op) {
  const int t = threadIdx.x;
  const int par = (int)(epoch & 1ull);
  if (t < n) {
    const double v = io[t];
    for (int p = 0; p < X.nranks; ++p)
      *reinterpret_cast<volatile double*>(&X.blk[p]->mail[par][X.rank][t]) = v;
  }
  xchg_signal_and_wait(X, ch, epoch);
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
