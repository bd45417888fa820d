// In-shared-memory FP64 complex FFT building blocks for one "line" of N = 2^L points (8 <= N <= 4096; N = 8 is a
// single register-resident radix-8 stage: one thread per line, no exchange), and of N = P * 2^L points with a small odd
// factor P (3, 5, ...: the second Fft template at the end of the file), which runs P power-of-two transforms side by
// side and joins them with one more exchange.
//
// Replaces the per-line FFTW executions of the reference (fftw.f90:61,97,159,176,198,215 and
// mhdrhs.f90:146,162,357): N/8 threads cooperate on a line, every thread holds 8 points in
// registers per stage, stages are radix-8 (the last one radix 2/4/8), data is exchanged between
// stages IN PLACE through a padded shared-memory line (one __syncthreads per exchange).
//
//   positions:  pos = sum_s d_s * w_s      (d_s = input digit n_s before stage s, output digit k_s after)
//   input index n = pos at the start;  output index k = sum_s k_s * v_s
//   after stage s (not last) output k_s is multiplied by W_N^(v_s * j' * k_s), j' = pos mod w_s.
//
// The index math is mirrored and validated in tools/fft_model.py (tests/test_fft_model.py).
#pragma once
#include "compat.h"

namespace laps {

constexpr int ilog2c(int n) { return n <= 1 ? 0 : 1 + ilog2c(n >> 1); }
constexpr int odd_part(int n) { return (n & 1) ? n : odd_part(n >> 1); }

template <int N>
struct Geom {
  static constexpr int P = odd_part(N);   // 1: a power of two
  static constexpr int M = N / P;         // the power-of-two part
  static constexpr bool POW2 = P == 1;
  static_assert(N >= 8 && N <= 4096 && (POW2 || (M >= 16 && P <= 15)), "line length must be 2^L in [8,4096] or P * 2^L with P odd <= 15, 2^L >= 16");
  // stage structure of the power-of-two transform (the members up to v() are meaningless when P > 1: RLAST = 0 there,
  // so that the shortcuts the kernels take for RLAST == 8 stay switched off)
  static constexpr int LOG2 = ilog2c(M);
  static constexpr int NSTAGE = POW2 ? (LOG2 + 2) / 3 : 0;
  static constexpr int RLAST = POW2 ? 1 << (LOG2 - 3 * (NSTAGE - 1)) : 0;
  static constexpr int NT = N / 8;  // threads per line
  LAPS_HD static constexpr int w(int s) { return s < NSTAGE - 1 ? (N >> (3 * (s + 1))) : 1; }
  LAPS_HD static constexpr int v(int s) { return 1 << (3 * s); }
  LAPS_HD static constexpr int pad(int i) { return i + (i >> 3) + (i >> 6) + (i >> 9); }
  // pad(base + e * w) == pad(base) + pad(e * w) for e < 8 and w a power of two, whenever the three bits of `base` at the
  // position of `e` are zero (every shift then splits without a carry): true for the stage-0 pattern (base = u < w) and
  // for the middle stages (base = (u / w) * 8 w + u % w).  The second term is a compile-time constant in the unrolled
  // loops below, so a padded address costs one addition instead of three shifts and three additions.
  // line pitch (in elements) congruent to `m` modulo 8: m=1 makes accesses that walk 8 lines at a
  // fixed position conflict free, m=2 serves 4 lines x 2 adjacent positions.
  LAPS_HD static constexpr int pitch(int m) {
    int p = pad(N - 1) + 1;
    while ((p & 7) != m) ++p;
    return p;
  }
};

// ------------------------------------------------------------------ butterflies (natural order out)
template <int DIR>
LAPS_D cplx mul_w4(cplx a) {  // a * W4,  W4 = -i (forward, DIR<0) or +i (inverse)
  return DIR < 0 ? mk(a.y, -a.x) : mk(-a.y, a.x);
}

template <int DIR>
LAPS_D void bfly8(cplx (&r)[8]) {
  const double h = 0.70710678118654752440;
  cplx a0 = cadd(r[0], r[4]), a4 = csub(r[0], r[4]);
  cplx a1 = cadd(r[1], r[5]), a5 = csub(r[1], r[5]);
  cplx a2 = cadd(r[2], r[6]), a6 = csub(r[2], r[6]);
  cplx a3 = cadd(r[3], r[7]), a7 = csub(r[3], r[7]);
  // odd branch twiddles W8^1, W8^2, W8^3
  if (DIR < 0) {
    a5 = mk((a5.x + a5.y) * h, (a5.y - a5.x) * h);
    a7 = mk((a7.y - a7.x) * h, -(a7.x + a7.y) * h);
  } else {
    a5 = mk((a5.x - a5.y) * h, (a5.x + a5.y) * h);
    a7 = mk(-(a7.x + a7.y) * h, (a7.x - a7.y) * h);
  }
  a6 = mul_w4<DIR>(a6);
  cplx b0 = cadd(a0, a2), b2 = csub(a0, a2);
  cplx b1 = cadd(a1, a3), b3 = mul_w4<DIR>(csub(a1, a3));
  cplx c0 = cadd(a4, a6), c2 = csub(a4, a6);
  cplx c1 = cadd(a5, a7), c3 = mul_w4<DIR>(csub(a5, a7));
  r[0] = cadd(b0, b1); r[4] = csub(b0, b1);
  r[2] = cadd(b2, b3); r[6] = csub(b2, b3);
  r[1] = cadd(c0, c1); r[5] = csub(c0, c1);
  r[3] = cadd(c2, c3); r[7] = csub(c2, c3);
}

template <int DIR>
LAPS_D void bfly4(cplx& r0, cplx& r1, cplx& r2, cplx& r3) {
  cplx b0 = cadd(r0, r2), b2 = csub(r0, r2);
  cplx b1 = cadd(r1, r3), b3 = mul_w4<DIR>(csub(r1, r3));
  r0 = cadd(b0, b1); r2 = csub(b0, b1);
  r1 = cadd(b2, b3); r3 = csub(b2, b3);
}

LAPS_D void bfly2(cplx& r0, cplx& r1) {
  cplx a = cadd(r0, r1), b = csub(r0, r1);
  r0 = a; r1 = b;
}

// r[e] *= w^e, e = 1..7.  Powers are built from the base twiddle by multiplication
// (3 squarings + 3 products), which keeps shared-memory/L1 traffic for twiddles negligible.
LAPS_D void twiddle8(cplx (&r)[8], cplx w) {
  cplx w2 = csqr(w);
  cplx w3 = cmul(w2, w);
  cplx w4 = csqr(w2);
  r[1] = cmul(r[1], w);
  r[2] = cmul(r[2], w2);
  r[3] = cmul(r[3], w3);
  r[4] = cmul(r[4], w4);
  r[5] = cmul(r[5], cmul(w4, w));
  r[6] = cmul(r[6], csqr(w3));
  r[7] = cmul(r[7], cmul(w4, w3));
}

// Barrier over the NT threads of one line (whole warps: NT % 32 == 0), else over the CTA.  `bar` is
// in 1..NBAR (one named barrier per line of the CTA); the ids are immediates so that ptxas reserves
// NBAR + 1 hardware barriers for the CTA, not all sixteen.
template <int NT, int NBAR>
LAPS_D void group_barrier(int bar) {
#ifdef LAPS_EMU_BUILD
  (void)bar;
  __syncthreads();
#else
  static_assert(NBAR <= 15, "a CTA has 16 named barriers");
  if constexpr (NT % 32 == 0) {
#define LAPS_BAR_CASE(I) case I: if constexpr (I <= NBAR) asm volatile("bar.sync " #I ", %0;" ::"n"(NT) : "memory"); break;
    switch (bar) {
      LAPS_BAR_CASE(1) LAPS_BAR_CASE(2) LAPS_BAR_CASE(3) LAPS_BAR_CASE(4) LAPS_BAR_CASE(5)
      LAPS_BAR_CASE(6) LAPS_BAR_CASE(7) LAPS_BAR_CASE(8) LAPS_BAR_CASE(9) LAPS_BAR_CASE(10)
      LAPS_BAR_CASE(11) LAPS_BAR_CASE(12) LAPS_BAR_CASE(13) LAPS_BAR_CASE(14) LAPS_BAR_CASE(15)
      default: break;
    }
#undef LAPS_BAR_CASE
  } else {
    __syncthreads();
  }
#endif
}

// ------------------------------------------------------------------ staged FFT of one line
// `tw` points to the forward table tw[m] = exp(-2 pi i m / N), m < N (global memory, L1 resident).
template <int N, int DIR, int P = Geom<N>::P>
struct Fft;

template <int N, int DIR>
struct Fft<N, DIR, 1> {
  typedef Geom<N> G;

  LAPS_D static cplx twid(const cplx* __restrict__ tw, int m) {
    cplx w = __ldg(tw + m);
    if (DIR > 0) w.y = -w.y;
    return w;
  }

  // The base twiddle a thread needs in stage S depends only on its position u in the line, so a
  // persistent kernel can fetch it once (tw_stage) and pass it to the *_w variants below.
  template <int S>
  LAPS_D static cplx tw_stage(const cplx* __restrict__ tw, int u) {
    if (S == 0) return twid(tw, u);                       // v_0 = 1, j' = u
    constexpr int ws = G::w(S);
    return twid(tw, G::v(S) * (u & (ws - 1)));
  }

  // Stage 0.  r[e] = x[u + e*N/8] on entry.  Leaves the twiddled outputs in the smem line.
  LAPS_D static void first_w(cplx (&r)[8], int u, cplx* __restrict__ line, cplx w) {
    bfly8<DIR>(r);
    if constexpr (G::NSTAGE == 1) return;   // N = 8: r[e] = X[e] already, nothing goes through the line
    twiddle8(r, w);
    const int pu = G::pad(u);
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) line[pu + G::pad(e * G::w(0))] = r[e];
  }
  LAPS_D static void first(cplx (&r)[8], int u, cplx* __restrict__ line, const cplx* __restrict__ tw) {
    first_w(r, u, line, tw_stage<0>(tw, u));
  }

  // Middle stage S (1 <= S < NSTAGE-1), in place.
  template <int S>
  LAPS_D static void middle_w(int u, cplx* __restrict__ line, cplx w) {
    constexpr int ws = G::w(S);
    const int jp = u & (ws - 1);
    const int pb = G::pad((u / ws) * (8 * ws) + jp);
    cplx r[8];
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) r[e] = line[pb + G::pad(e * ws)];
    bfly8<DIR>(r);
    twiddle8(r, w);
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) line[pb + G::pad(e * ws)] = r[e];
  }
  template <int S>
  LAPS_D static void middle(int u, cplx* __restrict__ line, const cplx* __restrict__ tw) {
    middle_w<S>(u, line, tw_stage<S>(tw, u));
  }

  // Base position (8 consecutive slots base..base+7) read by thread u in the last stage.
  LAPS_D static int last_base(int u) {
    constexpr int s = G::NSTAGE - 1;
    if constexpr (G::RLAST == 8) {
      int base = 0;
      LAPS_UNROLL
      for (int t = 0; t < s; ++t) base += ((u >> (3 * t)) & 7) * G::w(t);
      return base;
    } else {
      constexpr int vsm1 = G::v(s - 1);
      const int qlo = u & (vsm1 - 1);
      int o = u / vsm1;
      LAPS_UNROLL
      for (int t = 0; t < s - 1; ++t) o += ((qlo >> (3 * t)) & 7) * (G::w(t) / 8);
      return 8 * o;
    }
  }

  // padded position of element u + e * NT (the stage-0 input pattern; for RLAST == 8 also the output pattern kout):
  // pad(u) + a constant, see Geom::pad
  LAPS_D static int pad_in(int pu /* = G::pad(u) */, int e) { return pu + G::pad(e * G::NT); }
  LAPS_D static int in_pos(int /*u*/, int pu, int e) { return pad_in(pu, e); }

  // Output index k held in register slot e of thread u after the last stage.
  LAPS_D static int kout(int u, int e) {
    if constexpr (G::RLAST == 8) {
      return u + e * (N / 8);
    } else {
      constexpr int s = G::NSTAGE - 1;
      constexpr int R = G::RLAST;
      constexpr int vsm1 = G::v(s - 1);
      const int qlo = u & (vsm1 - 1);
      const int h = u / vsm1;
      const int i = e / R, ee = e % R;
      return qlo + ((8 / R) * h + i) * vsm1 + ee * (N / R);
    }
  }

  // Last stage: reads the line, leaves X[kout(u,e)] in r[e].
  LAPS_D static void last(cplx (&r)[8], int u, const cplx* __restrict__ line) {
    const int pb = G::pad(last_base(u));   // a multiple of 8: the eight slots are consecutive in the padded line too
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) r[e] = line[pb + e];
    if (G::RLAST == 8) {
      bfly8<DIR>(r);
    } else if (G::RLAST == 4) {
      bfly4<DIR>(r[0], r[1], r[2], r[3]);
      bfly4<DIR>(r[4], r[5], r[6], r[7]);
    } else {
      bfly2(r[0], r[1]); bfly2(r[2], r[3]); bfly2(r[4], r[5]); bfly2(r[6], r[7]);
    }
  }

  // All stages after `first` up to and including `last`; every thread of the CTA must call it
  // (it contains the block barriers).  `line` is this thread's line.
  LAPS_D static void finish(cplx (&r)[8], int u, cplx* __restrict__ line, const cplx* __restrict__ tw) {
    if constexpr (G::NSTAGE == 1) return;
    __syncthreads();
    if constexpr (G::NSTAGE >= 3) { middle<1>(u, line, tw); __syncthreads(); }
    if constexpr (G::NSTAGE >= 4) { middle<2>(u, line, tw); __syncthreads(); }
    last(r, u, line);
  }
  // The same with barriers over ONE line's threads only (named barrier `bar`, 1..15): for kernels in
  // which every line is transformed by whole warps of its own, so that lines need not wait for each other.
  template <int NBAR>
  LAPS_D static void finish_g(cplx (&r)[8], int u, cplx* __restrict__ line, const cplx* __restrict__ tw, int bar) {
    if constexpr (G::NSTAGE == 1) return;
    group_barrier<G::NT, NBAR>(bar);
    if constexpr (G::NSTAGE >= 3) { middle<1>(u, line, tw); group_barrier<G::NT, NBAR>(bar); }
    if constexpr (G::NSTAGE >= 4) { middle<2>(u, line, tw); group_barrier<G::NT, NBAR>(bar); }
    last(r, u, line);
  }
  LAPS_D static void finish_w(cplx (&r)[8], int u, cplx* __restrict__ line, cplx w1, cplx w2) {
    if constexpr (G::NSTAGE == 1) return;
    __syncthreads();
    if constexpr (G::NSTAGE >= 3) { middle_w<1>(u, line, w1); __syncthreads(); }
    if constexpr (G::NSTAGE >= 4) { middle_w<2>(u, line, w2); __syncthreads(); }
    last(r, u, line);
  }
};

// ------------------------------------------------------------------ N = P * M, P odd (3, 5, ...), M = 2^L >= 16
// FFTW plans any length (fftw.f90:27-33); this covers the lengths with one small odd factor.  Decimation in time over
// the P residue classes of the input index:
//   X[k + j M] = sum_q  w_P^(j q) * ( W_N^(q k) * Y_q[k] ),     Y_q = FFT_M( x[P n + q], n < M ),   k < M, j < P.
// Thread u of the N/8 threads holds x[u + e N/8] = x[P (v + e M/8) + q] with q = u % P, v = u / P: exactly the stage-0
// input pattern of lane v of the M-point transform of class q, so the P transforms run side by side on the same
// registers, each in its own part of the line (class q at pad(q M) — pad(q M + i) = pad(q M) + pad(i) for i < M, so the
// parts are disjoint and a part is addressed by the M-point transform's own padded offsets).  One more exchange joins
// them: every thread parks its twiddled Y_q[k], then forms output block j = q at the k's it holds (P-term sums, the
// thread's own term from its register; each radix-P butterfly is evaluated by P threads, one output each, which keeps
// eight values per thread throughout).
// Same interface as the power-of-two transform (first / finish / finish_g / kout / in_pos) with RLAST = 0; the
// stage-level entry points (first_w, middle_w, finish_w) of the pipelined z pass do not exist here.
// `tw`: W_N table (N entries) followed by the W_M table (M entries), see twiddle_table() in solver.cu.
template <int N, int DIR, int P>
struct Fft {
  typedef Geom<N> G;
  static constexpr int M = G::M;
  typedef Geom<M> GM;
  typedef Fft<M, DIR, 1> FM;
  static_assert(GM::NSTAGE >= 2, "the sub-transforms go through the line");

  LAPS_D static cplx twid(const cplx* __restrict__ tw, int m) {
    cplx w = __ldg(tw + m);
    if (DIR > 0) w.y = -w.y;
    return w;
  }
  template <int NBAR>
  LAPS_D static void sync(int bar) {
    if constexpr (NBAR == 0) { (void)bar; __syncthreads(); }
    else group_barrier<G::NT, NBAR>(bar);
  }

  LAPS_D static void first(cplx (&r)[8], int u, cplx* __restrict__ line, const cplx* __restrict__ tw) {
    FM::first(r, u / P, line + G::pad((u % P) * M), tw + N);
  }

  LAPS_D static int in_pos(int u, int /*pu*/, int e) { return G::pad(u + e * G::NT); }

  LAPS_D static int kout(int u, int e) { return FM::kout(u / P, e) + (u % P) * M; }

  template <int NBAR>
  LAPS_D static void finish_b(cplx (&r)[8], int u, cplx* __restrict__ line, const cplx* __restrict__ tw, int bar) {
    const int q = u % P, v = u / P;
    cplx* sub = line + G::pad(q * M);
    const cplx* twm = tw + N;
    sync<NBAR>(bar);
    if constexpr (GM::NSTAGE >= 3) { FM::template middle<1>(v, sub, twm); sync<NBAR>(bar); }
    if constexpr (GM::NSTAGE >= 4) { FM::template middle<2>(v, sub, twm); sync<NBAR>(bar); }
    FM::last(r, v, sub);
    sync<NBAR>(bar);   // every last-stage slot has been read: the parts can be refilled
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) {
      const int k = FM::kout(v, e);
      if (q != 0) r[e] = cmul(r[e], twid(tw, q * k));
      sub[GM::pad(k)] = r[e];
    }
    cplx wj[P];        // w_P^(q t), t < P: this thread forms block j = q
    LAPS_UNROLL
    for (int t = 1; t < P; ++t) wj[t] = twid(tw, ((q * t) % P) * M);
    sync<NBAR>(bar);
    LAPS_UNROLL
    for (int e = 0; e < 8; ++e) {   // the term of this thread's own class stays in its register: P - 1 reads per output
      const int pk = GM::pad(FM::kout(v, e));
      cplx acc = r[e];
      if (q != 0) acc = line[pk];
      LAPS_UNROLL
      for (int t = 1; t < P; ++t) {
        cplx y = r[e];
        if (t != q) y = line[G::pad(t * M) + pk];
        acc = cadd(acc, cmul(y, wj[t]));
      }
      r[e] = acc;
    }
  }
  LAPS_D static void finish(cplx (&r)[8], int u, cplx* __restrict__ line, const cplx* __restrict__ tw) {
    finish_b<0>(r, u, line, tw, 0);
  }
  template <int NBAR>
  LAPS_D static void finish_g(cplx (&r)[8], int u, cplx* __restrict__ line, const cplx* __restrict__ tw, int bar) {
    finish_b<NBAR>(r, u, line, tw, bar);
  }
};

}  // namespace laps
