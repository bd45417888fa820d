"""Executable model of the in-shared-memory FFT used by laps_b200/csrc/fft_core.cuh.

Mirrors the index math of the CUDA code one to one (digit positions, padded shared-memory
layout, per-stage twiddle exponents, thread->element mappings) so that it can be validated
against numpy.fft here, and so that shared-memory bank conflicts of every access pattern can
be counted without a GPU.  Development tool; also imported by tests/test_fft_model.py.
"""
import numpy as np


class Geom:
    def __init__(self, N):
        assert N >= 16 and (N & (N - 1)) == 0
        self.N = N
        self.log2 = N.bit_length() - 1
        self.nstage = (self.log2 + 2) // 3
        self.rlast = 1 << (self.log2 - 3 * (self.nstage - 1))
        self.nt = N // 8
        self.radix = [8] * (self.nstage - 1) + [self.rlast]
        # position weight of digit s and output weight of digit s
        self.w = []
        self.v = []
        acc = 1
        for s in range(self.nstage):
            self.v.append(acc)
            acc *= self.radix[s]
            self.w.append(N // acc)

    @staticmethod
    def pad(i):
        return i + (i >> 3) + (i >> 6) + (i >> 9)

    def pitch(self, want_mod8=1):
        p = self.pad(self.N - 1) + 1
        while p % 8 != want_mod8:
            p += 1
        return p

    # ---- thread -> positions -------------------------------------------------------
    def stage_positions(self, s, u):
        """The 8 in-place positions thread u touches at stage s (read and written)."""
        g = self
        if s < g.nstage - 1:
            ws = g.w[s]
            base = (u // ws) * (8 * ws) + (u % ws)
            return [base + e * ws for e in range(8)], base % ws
        # last stage
        R = g.rlast
        if R == 8:
            q = u
            base = 0
            for t in range(s):
                kt = (q // g.v[t]) % 8
                base += kt * g.w[t]
            return [base + e for e in range(8)], 0
        vsm1 = g.v[s - 1]
        qlo = u % vsm1
        h = u // vsm1
        o = h
        for t in range(s - 1):
            kt = (qlo // g.v[t]) % 8
            o += kt * (g.w[t] // 8)
        return [8 * o + c for c in range(8)], 0

    def last_stage_outputs(self, u):
        """Output indices k for the 8 register slots after the last stage."""
        g = self
        s = g.nstage - 1
        R = g.rlast
        if R == 8:
            return [u + e * (g.N // 8) for e in range(8)]
        vsm1 = g.v[s - 1]
        qlo = u % vsm1
        h = u // vsm1
        out = []
        per = 8 // R
        for i in range(per):
            for e in range(R):
                out.append(qlo + (per * h + i) * vsm1 + e * (g.N // R))
        return out


def butterfly(vals, R, sign):
    """R-point DFT of the list vals (sign=-1 forward)."""
    vals = np.asarray(vals)
    k = np.arange(R)
    M = np.exp(sign * 2j * np.pi * np.outer(k, k) / R)
    return M @ vals


def fft_model(x, sign=-1):
    """Run the staged in-place algorithm on one line x (complex, length N)."""
    N = len(x)
    g = Geom(N)
    tw = np.exp(sign * 2j * np.pi * np.arange(N) / N)
    sm = np.zeros(g.pad(N - 1) + 1, dtype=complex)
    # stage 0 reads its inputs from "global" in natural order (positions == indices)
    regs = {}
    for u in range(g.nt):
        pos, jp = g.stage_positions(0, u)
        r = butterfly([x[p] for p in pos], 8, sign)
        if g.nstage > 1:
            b = g.v[0] * jp
            r = r * tw[(b * np.arange(8)) % N]
            for e, p in enumerate(pos):
                sm[g.pad(p)] = r[e]
        else:
            regs[u] = r
    for s in range(1, g.nstage):
        last = s == g.nstage - 1
        new = {}
        for u in range(g.nt):
            pos, jp = g.stage_positions(s, u)
            r = np.array([sm[g.pad(p)] for p in pos])
            R = g.radix[s]
            if R == 8:
                r = butterfly(r, 8, sign)
            else:
                per = 8 // R
                r = np.concatenate([butterfly(r[i * R:(i + 1) * R], R, sign) for i in range(per)])
            if not last:
                b = g.v[s] * jp
                r = r * tw[(b * np.arange(8)) % N]
            new[u] = (pos, r)
        for u, (pos, r) in new.items():
            if last:
                regs[u] = r
            else:
                for e, p in enumerate(pos):
                    sm[g.pad(p)] = r[e]
    out = np.zeros(N, dtype=complex)
    for u in range(g.nt):
        for e, k in enumerate(g.last_stage_outputs(u)):
            out[k] = regs[u][e]
    return out


# ---- line lengths with an odd factor: N = P * M, M a power of two (Fft<N, DIR, P> of fft_core.cuh) -----------------
def odd_part(n):
    while n % 2 == 0:
        n //= 2
    return n


def fft_model_odd(x, sign=-1):
    """The composite transform as the CUDA code does it: thread u of the N/8 threads is lane v = u // P of the M-point
    transform of residue class q = u % P (its registers x[u + e N/8] ARE that lane's stage-0 inputs), the P transforms
    run side by side, then one exchange: park W_N^(q k) Y_q[k] in part q of the line, form block j = q as P-term sums."""
    N = len(x)
    P = odd_part(N)
    M = N // P
    g, gm = None, Geom(M)
    nt = N // 8
    twN = np.exp(sign * 2j * np.pi * np.arange(N) / N)
    # the registers of thread u on entry, and the claim that they are lane v's inputs of class q
    for u in range(nt):
        q, v = u % P, u // P
        pos, _ = gm.stage_positions(0, v)
        assert [u + e * nt for e in range(8)] == [P * p_ + q for p_ in pos]
    Y = [fft_model(x[q::P], sign) for q in range(P)]           # the M-point transforms (model above)
    part = lambda q: Geom.pad(q * M)                             # noqa: E731  part q of the padded line
    for q in range(P):                                           # pad(q M + i) = pad(q M) + pad(i): the parts are disjoint
        assert all(Geom.pad(q * M + i) == part(q) + Geom.pad(i) for i in range(M))
    sm = np.zeros(Geom.pad(N - 1) + 1, dtype=complex)
    for u in range(nt):
        q, v = u % P, u // P
        for k in gm.last_stage_outputs(v):
            sm[part(q) + Geom.pad(k)] = Y[q][k] * twN[(q * k) % N]
    out = np.zeros(N, dtype=complex)
    for u in range(nt):
        q, v = u % P, u // P
        wj = [twN[((q * t) % P) * M] for t in range(P)]
        for k in gm.last_stage_outputs(v):
            out[k + q * M] = sum(wj[t] * sm[part(t) + Geom.pad(k)] for t in range(P))     # kout(u, e) = kout_M(v, e) + q M
    return out


def report_conflicts_odd(N, TL):
    """Wavefronts per ideal wavefront of the two phases of the radix-P exchange (mapping A: a line's threads are consecutive)."""
    P = odd_part(N)
    M = N // P
    gm = Geom(M)
    nt = N // 8
    LP = Geom.pad(N - 1) + 1
    while LP % 8 != 1:
        LP += 1
    res = {}
    for phase in ("park", "read"):
        tot = cnt = 0
        for w0 in range(0, TL * nt, 32):
            for e in range(8):
                for t in range(1 if phase == "park" else P):
                    addrs = []
                    for tid in range(w0, min(w0 + 32, TL * nt)):
                        l, u = tid // nt, tid % nt
                        q, v = u % P, u // P
                        k = gm.last_stage_outputs(v)[e]
                        if phase == "read" and t == q:
                            continue                                # the thread's own term stays in its register
                        addrs.append(l * LP + Geom.pad((q if phase == "park" else t) * M) + Geom.pad(k))
                    if addrs:
                        tot += conflicts_16B(addrs)
                        cnt += (len(addrs) + 7) // 8
        res[phase] = tot / cnt
    return res


# ---- bank conflict counting -------------------------------------------------------
def conflicts_16B(addrs_elems):
    """addrs_elems: per-lane element addresses (16-byte elements) of one warp-wide LDS/STS.128.
    Hardware processes a quarter warp (8 lanes x 16 B = 128 B) per wavefront when conflict free.
    Returns the number of wavefronts (ideal: len/8)."""
    wf = 0
    for q in range(0, len(addrs_elems), 8):
        lanes = addrs_elems[q:q + 8]
        bankgroups = {}
        for a in lanes:
            bankgroups.setdefault(a % 8, set()).add(a)
        wf += max(len(v) for v in bankgroups.values())
    return wf


def report_conflicts(N, TL, mappingB_last=False):
    g = Geom(N)
    LP = g.pitch(1)
    nthreads = TL * g.nt
    res = {}
    for s in range(g.nstage):
        for mapping in ("A", "B"):
            worst = 0
            tot = 0
            cnt = 0
            for w0 in range(0, nthreads, 32):
                for e in range(8):
                    addrs = []
                    for tid in range(w0, min(w0 + 32, nthreads)):
                        if mapping == "A":
                            l, u = tid // g.nt, tid % g.nt
                        else:
                            l, u = tid % TL, tid // TL
                        pos, _ = g.stage_positions(s, u)
                        addrs.append(l * LP + g.pad(pos[e]))
                    wf = conflicts_16B(addrs)
                    ideal = (len(addrs) + 7) // 8
                    worst = max(worst, wf / ideal)
                    tot += wf
                    cnt += ideal
            res[(s, mapping)] = (tot / cnt, worst)
    return res


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    for N in (16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
        x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        for sign in (-1, 1):
            y = fft_model(x, sign)
            ref = np.fft.fft(x) if sign < 0 else np.fft.ifft(x) * N
            err = np.abs(y - ref).max() / np.abs(ref).max()
            assert err < 1e-12, (N, sign, err)
        g = Geom(N)
        print(f"N={N}: stages={g.radix} w={g.w} v={g.v} pitch={g.pitch(1)} OK")
    for N, TL in ((64, 32), (256, 8), (512, 4), (512, 8), (1024, 4), (2048, 2)):
        res = report_conflicts(N, TL)
        print(f"N={N} TL={TL}: " + "  ".join(f"s{s}{m}:{a:.2f}/{w:.1f}" for (s, m), (a, w) in sorted(res.items())))
    for N in (48, 80, 96, 160, 192, 320, 384, 640, 768, 1280, 1536):
        x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        for sign in (-1, 1):
            ref = np.fft.fft(x) if sign < 0 else np.fft.ifft(x) * N
            assert np.abs(fft_model_odd(x, sign) - ref).max() / np.abs(ref).max() < 1e-12, (N, sign)
        print(f"N={N} = {odd_part(N)} x {N // odd_part(N)}: OK; exchange wavefronts / ideal (4 lines per CTA): {report_conflicts_odd(N, 4)}")
