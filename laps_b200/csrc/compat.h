// Build-mode glue.  The product is built by nvcc for sm_100a.  The only other mode (LAPS_EMU) is
// the test-only CPU emulation under tests/emu/, which compiles these same sources with g++ to
// check index math in a container without a GPU; it is never shipped or loaded by the package.
#pragma once

#ifdef LAPS_EMU_BUILD
#include "cuda_emu.h"
#define LAPS_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::g_dyn_smem)
#define LAPS_UNROLL
#define LAPS_GRID_CONSTANT
#else
#include <cuda_runtime.h>
#define LAPS_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw_[]; \
  type* name = reinterpret_cast<type*>(name##_raw_)
#define LAPS_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define LAPS_UNROLL _Pragma("unroll")
// kernel parameter structs that are indexed dynamically stay in the constant bank (no local-memory copy)
#define LAPS_GRID_CONSTANT __grid_constant__
#endif

#include <cstdint>

#define LAPS_HD __host__ __device__ __forceinline__
#define LAPS_D __device__ __forceinline__

namespace laps {

typedef double2 cplx;

LAPS_HD cplx mk(double a, double b) { return make_double2(a, b); }
LAPS_HD cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
LAPS_HD cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
LAPS_HD cplx cmul(cplx a, cplx b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
LAPS_HD cplx csqr(cplx a) { return mk(a.x * a.x - a.y * a.y, 2.0 * a.x * a.y); }
LAPS_HD cplx cconj(cplx a) { return mk(a.x, -a.y); }
LAPS_HD cplx cscale(cplx a, double s) { return mk(a.x * s, a.y * s); }
// multiply by i*s (s real): (x + i y) * (i s) = -s y + i s x
LAPS_HD cplx cmul_i(cplx a, double s) { return mk(-s * a.y, s * a.x); }

// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256): two adjacent complex128 values, 32-byte aligned.
LAPS_D void ld256(const cplx* p, cplx& a, cplx& b) {
#ifdef LAPS_EMU_BUILD
  a = p[0]; b = p[1];
#else
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
#endif
}
LAPS_D void st256(cplx* p, cplx a, cplx b) {
#ifdef LAPS_EMU_BUILD
  p[0] = a; p[1] = b;
#else
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
#endif
}

// Asynchronous 16-byte global -> shared copies (LDGSTS, L2-only caching).  Used with THREAD-PRIVATE
// landing slots: the thread that issues a copy is the only one that reads the slot, so
// cp_async_wait<>() is the only synchronisation needed.
LAPS_D void cp_async16(void* smem_dst, const void* gsrc) {
#ifdef LAPS_EMU_BUILD
  *reinterpret_cast<cplx*>(smem_dst) = *reinterpret_cast<const cplx*>(gsrc);
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
#endif
}
LAPS_D void cp_async_commit() {
#ifndef LAPS_EMU_BUILD
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int PENDING>
LAPS_D void cp_async_wait() {   // at most PENDING of this thread's most recent groups still in flight
#ifndef LAPS_EMU_BUILD
  asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
#endif
}

}  // namespace laps
