// Fiber-based CTA executor for cuda_emu.h — TEST INFRASTRUCTURE ONLY (see header).
#include "cuda_emu.h"

namespace emu {
uint3e g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;
unsigned char* g_dyn_smem = nullptr;
double g_shfl_scratch[2048];
unsigned long long g_shfl_scratch_u[2048];

namespace {
constexpr size_t kStack = 256 * 1024;
struct Fiber {
  ucontext_t ctx;
  unsigned char* stack = nullptr;
  bool done = false;
  uint3e tid;
};
std::vector<Fiber> g_fibers;
ucontext_t g_sched;
int g_cur = -1;
const std::function<void()>* g_body = nullptr;

void trampoline() {
  (*g_body)();
  g_fibers[g_cur].done = true;
  swapcontext(&g_fibers[g_cur].ctx, &g_sched);
}
}  // namespace

void yield_barrier() {
  int me = g_cur;
  swapcontext(&g_fibers[me].ctx, &g_sched);
  // resumed: restore the thread identity (the scheduler sets it before switching in)
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const unsigned nthr = block.x * block.y * block.z;
  if (g_fibers.size() < nthr) {
    size_t old = g_fibers.size();
    g_fibers.resize(nthr);
    for (size_t i = old; i < nthr; ++i) g_fibers[i].stack = (unsigned char*)std::malloc(kStack);
  }
  std::vector<unsigned char> dyn(smem + 64);
  g_dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
  g_blockDim = block;
  g_gridDim = grid;
  g_body = &body;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_blockIdx = {bx, by, bz};
        std::memset(g_dyn_smem, 0xA5, smem);  // poison: uninitialised reads show up as garbage
        unsigned t = 0;
        for (unsigned tz = 0; tz < block.z; ++tz)
          for (unsigned ty = 0; ty < block.y; ++ty)
            for (unsigned tx = 0; tx < block.x; ++tx, ++t) {
              Fiber& f = g_fibers[t];
              f.done = false;
              f.tid = {tx, ty, tz};
              getcontext(&f.ctx);
              f.ctx.uc_stack.ss_sp = f.stack;
              f.ctx.uc_stack.ss_size = kStack;
              f.ctx.uc_link = &g_sched;
              makecontext(&f.ctx, (void (*)())trampoline, 0);
            }
        unsigned remaining = nthr;
        while (remaining) {
          unsigned ran_done = 0, waiting = 0;
          for (unsigned i = 0; i < nthr; ++i) {
            Fiber& f = g_fibers[i];
            if (f.done) continue;
            g_cur = (int)i;
            g_threadIdx = f.tid;
            swapcontext(&g_sched, &f.ctx);
            if (f.done) { ++ran_done; --remaining; } else { ++waiting; }
          }
          if (waiting && ran_done && remaining) {
            // some threads exited while others wait at a barrier: legal in CUDA only if the
            // exited threads never reach it; the emulator treats exited threads as arrived.
          }
        }
      }
  g_body = nullptr;
  g_dyn_smem = nullptr;
}
}  // namespace emu
