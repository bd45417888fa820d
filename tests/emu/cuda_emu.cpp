// Fiber-based CTA executor for cuda_emu.h — TEST INFRASTRUCTURE ONLY (see header).
#include "cuda_emu.h"

#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <map>
#include <string>

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

// ---- shared-memory backed "device" allocations --------------------------------------------
namespace {
struct Region { std::string name; size_t size; bool owner; };
std::map<void*, Region> g_regions;
unsigned g_counter = 0;
struct Cleanup {
  ~Cleanup() { for (auto& kv : g_regions) if (kv.second.owner) shm_unlink(kv.second.name.c_str()); }
} g_cleanup;
}  // namespace

void* shm_alloc(size_t n) {
  char name[64];
  std::snprintf(name, sizeof(name), "/laps_emu_%d_%u", (int)getpid(), g_counter++);
  int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0) return nullptr;
  if (ftruncate(fd, (off_t)n) != 0) { close(fd); shm_unlink(name); return nullptr; }
  void* p = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) { shm_unlink(name); return nullptr; }
  g_regions[p] = Region{name, n, true};
  return p;   // fresh shared memory is zero-filled, like the calloc it replaces
}

void shm_free(void* p) {
  if (!p) return;
  auto it = g_regions.find(p);
  if (it == g_regions.end()) return;
  munmap(p, it->second.size);
  if (it->second.owner) shm_unlink(it->second.name.c_str());
  g_regions.erase(it);
}

int shm_export(void* p, char name_out[64]) {
  auto it = g_regions.find(p);
  if (it == g_regions.end()) return 1;
  std::memset(name_out, 0, 64);
  std::snprintf(name_out, 48, "%s", it->second.name.c_str());
  unsigned long long sz = it->second.size;
  std::memcpy(name_out + 48, &sz, 8);
  return 0;
}

void* shm_import(const char name[64]) {
  unsigned long long sz = 0;
  std::memcpy(&sz, name + 48, 8);
  int fd = shm_open(name, O_RDWR, 0600);
  if (fd < 0) return nullptr;
  void* p = mmap(nullptr, (size_t)sz, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) return nullptr;
  g_regions[p] = Region{std::string(name), (size_t)sz, false};
  return p;
}

void shm_unmap(void* p) { shm_free(p); }

void spin_pause() { sched_yield(); }
unsigned long long now_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}

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
