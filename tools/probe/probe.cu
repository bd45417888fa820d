// Hardware probe for design decisions (not part of the product library).
// Measures: FP64 DFMA throughput, plain copy bandwidth, and copy bandwidth when one side
// is accessed in CHUNK-byte pieces at a large stride (the access pattern of the
// transposing FFT passes), for CHUNK = 32/64/128/256/512 B.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[i];
}

// Transposing copy: in is [A][B] (B fastest, 16-B elements), out is [B/T][A][T]:
// i.e. out chunks of T elements at stride A*T. Reads are contiguous rows; writes are
// T*16-byte chunks. mode 0: strided writes; mode 1: strided reads (swap roles).
template <int T>
__global__ void chunk_copy(const double2* __restrict__ in, double2* __restrict__ out,
                           int A, int B, int mode) {
  // each CTA handles T consecutive "a" rows? No: emulate an FFT pass: a CTA owns T rows
  // a0..a0+T-1 (each B contiguous), writes out[b][a0..a0+T-1] chunks.
  int a0 = blockIdx.x * T;
  extern __shared__ double2 sm[];
  // load T rows coalesced into smem [t][b] with odd pitch
  const int P = B + 1;
  if (mode == 0) {
    for (int idx = threadIdx.x; idx < T * B; idx += blockDim.x) {
      int t = idx / B, b = idx % B;
      sm[t * P + b] = in[(size_t)(a0 + t) * B + b];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < T * B; idx += blockDim.x) {
      int b = idx / T, t = idx % T;
      out[(size_t)b * A + a0 + t] = sm[t * P + b];
    }
  } else {
    for (int idx = threadIdx.x; idx < T * B; idx += blockDim.x) {
      int b = idx / T, t = idx % T;
      sm[t * P + b] = in[(size_t)b * A + a0 + t];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < T * B; idx += blockDim.x) {
      int t = idx / B, b = idx % B;
      out[(size_t)(a0 + t) * B + b] = sm[t * P + b];
    }
  }
}

template <int T>
void run_chunk(const double2* in, double2* out, int A, int B, int mode) {
  size_t smem = (size_t)T * (B + 1) * sizeof(double2);
  CK(cudaFuncSetAttribute(chunk_copy<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int threads = 256;
  for (int it = 0; it < 2; ++it) chunk_copy<T><<<A / T, threads, smem>>>(in, out, A, B, mode);
  CK(cudaEventRecord(e0));
  const int reps = 5;
  for (int it = 0; it < reps; ++it) chunk_copy<T><<<A / T, threads, smem>>>(in, out, A, B, mode);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  double gb = 2.0 * A * (double)B * 16 / 1e9;
  printf("chunk_copy T=%d (%d B chunks) mode=%s: %.3f ms  %.1f GB/s\n", T, T * 16,
         mode == 0 ? "strided-write" : "strided-read", ms, gb / (ms * 1e-3));
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device: %s  sm_%d%d  SMs=%d  smem/SM=%zu  smem/block optin=%zu  L2=%d MB  mem=%.1f GB\n",
         p.name, p.major, p.minor, p.multiProcessorCount, p.sharedMemPerMultiprocessor,
         p.sharedMemPerBlockOptin, p.l2CacheSize >> 20, p.totalGlobalMem / 1e9);
  int v; cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, 0); printf("clock kHz=%d\n", v);
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, 0); printf("max persisting L2=%d MB\n", v >> 20);

  // --- DFMA ---
  {
    int blocks = p.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    double* out; CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
    dfma_kernel<<<blocks, threads>>>(out, 1024);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    dfma_kernel<<<blocks, threads>>>(out, iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * 8 * (double)iters * blocks * threads;
    printf("DFMA: %.3f ms  %.2f TFLOP/s FP64\n", ms, flops / (ms * 1e-3) / 1e12);
    cudaFree(out);
  }
  // --- plain copy ---
  size_t n = (size_t)1 << 28;  // 256 Mi double2 = 4 GiB
  double2 *a, *b; CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
  CK(cudaMemset(a, 1, n * 16)); CK(cudaMemset(b, 0, n * 16));
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int it = 0; it < 2; ++it) copy_kernel<<<p.multiProcessorCount * 16, 512>>>(a, b, n);
    CK(cudaEventRecord(e0));
    for (int it = 0; it < 5; ++it) copy_kernel<<<p.multiProcessorCount * 16, 512>>>(a, b, n);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
    printf("copy 4 GiB double2: %.3f ms  %.1f GB/s (r+w)\n", ms, 2.0 * n * 16 / 1e9 / (ms * 1e-3));
    CK(cudaEventRecord(e0));
    for (int it = 0; it < 5; ++it) CK(cudaMemcpyAsync(b, a, n * 16, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
    printf("cudaMemcpy D2D 4 GiB: %.3f ms  %.1f GB/s (r+w)\n", ms, 2.0 * n * 16 / 1e9 / (ms * 1e-3));
  }
  // --- chunked transposing copies: A rows of B=512 elements (8 KB rows) ---
  {
    int B = 512; int A = (int)(n / B);
    for (int mode = 0; mode < 2; ++mode) {
      run_chunk<2>(a, b, A, B, mode);
      run_chunk<4>(a, b, A, B, mode);
      run_chunk<8>(a, b, A, B, mode);
      run_chunk<16>(a, b, A, B, mode);
      run_chunk<32>(a, b, A, B, mode);
    }
  }
  CK(cudaDeviceSynchronize());
  printf("probe done\n");
  return 0;
}
