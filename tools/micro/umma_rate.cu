// Micro-benchmark (developer aid): cycles per tcgen05.mma (cta_group::1, kind::f16, bf16, M = 128, K = 16) as a
// function of N with both operands in shared memory (SS mode), K-major SWIZZLE_128B tiles, nothing else running on
// the SM.  Answers: is the tensor pipe paced by its floor (128 * N / 256 cycles) or by the shared-memory operand fetch?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_of(uint32_t addr) {
  // K-major SWIZZLE_128B, BK = 64 bf16 (128 B rows), 8-row groups 1024 B apart
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int n_mma, int n_acc, int traffic, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t s0 = (smem_u32(raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw + (s0 - smem_u32(raw)))[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    const uint32_t idesc = idesc_bf16(128, N);
    const uint32_t a0 = s0, b0 = s0 + 16384;      // A: 128 x 64 bf16 (16 KB); B: up to 256 x 64 bf16 (32 KB)
    const uint32_t acc_mask = (uint32_t)n_acc - 1u, acc_stride = (uint32_t)N;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t k = (uint32_t)(i & 3) * 32u;  // the four K = 16 slices of the 64-wide tile
      const uint64_t da = desc_of(a0 + k), db = desc_of(b0 + k);
      asm volatile(
          "{\n\t.reg .pred pe, p;\n\t"
          "elect.sync _|pe, 0xffffffff;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (uint32_t)(i & acc_mask) * acc_stride),
          "l"(da), "l"(db), "r"(idesc), "r"(i >= n_acc ? 1u : 0u)
          : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred pe;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(&bar))
        : "memory");
    long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  else if (traffic) {
    // generic-proxy shared-memory traffic beside the MMAs: each lane streams 16-byte loads + stores over a private 8 KB
    const uint32_t base = s0 + 49152 + (warp - 1) * 8192 + (threadIdx.x & 31) * 16;
    volatile int* flag = reinterpret_cast<volatile int*>(&tmem_base_s) + 0;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int it = 0; it < traffic; ++it) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + u * 512) : "memory");
        acc.x ^= v.x;
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(base + u * 512), "r"(acc.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
    }
    if (acc.x == 0x12345678u) *flag = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 74 * 1024 + 1024);
  const int n_mma = 2048;
  printf("M=128 K=16 bf16, SS mode, %d back-to-back MMAs on one SM\n", n_mma);
  printf("%5s %10s %12s %12s %10s %12s\n", "N", "floor cyc", "cyc/MMA", "issue cyc", "fetch B", "B/cyc");
  for (int traffic : {0, 400})
    for (int n_acc : {1, 2, 4})
      for (int N : {32, 64, 96, 128, 192, 256}) {
        if (n_acc * N > 512) continue;
        long long h[2];
        for (int rep = 0; rep < 2; ++rep) {
          rate_kernel<<<1, 128, 74 * 1024 + 1024>>>(N, n_mma, n_acc, traffic, d);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        }
        const double cyc = (double)h[1] / n_mma;
        const int bytes = 128 * 32 + N * 32;
        printf("%5d %10d %12.1f %12.1f %10d %12.1f  accumulators=%d%s\n", N, 128 * N / 256, cyc, (double)h[0] / n_mma, bytes,
               bytes / cyc, n_acc, traffic ? "  + 3 warps of ld/st.shared.v4 traffic" : "");
      }
  return 0;
}
