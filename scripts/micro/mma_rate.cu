// Micro-benchmark: cycles per tcgen05.mma (kind::tf32 / kind::f16, M=128, cta_group::1) for
// several N and operand sources.  One CTA per SM, one thread issues REPS MMAs back to back.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc(int fmt, int M, int N) {   // fmt: 2 tf32, 0 f16, 1 bf16
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int KIND, bool TS>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_t, uint64_t a_d, uint64_t b_d, uint32_t id, uint32_t acc) {
  if (KIND == 0) {
    if (TS) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a_t), "l"(b_d), "r"(id), "r"(acc) : "memory");
    else asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a_d), "l"(b_d), "r"(id), "r"(acc) : "memory");
  } else {
    if (TS) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a_t), "l"(b_d), "r"(id), "r"(acc) : "memory");
    else asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a_d), "l"(b_d), "r"(id), "r"(acc) : "memory");
  }
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(pred));
  return pred != 0;
}

template <int KIND, bool TS, int N, int NACC, bool ELECT = false>
__global__ void __launch_bounds__(32, 1) rate_kernel(int reps, long long* cycles) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  // zero smem operands so no NaNs
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 32) reinterpret_cast<uint32_t*>(raw)[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_ptr;
  const uint64_t a_d = make_desc(base), b_d = make_desc(base + 16384);
  const uint32_t id = idesc(KIND == 0 ? 2 : 1, 128, N);
  long long t0 = 0, t1 = 0;
  if (ELECT) {
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (elect_one())
          mma<KIND, TS>(tm + ((r * 4 + k) % NACC) * N, tm + 480 + 8 * k, a_d + 2 * k, b_d + 2 * k, id, 1);
        __syncwarp();
      }
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra.uni D2;\nbra.uni W2;\nD2:\n}" ::"r"(smem_u32(&bar)) : "memory");
    t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = t1 - t0;
  } else if (threadIdx.x == 0) {
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // accumulators rotate over NACC tiles; A (TS) sits in the last 32 columns
        mma<KIND, TS>(tm + ((r * 4 + k) % NACC) * N, tm + 480 + 8 * k, a_d + 2 * k, b_d + 2 * k, id, 1);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra.uni D;\nbra.uni W;\nD:\n}" ::"r"(smem_u32(&bar)) : "memory");
    t1 = clock64();
    if (blockIdx.x == 0) *cycles = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

template <int KIND, bool TS, int N, int NACC, bool ELECT = false>
void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 8);
  const int reps = 2048;
  auto k = rate_kernel<KIND, TS, N, NACC, ELECT>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  k<<<grid, 32, 64 * 1024>>>(reps, d);
  k<<<grid, 32, 64 * 1024>>>(reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-34s grid=%3d  %8.1f cycles/MMA   (%s)\n", name, grid, (double)h / (reps * 4), cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {148}) {
    run<1, true, 144, 1, true>("f16 TS N=144 1 acc ELECT", grid);
    run<1, true, 144, 2, true>("f16 TS N=144 2 acc ELECT", grid);
    run<1, true, 128, 2, true>("f16 TS N=128 2 acc ELECT", grid);
    run<1, true, 64, 2, true>("f16 TS N=64 2 acc ELECT", grid);
    run<1, true, 32, 4, true>("f16 TS N=32 4 acc ELECT", grid);
    run<1, false, 64, 2, true>("f16 SS N=64 2 acc ELECT", grid);
    run<1, false, 144, 2, true>("f16 SS N=144 2 acc ELECT", grid);
    run<0, false, 64, 2, true>("tf32 SS N=64 2 acc ELECT", grid);
    run<0, true, 64, 2, true>("tf32 TS N=64 2 acc ELECT", grid);
    run<0, true, 32, 4, true>("tf32 TS N=32 4 acc ELECT", grid);
    run<0, true, 32, 1, true>("tf32 TS N=32 1 acc ELECT", grid);
    run<0, true, 144, 1, true>("tf32 TS N=144 1 acc ELECT", grid);
    run<0, true, 80, 2, true>("tf32 TS N=80 2 acc ELECT", grid);
  }
  return 0;
}
