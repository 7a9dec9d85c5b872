// Micro-benchmark: cost of tcgen05.commit and of mbarrier waits interleaved with MMA groups.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0; d |= (uint64_t)((saddr & 0x3FFFF) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
__host__ __device__ constexpr uint32_t idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a), "l"(b), "r"(id) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra.uni D;\nbra.uni W;\nD:\n}" ::"r"(bar), "r"(parity) : "memory");
}
// MODE 0: groups of 8 TS N=128 MMAs, nothing else.  1: + one commit per group (to a barrier nobody waits on)
// 2: + commit + wait on an already-complete barrier per group.  3: 6 SS N=64 + 2 commits per iteration
// 4: 6 SS N=64 only.  5: groups of 8 TS MMAs + wait on a complete barrier (no commit)
template <int MODE>
__global__ void __launch_bounds__(32, 1) k(int reps, long long* cycles) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_ptr;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 32) reinterpret_cast<uint32_t*>(raw)[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_ptr;
  const uint64_t a_d = make_desc(base), b_d = make_desc(base + 16384);
  // bars[1]: completed once up front so that parity-0 waits always succeed immediately
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
  __syncwarp();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (MODE <= 2 || MODE == 5) {
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) mma_ts(tm, tm + 384 + 8 * kk, b_d + 2 * (kk & 3), idesc(128, 128));
        if (MODE == 1 || MODE == 2) commit(smem_u32(&bars[2]));
      }
      __syncwarp();
      if (MODE == 2 || MODE == 5) wait(smem_u32(&bars[1]), 0);
    } else {
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 6; ++kk) mma_ss(tm + 256, a_d + 2 * (kk & 3), b_d + 2 * (kk & 3), idesc(128, 64));
        if (MODE == 3) { commit(smem_u32(&bars[2])); commit(smem_u32(&bars[3])); }
      }
      __syncwarp();
    }
  }
  if (elect_one()) commit(smem_u32(&bars[0]));
  __syncwarp();
  wait(smem_u32(&bars[0]), 0);
  long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}
template <int MODE> void run(const char* name) {
  long long* d; cudaMalloc(&d, 8);
  const int reps = 1024;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  k<MODE><<<148, 32, 64 * 1024>>>(reps, d);
  k<MODE><<<148, 32, 64 * 1024>>>(reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-52s %8.1f cycles/iteration (%s)\n", name, (double)h / reps, cudaGetErrorString(e));
}
int main() {
  run<0>("8 TS N=128 MMAs (expect 512)");
  run<1>("8 TS N=128 MMAs + 1 commit");
  run<2>("8 TS N=128 MMAs + 1 commit + 1 complete wait");
  run<5>("8 TS N=128 MMAs + 1 complete wait");
  run<4>("6 SS N=64 MMAs (expect 288)");
  run<3>("6 SS N=64 MMAs + 2 commits");
  return 0;
}
