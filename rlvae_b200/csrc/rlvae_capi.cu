// extern "C" entry points of librlvae_b200.so (declared in include/rlvae_b200.h) plus the
// table-packing kernels behind rlvae_tables_create.
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#include <cuda_fp16.h>

#include "rlvae_internal.h"

namespace rlvae {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
static std::atomic<long long> g_launch_count{0};
void count_launch() { g_launch_count.fetch_add(1, std::memory_order_relaxed); }

// ---- per-kernel timing of rlvae_metric_eval on the packed tensor path (bench.py's roofline block):
// CUDA events recorded on the launching stream around the forward launch, the fallback pass and the
// gradient launch of every call while profiling is on.  Not thread-safe; a measurement aid.
static struct {
  bool on = false;
  int cap = 0, n = 0;
  cudaEvent_t* ev = nullptr;     // 4 per record
} g_prof;
void prof_mark(int slot, cudaStream_t s) {
  if (g_prof.on && g_prof.n < g_prof.cap) cudaEventRecord(g_prof.ev[4 * g_prof.n + slot], s);
}
static void prof_next() {
  if (g_prof.on && g_prof.n < g_prof.cap) ++g_prof.n;
}

__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// centroid-side tables: padded c, ||c||^2, [hi|lo] stack and the exp2 bias of the tensor path
__global__ void pack_centroids_kernel(const float* __restrict__ c_in, int K, int Kpad, int d,
                                      float inv_T2_log2e, float* __restrict__ c, float* __restrict__ cn,
                                      float* __restrict__ cstack, float* __restrict__ cbias,
                                      float* __restrict__ ct_hi, float* __restrict__ ct_lo,
                                      float* __restrict__ cn_inf, float* __restrict__ cmask,
                                      float* __restrict__ stats) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Kpad) return;
  float nrm = 0.f, amax = 0.f;
  for (int j = 0; j < d; ++j) {
    const float v = (k < K) ? c_in[(int64_t)k * d + j] : 0.f;
    c[(int64_t)k * d + j] = v;
    nrm = fmaf(v, v, nrm);
    amax = fmaxf(amax, fabsf(v));
  }
  cn[k] = nrm;
  if (k < K) {
    atomicAdd(&stats[0], nrm);                                  // sum ||c||^2
    atomicMax(reinterpret_cast<int*>(&stats[1]), __float_as_int(nrm));  // max (nrm >= 0)
    if (isfinite(amax)) atomicMax(reinterpret_cast<int*>(&stats[4]), __float_as_int(amax));  // max |c_kj|
  }
  if (cstack != nullptr) {  // d == 16
    for (int j = 0; j < 16; ++j) {
      const float v = (k < K) ? c_in[(int64_t)k * 16 + j] : 0.f;
      const float hi = tf32_hi(v);
      cstack[(int64_t)k * 32 + j] = hi;
      cstack[(int64_t)k * 32 + 16 + j] = v - hi;
      // B operand of the gradient kernel's final contraction: transposed [32, Kpad] (centroid index
      // contiguous), rows 0..15 = c, row 16 = 1 (so that the same GEMM also yields sum_k u), rest 0
      ct_hi[(int64_t)j * Kpad + k] = hi;
      ct_lo[(int64_t)j * Kpad + k] = v - hi;
      ct_hi[(int64_t)(16 + j) * Kpad + k] = (j == 0) ? 1.f : 0.f;
      ct_lo[(int64_t)(16 + j) * Kpad + k] = 0.f;
    }
    cbias[k] = (k < K) ? -nrm * inv_T2_log2e : -1.0e30f;
    cn_inf[k] = (k < K) ? nrm : 3.0e38f;
    cmask[k] = (k < K) ? 0.f : -1.0e30f;
  }
}

// matrix-side tables: padded natural M, and for d == 16 the TF32 hi/lo splits in both the
// natural [Kpad,256] and the transposed [256,Kpad] (centroid-contiguous) layouts.
__global__ void pack_matrices_kernel(const float* __restrict__ m_in, int K, int Kpad, int dd,
                                     float* __restrict__ M, float* __restrict__ mt_hi,
                                     float* __restrict__ mt_lo, float* __restrict__ mn_hi,
                                     float* __restrict__ mn_lo, float* __restrict__ stats) {
  const int64_t total = (int64_t)Kpad * dd;
  float amax = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / dd), q = (int)(i - (int64_t)k * dd);
    const float v = (k < K) ? m_in[i] : 0.f;
    M[i] = v;
    amax = fmaxf(amax, fabsf(v));
    if (mt_hi != nullptr) {
      const float hi = tf32_hi(v);
      mn_hi[i] = hi;
      mn_lo[i] = v - hi;
      mt_hi[(int64_t)q * Kpad + k] = hi;
      mt_lo[(int64_t)q * Kpad + k] = v - hi;
    }
  }
  if (amax > 0.f && isfinite(amax)) atomicMax(reinterpret_cast<int*>(&stats[3]), __float_as_int(amax));
}

// mean centroid (d == 16): the expanded distance form is evaluated about it (see pack_c16h_kernel)
__global__ void centroid_mean_kernel(const float* __restrict__ c, int K, float* __restrict__ shift) {
  __shared__ double part[256][17];
  double acc[16];
  for (int j = 0; j < 16; ++j) acc[j] = 0.0;
  for (int k = threadIdx.x; k < K; k += blockDim.x)
    for (int j = 0; j < 16; ++j) acc[j] += (double)c[(int64_t)k * 16 + j];
  for (int j = 0; j < 16; ++j) part[threadIdx.x][j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 16) {
    double tot = 0.0;
    for (int i = 0; i < (int)blockDim.x; ++i) tot += part[i][threadIdx.x];
    const float m = (float)(tot / (double)K);
    shift[threadIdx.x] = isfinite(m) ? m : 0.f;
  }
}

// split-fp16 centroid rows for GEMM1 (d == 16): [Kpad, 64] fp16 = [fp16(2^ec c~) (16) | residual (16) | 0 (32)]
// (one 128-byte swizzle row per centroid, like cstack), with c~ = c - shift.  ||z-c||^2 is invariant
// under a common translation, and the cancellation error of ||z~||^2+||c~||^2-2 z~.c~ scales with
// ||c~||^2, so centring the table widens the range of temperatures the expanded form can serve.
// bias_h[k] = -||c~_k||^2 log2(e)/T^2 (padding: -1e30); stats[5] accumulates sum ||c~||^2, [6] max |c~|.
__global__ void pack_c16h_kernel(const float* __restrict__ c, const float* __restrict__ shift, int K, int Kpad,
                                 float scale, float inv_T2_log2e, __half* __restrict__ out,
                                 float* __restrict__ bias_h, float* __restrict__ ctc_hi,
                                 float* __restrict__ ctc_lo) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Kpad) return;
  float nrm = 0.f;
  for (int j = 0; j < 16; ++j) {
    const float cv = (k < K) ? c[(int64_t)k * 16 + j] - shift[j] : 0.f;
    nrm = fmaf(cv, cv, nrm);
    const float v = scale * cv;
    const __half h = __float2half_rn(v);
    out[(int64_t)k * 64 + j] = h;
    out[(int64_t)k * 64 + 16 + j] = __float2half_rn(v - __half2float(h));
    const float thi = tf32_hi(cv);
    ctc_hi[(int64_t)j * Kpad + k] = thi;
    ctc_lo[(int64_t)j * Kpad + k] = cv - thi;
  }
  for (int j = 32; j < 64; ++j) out[(int64_t)k * 64 + j] = __float2half_rn(0.f);
  bias_h[k] = (k < K) ? -nrm * inv_T2_log2e : -1.0e30f;
}

// sum_k ||M_k||_F into out[0] (bound on what the un-refined weights of the hybrid mode can add up to)
__global__ void matrix_norm_sum_kernel(const float* __restrict__ M, int K, int dd, float* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  float nrm = 0.f;
  if (k < K) {
    float s = 0.f;
    for (int i = 0; i < dd; ++i) { const float v = M[(int64_t)k * dd + i]; s = fmaf(v, v, s); }
    nrm = sqrtf(s);
  }
  for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
  if ((threadIdx.x & 31) == 0 && isfinite(nrm)) atomicAdd(out, nrm);
}

// sum and max over the centred centroids (for the accuracy gate and the fp16 scale)
__global__ void centred_stats_kernel(const float* __restrict__ c, const float* __restrict__ shift, int K,
                                     float* __restrict__ out2) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float nrm = 0.f, amax = 0.f;
  for (int j = 0; j < 16; ++j) {
    const float cv = c[(int64_t)k * 16 + j] - shift[j];
    nrm = fmaf(cv, cv, nrm);
    amax = fmaxf(amax, fabsf(cv));
  }
  if (isfinite(nrm)) atomicAdd(&out2[0], nrm);
  if (isfinite(amax)) atomicMax(reinterpret_cast<int*>(&out2[1]), __float_as_int(amax));
}

// split-fp16 natural tables [Kpad, 192] (136 packed columns, column 136 = 1, then zeros) for the gradient kernel.
// The constant column is multiplied by U'_136, which is 0 for every real U (the kernel zero-fills p >= 136) and 1
// in the kernel's unit-weight mode, where it makes t_k = <U', M'_k> = 1 exactly (pythae variant, sum_k w_k b_k).
__global__ void pack_sym_nat_h_kernel(const float* __restrict__ M, int Kpad, float scale, __half* __restrict__ hi_n,
                                      __half* __restrict__ lo_n) {
  const int64_t total = (int64_t)Kpad * 192;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx / 192), p = (int)(idx - (int64_t)k * 192);
    float v = 0.f;
    if (p < 136) {
      int i = 0, base = 0;
      while (p >= base + (16 - i)) { base += 16 - i; ++i; }
      const int j = i + (p - base);
      v = scale * 0.5f * (M[(int64_t)k * 256 + i * 16 + j] + M[(int64_t)k * 256 + j * 16 + i]);
    } else if (p == 136) {
      v = 1.f;
    }
    const __half h = __float2half_rn(v);
    hi_n[idx] = h;
    lo_n[idx] = __float2half_rn(v - __half2float(h));
  }
}

// pythae variant on the tensor path (d == 16, symmetric tables): b_k = sym(M_k) (c_k - shift), TF32 hi / lo,
// transposed [16, Kpad] like ctc_hi / ctc_lo (ref pythae rhvae_sampler.py:160-187: sum_k w_k M_k (c_k - z)
// = sum_k w_k b_k - (G^{-1} - lambda I)(z - shift))
__global__ void pack_pythae_bt_kernel(const float* __restrict__ c, const float* __restrict__ M,
                                      const float* __restrict__ shift, int K, int Kpad, float* __restrict__ bt_hi,
                                      float* __restrict__ bt_lo) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Kpad) return;
  float cc[16];
  for (int j = 0; j < 16; ++j) cc[j] = (k < K) ? c[(int64_t)k * 16 + j] - shift[j] : 0.f;
  for (int i = 0; i < 16; ++i) {
    float b = 0.f;
    if (k < K)
      for (int j = 0; j < 16; ++j)
        b = fmaf(0.5f * (M[(int64_t)k * 256 + i * 16 + j] + M[(int64_t)k * 256 + j * 16 + i]), cc[j], b);
    const float hi = tf32_hi(b);
    bt_hi[(int64_t)i * Kpad + k] = hi;
    bt_lo[(int64_t)i * Kpad + k] = b - hi;
  }
}

// one half-warp per point: lane i forms v_i = B_i - sum_e S_ei zt_e (S = G^{-1} - lambda I: off-diagonal entries from
// the packed G^{-1}, the diagonal from the kernel's lambda-free copy), lane j then contracts
// out_j = (1/T^2) sum_i G_ij v_i.
// Error bound -> fallback list.  B and S are table contractions accurate to ~2^-22 of THEIR size, while v is a
// difference of the two; the rounding dv ~ eps (|B| + |S|_F |zt| / 4) is then multiplied by G.  For a rounding
// vector of random direction |G dv| ~ |G|_F |dv| / 4, so with  err = |G|_F / (4 T^2) * (eps (|B| + |S|_F |zt| / 4)
// + floor |zt|)  (floor: the fp16 weight floor 2^-39 of the forward kernel times the table's rms Frobenius norm) a
// point whose err exceeds rtol |out| + atol goes to the list and is recomputed by pythae_exact_kernel, which forms
// c_k - z per centroid like the reference.  Far from every centroid and at moderate temperatures nothing is listed;
// next to a centroid at small T / small lambda (cond(G^{-1}) * |c| / |c_k - z| large) everything is.
__global__ void pythae_finish_sym_kernel(const float* __restrict__ z, const float* __restrict__ a_packed,
                                         const float* __restrict__ s_diag /* [N,16] diagonal without lambda */,
                                         const float* __restrict__ b, const float* __restrict__ g, int g_is_packed,
                                         const float* __restrict__ shift, int64_t n, float inv_T2, float floor_s,
                                         int* __restrict__ fail_ws /* [0] count, [1..] rows; NULL: no bound */,
                                         float* __restrict__ out) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t p = gid >> 4;
  const int i = (int)(gid & 15);
  const bool live = p < n;
  auto pidx = [](int r, int cidx) {        // packed index of (min, max)
    const int lo = r < cidx ? r : cidx, hi = r < cidx ? cidx : r;
    return lo * 16 - (lo * (lo - 1)) / 2 + (hi - lo);
  };
  float v = 0.f, ns = 0.f, zti = 0.f, bi = 0.f;
  if (live) {
    const float* ap = a_packed + p * kSymCols;
    bi = b[p * 16 + i];
    v = bi;
    zti = z[p * 16 + i] - shift[i];
    for (int e = 0; e < 16; ++e) {
      const float zt = z[p * 16 + e] - shift[e];
      const float m = (e == i) ? s_diag[p * 16 + i] : ap[pidx(e, i)];     // (G^{-1} - lambda I)_ei, lambda never added
      ns = fmaf(m, m, ns);
      v = fmaf(-m, zt, v);
    }
  }
  float o = 0.f, ng = 0.f;
  for (int q = 0; q < 16; ++q) {
    const float vq = __shfl_sync(0xffffffffu, v, q, 16);
    if (live) {
      const float gq = g_is_packed ? g[p * kSymCols + pidx(q, i)] : g[p * 256 + q * 16 + i];   // G^T v: G[q][i] v_q
      ng = fmaf(gq, gq, ng);
      o = fmaf(gq, vq, o);
    }
  }
  o *= inv_T2;
  if (live) out[p * 16 + i] = o;
  if (fail_ws != nullptr) {
    float nb = bi * bi, nz = zti * zti, no = o * o;
#pragma unroll
    for (int sft = 8; sft > 0; sft >>= 1) {
      ns += __shfl_xor_sync(0xffffffffu, ns, sft, 16);
      ng += __shfl_xor_sync(0xffffffffu, ng, sft, 16);
      nb += __shfl_xor_sync(0xffffffffu, nb, sft, 16);
      nz += __shfl_xor_sync(0xffffffffu, nz, sft, 16);
      no += __shfl_xor_sync(0xffffffffu, no, sft, 16);
    }
    if (live && i == 0) {
      const float eps = 2.3841858e-7f;                      // 2^-22
      const float zn = sqrtf(nz);
      const float dv = eps * (sqrtf(nb) + 0.25f * sqrtf(ns) * zn) + floor_s * zn;
      const float err = 0.25f * sqrtf(ng) * dv * inv_T2;
      if (!(err <= 2.0e-5f * sqrtf(no) + 1.0e-6f)) {         // also catches NaN
        const int slot = atomicAdd(fail_ws, 1);
        fail_ws[1 + slot] = (int)p;
      }
    }
  }
}

// split-fp16 packed-transposed tables [144, Kpad]: hi = fp16(scale * M), lo = fp16(scale * M - hi)
__global__ void pack_sym_h_kernel(const float* __restrict__ M, int Kpad, float scale, __half* __restrict__ hi_t,
                                  __half* __restrict__ lo_t) {
  const int64_t total = (int64_t)kSymCols * Kpad;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(idx / Kpad), k = (int)(idx - (int64_t)p * Kpad);
    float v = 0.f;
    if (p < 136) {
      int i = 0, base = 0;
      while (p >= base + (16 - i)) { base += 16 - i; ++i; }
      const int j = i + (p - base);
      v = scale * 0.5f * (M[(int64_t)k * 256 + i * 16 + j] + M[(int64_t)k * 256 + j * 16 + i]);
    }
    const __half h = __float2half_rn(v);
    hi_t[idx] = h;
    lo_t[idx] = __float2half_rn(v - __half2float(h));
  }
}

// symmetric tables: packed upper triangle (row p = (i<=j)), transposed [144, Kpad], TF32 hi/lo
__global__ void pack_sym_kernel(const float* __restrict__ M, int Kpad, float* __restrict__ hi_t,
                                float* __restrict__ lo_t) {
  const int64_t total = (int64_t)kSymCols * Kpad;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(idx / Kpad), k = (int)(idx - (int64_t)p * Kpad);
    float v = 0.f;
    if (p < 136) {
      int i = 0, base = 0;                 // invert p = 16 i - i (i-1)/2 + (j - i)
      while (p >= base + (16 - i)) { base += 16 - i; ++i; }
      const int j = i + (p - base);
      v = 0.5f * (M[(int64_t)k * 256 + i * 16 + j] + M[(int64_t)k * 256 + j * 16 + i]);
    }
    const float hi = tf32_hi(v);
    hi_t[idx] = hi;
    lo_t[idx] = v - hi;
  }
}

// packed natural layout [Kpad, 160] for the gradient kernel (row = centroid, 136 packed columns + 0s)
__global__ void pack_sym_nat_kernel(const float* __restrict__ M, int Kpad, float* __restrict__ hi_n,
                                    float* __restrict__ lo_n) {
  const int64_t total = (int64_t)Kpad * kSymNatCols;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx / kSymNatCols), p = (int)(idx - (int64_t)k * kSymNatCols);
    float v = 0.f;
    if (p < 136) {
      int i = 0, base = 0;
      while (p >= base + (16 - i)) { base += 16 - i; ++i; }
      const int j = i + (p - base);
      v = 0.5f * (M[(int64_t)k * 256 + i * 16 + j] + M[(int64_t)k * 256 + j * 16 + i]);
    }
    const float hi = tf32_hi(v);
    hi_n[idx] = hi;
    lo_n[idx] = v - hi;
  }
}

// One warp per matrix: M_k counts as symmetric when |M_ij - M_ji| <= 2^-22 max|M_k| for every pair.
// Tables made as L L^T in fp32 (the reference's metric.pt) are symmetric only up to rounding (~5e-9
// relative); the packed tables then hold (M_ij + M_ji)/2, which moves G^{-1} by < 1.2e-7 relative.
__global__ void symmetry_kernel(const float* __restrict__ m_in, int K, int d, int* __restrict__ asym) {
  const int warps_per_block = blockDim.x >> 5, lane = threadIdx.x & 31;
  const int dd = d * d;
  for (int k = blockIdx.x * warps_per_block + (threadIdx.x >> 5); k < K; k += gridDim.x * warps_per_block) {
    const float* m = m_in + (int64_t)k * dd;
    float amax = 0.f;
    for (int i = lane; i < dd; i += 32) amax = fmaxf(amax, fabsf(m[i]));
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    const float tol = 2.384185791015625e-07f * amax;
    bool bad = false;
    for (int i = lane; i < dd; i += 32) {
      const int r = i / d, c = i - r * d;
      if (r < c && !(fabsf(m[i] - m[c * d + r]) <= tol)) bad = true;
    }
    if (bad) atomicExch(asym, 1);
  }
}

__global__ void negate_copy_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = -x[i];
}
static int negate_copy(const float* x, float* y, int64_t n, cudaStream_t s) {
  negate_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, y, n);
  RLVAE_LAUNCH_OK();
  return 0;
}

static void free_tables(rlvae_tables* t) {
  if (t->Mh_hi) cudaFree(t->Mh_hi);
  if (t->Mh_lo) cudaFree(t->Mh_lo);
  if (t->Mnh_hi) cudaFree(t->Mnh_hi);
  if (t->Mnh_lo) cudaFree(t->Mnh_lo);
  if (t->c64h) cudaFree(t->c64h);
  if (t->ct64_hi) cudaFree(t->ct64_hi);
  if (t->ct64_lo) cudaFree(t->ct64_lo);
  t->ct64_hi = t->ct64_lo = nullptr;
  if (t->c16h) cudaFree(t->c16h);
  if (t->cbias_h) cudaFree(t->cbias_h);
  if (t->cshift) cudaFree(t->cshift);
  if (t->ctc_hi) cudaFree(t->ctc_hi);
  if (t->ctc_lo) cudaFree(t->ctc_lo);
  if (t->bt_hi) cudaFree(t->bt_hi);
  if (t->bt_lo) cudaFree(t->bt_lo);
  t->bt_hi = t->bt_lo = nullptr;
  t->cbias_h = t->cshift = t->ctc_hi = t->ctc_lo = nullptr;
  t->c64h = t->c16h = nullptr;
  t->Mh_hi = t->Mh_lo = t->Mnh_hi = t->Mnh_lo = nullptr;
  float** ptrs[] = {&t->c, &t->cn, &t->M, &t->cstack, &t->cbias, &t->Mt_hi, &t->Mt_lo,
                    &t->Mn_hi, &t->Mn_lo, &t->ct_hi, &t->ct_lo, &t->cn_inf, &t->cmask, &t->Mts_hi, &t->Mts_lo, &t->Mns_hi, &t->Mns_lo};
  for (float** p : ptrs) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
}

}  // namespace rlvae

using namespace rlvae;

extern "C" {

const char* rlvae_last_error(void) { return g_last_error.c_str(); }
int rlvae_abi_version(void) { return 2; }
long long rlvae_launch_count(int reset) {
  return reset ? g_launch_count.exchange(0) : g_launch_count.load();
}

int rlvae_profile_begin(int max_records) {
  RLVAE_REQUIRE(max_records >= 1 && max_records <= 4096, "profile_begin: max_records must be in [1,4096]");
  rlvae_profile_end();
  g_prof.ev = static_cast<cudaEvent_t*>(malloc(sizeof(cudaEvent_t) * 4 * (size_t)max_records));
  RLVAE_REQUIRE(g_prof.ev != nullptr, "profile_begin: out of host memory");
  for (int i = 0; i < 4 * max_records; ++i) RLVAE_CUDA_OK(cudaEventCreate(&g_prof.ev[i]));
  g_prof.cap = max_records;
  g_prof.n = 0;
  g_prof.on = true;
  return 0;
}

int rlvae_profile_count(void) { return g_prof.n; }

int rlvae_profile_read(int record, float ms[3]) {
  RLVAE_REQUIRE(g_prof.ev != nullptr && record >= 0 && record < g_prof.n && ms != nullptr,
                "profile_read: no such record");
  cudaEvent_t* e = g_prof.ev + 4 * record;
  RLVAE_CUDA_OK(cudaEventSynchronize(e[3]));
  for (int i = 0; i < 3; ++i) RLVAE_CUDA_OK(cudaEventElapsedTime(&ms[i], e[i], e[i + 1]));
  return 0;
}

int rlvae_profile_end(void) {
  if (g_prof.ev != nullptr) {
    for (int i = 0; i < 4 * g_prof.cap; ++i) cudaEventDestroy(g_prof.ev[i]);
    free(g_prof.ev);
  }
  g_prof.ev = nullptr;
  g_prof.cap = g_prof.n = 0;
  g_prof.on = false;
  return 0;
}

int rlvae_tables_create(rlvae_tables_t** out, const float* centroids, const float* matrices,
                        int n_centroids, int latent_dim, float temperature, float regularization,
                        void* stream) {
  RLVAE_REQUIRE(out != nullptr, "tables_create: out is NULL");
  *out = nullptr;
  RLVAE_REQUIRE(n_centroids >= 1, "tables_create: need at least one centroid");
  RLVAE_REQUIRE(latent_dim >= 1 && latent_dim <= kMaxLatentDim, "tables_create: latent_dim must be in [1,64]");
  RLVAE_REQUIRE(centroids != nullptr && matrices != nullptr, "tables_create: NULL table pointer");
  RLVAE_REQUIRE(temperature != 0.f, "tables_create: temperature must be non-zero");
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  rlvae_tables* t = new (std::nothrow) rlvae_tables();
  RLVAE_REQUIRE(t != nullptr, "tables_create: out of host memory");
  const int K = n_centroids, d = latent_dim, dd = d * d;
  const int Kpad = (K + kKPad - 1) / kKPad * kKPad;
  t->K = K; t->d = d; t->Kpad = Kpad;
  t->T = temperature; t->T2 = temperature * temperature; t->lambda = regularization;
  const bool tc = (d == 16);

  float* stats = nullptr;  // [0] sum ||c||^2, [1] max ||c||^2, [2] (int) asymmetric flag, [3] max |M|, [4] max |c|
#define ALLOC(ptr, elems)                                                          \
  do {                                                                             \
    cudaError_t _e = cudaMalloc(&(ptr), sizeof(float) * (size_t)(elems));          \
    if (_e != cudaSuccess) {                                                       \
      set_error(std::string("tables_create: cudaMalloc: ") + cudaGetErrorString(_e)); \
      free_tables(t); delete t; if (stats) cudaFree(stats);                        \
      return 1;                                                                    \
    }                                                                              \
  } while (0)
  ALLOC(stats, 8);
  ALLOC(t->c, (size_t)Kpad * d);
  ALLOC(t->cn, Kpad);
  ALLOC(t->M, (size_t)Kpad * dd);
  if (tc) {
    ALLOC(t->cstack, (size_t)Kpad * 32);
    ALLOC(t->ct_hi, (size_t)Kpad * 32);
    ALLOC(t->ct_lo, (size_t)Kpad * 32);
    ALLOC(t->cbias, Kpad);
    ALLOC(t->cn_inf, Kpad);
    ALLOC(t->cmask, Kpad);
    ALLOC(t->Mt_hi, (size_t)Kpad * dd);
    ALLOC(t->Mt_lo, (size_t)Kpad * dd);
    ALLOC(t->Mn_hi, (size_t)Kpad * dd);
    ALLOC(t->Mn_lo, (size_t)Kpad * dd);
  }
#undef ALLOC
  auto fail = [&](int code) { free_tables(t); delete t; cudaFree(stats); return code; };
#define OK_OR_FAIL(expr)                                                           \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));               \
      return fail(1);                                                              \
    }                                                                              \
  } while (0)
  OK_OR_FAIL(cudaMemsetAsync(stats, 0, 8 * sizeof(float), s));
  const float inv_T2_log2e = 1.4426950408889634f / t->T2;
  pack_centroids_kernel<<<(Kpad + 127) / 128, 128, 0, s>>>(centroids, K, Kpad, d, inv_T2_log2e, t->c,
                                                           t->cn, t->cstack, t->cbias, t->ct_hi, t->ct_lo, t->cn_inf, t->cmask, stats);
  OK_OR_FAIL(cudaGetLastError());
  pack_matrices_kernel<<<592, 256, 0, s>>>(matrices, K, Kpad, dd, t->M, t->Mt_hi, t->Mt_lo, t->Mn_hi,
                                           t->Mn_lo, stats);
  OK_OR_FAIL(cudaGetLastError());
  symmetry_kernel<<<296, 256, 0, s>>>(matrices, K, d, reinterpret_cast<int*>(stats + 2));
  OK_OR_FAIL(cudaGetLastError());
  float h_stats[8];
  OK_OR_FAIL(cudaMemcpyAsync(h_stats, stats, sizeof(h_stats), cudaMemcpyDeviceToHost, s));
  OK_OR_FAIL(cudaStreamSynchronize(s));
  cudaFree(stats);
  stats = nullptr;
  int asym;
  std::memcpy(&asym, &h_stats[2], sizeof(int));
  t->symmetric = asym ? 0 : 1;
  t->r2max = h_stats[1];
  t->m_absmax = h_stats[3];
  t->c_absmax = h_stats[4];
  const float r2mean = h_stats[0] / (float)K;
  // Accuracy gate of the expanded form ||z||^2+||c||^2-2 z.c (DESIGN.md "precision"): its
  // absolute error ~2.5e-7*mean||c||^2 becomes a relative error /T^2 in every weight.  G^{-1} was measured at up to
  // 8x that figure against the CUDA-core path (randomised cases right at the old gate of 2e-6: 1.6e-5), hence 1.2e-6.
  constexpr float kExpandedGate = 1.2e-6f;
  t->tensor_auto = 0;
  if (tc) {
    int rc = tc_build_descriptors(t);
    if (rc != 0) return fail(rc);
    t->tensor_capable = 1;
    const float rel = 2.5e-7f * fmaxf(r2mean, 1.f) / t->T2;
    t->expanded_ok = (rel < kExpandedGate) ? 1 : 0;
    t->tensor_auto = t->expanded_ok;
    if (t->symmetric) {   // 136 instead of 256 accumulated columns
      cudaError_t e1 = cudaMalloc(&t->Mts_hi, sizeof(float) * (size_t)kSymCols * Kpad);
      cudaError_t e2 = cudaMalloc(&t->Mts_lo, sizeof(float) * (size_t)kSymCols * Kpad);
      if (e1 == cudaSuccess) e1 = cudaMalloc(&t->Mns_hi, sizeof(float) * (size_t)kSymNatCols * Kpad);
      if (e2 == cudaSuccess) e2 = cudaMalloc(&t->Mns_lo, sizeof(float) * (size_t)kSymNatCols * Kpad);
      if (e1 != cudaSuccess || e2 != cudaSuccess) {
        set_error("tables_create: cudaMalloc (packed tables) failed");
        return fail(1);
      }
      pack_sym_kernel<<<592, 256, 0, s>>>(t->M, Kpad, t->Mts_hi, t->Mts_lo);
      OK_OR_FAIL(cudaGetLastError());
      pack_sym_nat_kernel<<<592, 256, 0, s>>>(t->M, Kpad, t->Mns_hi, t->Mns_lo);
      OK_OR_FAIL(cudaGetLastError());
      OK_OR_FAIL(cudaStreamSynchronize(s));
      rc = tc_build_sym_descriptors(t);
      if (rc != 0) return fail(rc);
      // split-fp16 tables: M' = 2^e M with max|M'| in (2^13, 2^14]
      if (t->m_absmax > 0.f) {
        int ex = 0;
        frexpf(t->m_absmax, &ex);                    // m_absmax = f * 2^ex, f in [0.5, 1)
        const int e = 14 - ex;
        if (e > -60 && e < 60) {
          cudaError_t e3 = cudaMalloc(&t->Mh_hi, sizeof(__half) * (size_t)kSymCols * Kpad);
          cudaError_t e4 = cudaMalloc(&t->Mh_lo, sizeof(__half) * (size_t)kSymCols * Kpad);
          if (e3 == cudaSuccess) e3 = cudaMalloc(&t->Mnh_hi, sizeof(__half) * (size_t)192 * Kpad);
          if (e4 == cudaSuccess) e4 = cudaMalloc(&t->Mnh_lo, sizeof(__half) * (size_t)192 * Kpad);
          if (e3 != cudaSuccess || e4 != cudaSuccess) {
            set_error("tables_create: cudaMalloc (fp16 tables) failed");
            return fail(1);
          }
          pack_sym_h_kernel<<<592, 256, 0, s>>>(t->M, Kpad, ldexpf(1.f, e), static_cast<__half*>(t->Mh_hi),
                                                static_cast<__half*>(t->Mh_lo));
          OK_OR_FAIL(cudaGetLastError());
          pack_sym_nat_h_kernel<<<592, 256, 0, s>>>(t->M, Kpad, ldexpf(1.f, e), static_cast<__half*>(t->Mnh_hi),
                                                    static_cast<__half*>(t->Mnh_lo));
          OK_OR_FAIL(cudaGetLastError());
          // split-fp16 centroid rows for GEMM1, centred on the mean centroid:
          // c' = 2^ec (c - shift) with max|c'| in [2^13, 2^14)
          if (cudaMalloc(&t->c16h, sizeof(__half) * (size_t)Kpad * 64) != cudaSuccess ||
              cudaMalloc(&t->cbias_h, sizeof(float) * (size_t)Kpad) != cudaSuccess ||
              cudaMalloc(&t->ctc_hi, sizeof(float) * (size_t)Kpad * 16) != cudaSuccess ||
              cudaMalloc(&t->ctc_lo, sizeof(float) * (size_t)Kpad * 16) != cudaSuccess ||
              cudaMalloc(&t->cshift, sizeof(float) * 20) != cudaSuccess) {
            set_error("tables_create: cudaMalloc (fp16 centroid rows) failed");
            return fail(1);
          }
          OK_OR_FAIL(cudaMemsetAsync(t->cshift, 0, sizeof(float) * 20, s));
          const char* ce = getenv("RLVAE_TC_CENTRE");   // "0": keep the table un-centred (A/B only)
          if (ce == nullptr || ce[0] != '0') {
            centroid_mean_kernel<<<1, 256, 0, s>>>(t->c, K, t->cshift);
            OK_OR_FAIL(cudaGetLastError());
          }
          centred_stats_kernel<<<(K + 127) / 128, 128, 0, s>>>(t->c, t->cshift, K, t->cshift + 16);
          OK_OR_FAIL(cudaGetLastError());
          matrix_norm_sum_kernel<<<(K + 127) / 128, 128, 0, s>>>(t->M, K, dd, t->cshift + 18);
          OK_OR_FAIL(cudaGetLastError());
          float h_cs[20];
          OK_OR_FAIL(cudaMemcpyAsync(h_cs, t->cshift, sizeof(h_cs), cudaMemcpyDeviceToHost, s));
          OK_OR_FAIL(cudaStreamSynchronize(s));
          int exc = 0;
          frexpf(h_cs[17] > 0.f ? h_cs[17] : 1.f, &exc);
          int ec = 14 - exc;
          ec = ec > 50 ? 50 : (ec < -50 ? -50 : ec);
          pack_c16h_kernel<<<(Kpad + 127) / 128, 128, 0, s>>>(t->c, t->cshift, K, Kpad, ldexpf(1.f, ec), inv_T2_log2e,
                                                              static_cast<__half*>(t->c16h), t->cbias_h, t->ctc_hi, t->ctc_lo);
          OK_OR_FAIL(cudaGetLastError());
          t->c16_unscale = ldexpf(1.f, -ec);
          if (cudaMalloc(&t->bt_hi, sizeof(float) * (size_t)Kpad * 16) != cudaSuccess ||
              cudaMalloc(&t->bt_lo, sizeof(float) * (size_t)Kpad * 16) != cudaSuccess) {
            set_error("tables_create: cudaMalloc (pythae table) failed");
            return fail(1);
          }
          pack_pythae_bt_kernel<<<(Kpad + 127) / 128, 128, 0, s>>>(t->c, t->M, t->cshift, K, Kpad, t->bt_hi, t->bt_lo);
          OK_OR_FAIL(cudaGetLastError());
          OK_OR_FAIL(cudaStreamSynchronize(s));
          {   // the gate of the expanded form now looks at the centred norms
            const float r2c = h_cs[16] / (float)K;
            const float relc = 2.5e-7f * fmaxf(r2c, 1.f) / t->T2;
            t->expanded_ok = (relc < kExpandedGate) ? 1 : 0;
            // Hybrid mode: the un-refined weights (each below 2^-bits, relative error relc) move G^{-1} by
            // at most relc * 2^-bits * sum_k ||M_k||_F <= 1e-6 lambda, and G^{-1} >= lambda I.
            const float mf = h_cs[18];
            t->m_fro_rms = (mf > 0.f && isfinite(mf)) ? mf / sqrtf((float)K) : 0.f;
            if (t->lambda > 0.f && isfinite(t->lambda) && mf > 0.f && isfinite(mf)) {
              const float bits = log2f(relc * mf / (1.0e-6f * t->lambda));
              if (isfinite(bits) && bits < 48.f) {
                t->hybrid_ok = 1;
                t->hybrid_bits = fmaxf(bits, 8.f);
              }
            }
          }
          t->h16_out_scale = ldexpf(1.f, -(14 + e));
          t->h16_m_unscale = ldexpf(1.f, -e);
          t->tensor_auto = 1;    // the split-fp16 kernels have an exact-distance mode: no restriction on T
          rc = tc_build_h16_descriptors(t);
          if (rc != 0) return fail(rc);
          rc = tc_build_ct_centred_descriptors(t);
          if (rc != 0) return fail(rc);
          // Positive semi-definiteness certificate: eigenvalues of every (symmetrised) M_k by the batched
          // Jacobi kernel.  With all M_k >= 0 and lambda > 0, G^{-1}(z) = sum_k w_k M_k + lambda I is
          // positive definite for every z, so the fused kernels' Cholesky cannot fail except by rounding.
          if (t->lambda > 0.f && isfinite(t->lambda)) {
            float* eig = nullptr;
            if (cudaMalloc(&eig, sizeof(float) * (size_t)K * 16) == cudaSuccess) {
              bool okc = launch_sym16_eigvalsh(t->M, K, 0, eig, s) == 0;
              float* h_eig = okc ? static_cast<float*>(malloc(sizeof(float) * (size_t)K * 16)) : nullptr;
              if (h_eig != nullptr &&
                  cudaMemcpyAsync(h_eig, eig, sizeof(float) * (size_t)K * 16, cudaMemcpyDeviceToHost, s) == cudaSuccess &&
                  cudaStreamSynchronize(s) == cudaSuccess) {
                float emin = 0.f;
                bool finite = true;
                for (int64_t i = 0; i < (int64_t)K * 16; i += 16) {     // ascending: entry 0 is the smallest
                  if (!std::isfinite(h_eig[i])) finite = false;
                  emin = fminf(emin, h_eig[i]);
                }
                t->psd_certified = (finite && emin >= -1.0e-6f * t->m_absmax) ? 1 : 0;
              }
              free(h_eig);
              cudaFree(eig);
            }
          }
        }
      }
    }
  }
  if (d == 64 && t->symmetric) {   // split-fp16 forward path for the large-metric configuration
    int rc = tc_build_h64_tables(t, s);
    if (rc != 0) return fail(rc);
    if (t->c64h != nullptr) {
      t->tensor_capable = 1;
      const float rel = 2.5e-7f * fmaxf(t->r2mean_centred, 1.f) / t->T2;     // centred table, as for d = 16
      t->expanded_ok = (rel < kExpandedGate) ? 1 : 0;
      t->tensor_auto = t->expanded_ok;
    }
  }
  *out = t;
  return 0;
}

int rlvae_tables_destroy(rlvae_tables_t* t) {
  if (t == nullptr) return 0;
  free_tables(t);
  delete t;
  return 0;
}

int rlvae_tables_info(const rlvae_tables_t* t, int64_t info[12]) {
  RLVAE_REQUIRE(t != nullptr && info != nullptr, "tables_info: NULL argument");
  info[0] = t->K; info[1] = t->d; info[2] = t->Kpad; info[3] = t->symmetric;
  info[4] = t->tensor_capable; info[5] = t->tensor_auto; info[6] = t->expanded_ok;
  info[7] = (t->d == 16 && t->symmetric && t->c16h != nullptr) ? h16_mode(t) : 0;
  info[8] = t->psd_certified;
  info[9] = info[10] = info[11] = 0;
  return 0;
}

static int resolve_path(const rlvae_tables* t, int path, bool* use_tc) {
  if (path == RLVAE_PATH_DIRECT) { *use_tc = false; return 0; }
  if (path == RLVAE_PATH_TENSOR) {
    RLVAE_REQUIRE(t->tensor_capable, "tensor path requested but not available (latent_dim must be 16)");
    *use_tc = true; return 0;
  }
  RLVAE_REQUIRE(path == RLVAE_PATH_AUTO, "unknown path selector");
  *use_tc = t->tensor_capable && t->tensor_auto;
  return 0;
}

// RLVAE_TC_FWD=tf32 selects the 3xTF32 forward kernel for symmetric tables (default: split fp16)
static bool use_h16(const rlvae_tables* t) {
  static const int v = [] {
    const char* e = getenv("RLVAE_TC_FWD");
    return (e != nullptr && e[0] == 't') ? 0 : 1;
  }();   // initialised once, thread-safe (C++11 magic static)
  return v == 1 && t->Mh_hi != nullptr;
}

// Symmetric tables on the tensor path: packed G^{-1} [N,144] into a_packed plus any of
// { packed G, lad_scale * log|det G^{-1}|, sign, diag(G) }.  fail_ws: 1 + n ints.
// a_full (optional): the expanded [N,16,16] G^{-1} as well (same kernel on the split-fp16 path).
static int sym_forward(const rlvae_tables* t, const float* z, int64_t n, float* a_packed, float* g_packed,
                       float* lad, float lad_scale, float* sgn, float* diag, int* fail_ws, cudaStream_t s,
                       float* a_full = nullptr, float* g_full = nullptr, int a_packed_wanted = 1,
                       float* s_diag = nullptr /* split-fp16 path only: diagonal of sum_k w_k M_k without lambda */) {
  if (use_h16(t)) {
    // a pivoting fallback can only write the packed G: expand it afterwards for the (rare) failures by
    // keeping g_packed alongside g_full
    return launch_inverse_metric_h16(t, z, n, a_packed, g_packed, lad, lad_scale, sgn, diag, fail_ws, s, a_full,
                                     g_full, a_packed_wanted, s_diag);
  }
  RLVAE_REQUIRE(s_diag == nullptr, "the lambda-free diagonal is an output of the split-fp16 kernel only");
  RLVAE_REQUIRE(a_packed != nullptr, "symmetric 3xTF32 path needs the packed buffer");
  if (int rc = launch_inverse_metric_tc_sym(t, z, n, a_packed, s)) return rc;
  if (a_full != nullptr) { if (int rc = launch_unpack_sym16(a_packed, n, a_full, s)) return rc; }
  if (g_packed || lad || sgn || diag) {
    if (int rc = launch_sym16_inverse(a_packed, n, g_packed, lad, lad_scale, sgn, diag, fail_ws, s)) return rc;
  }
  if (g_full != nullptr) return launch_unpack_sym16(g_packed, n, g_full, s);
  return 0;
}

// does this (tables, path selector) pair run the packed symmetric tensor kernels?
static int sym_tensor_path(const rlvae_tables* t, int path, bool* yes) {
  bool use_tc;
  if (int rc = resolve_path(t, path, &use_tc)) return rc;
  *yes = use_tc && t->symmetric && t->Mts_hi != nullptr;
  return 0;
}

static int inverse_metric_full(const rlvae_tables* t, const float* z, int64_t n, float* buf, int path,
                               cudaStream_t s, float* scratch);

int64_t rlvae_inverse_metric_workspace(int64_t n, int d) {
  return d == 16 ? (int64_t)sizeof(float) * n * kSymCols : (d == 64 ? (int64_t)sizeof(float) * n * kSym64Cols : 0);
}

int rlvae_inverse_metric(const rlvae_tables_t* t, const float* z, int64_t n, float* ginv, void* work,
                         int path, void* stream) {
  RLVAE_REQUIRE(t != nullptr, "inverse_metric: tables handle is NULL (metric not loaded)");
  RLVAE_REQUIRE(n >= 0, "inverse_metric: negative batch");
  if (n == 0) return 0;
  RLVAE_REQUIRE(z != nullptr && ginv != nullptr, "inverse_metric: NULL pointer");
  bool use_tc;
  if (int rc = resolve_path(t, path, &use_tc)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!use_tc) return launch_inverse_metric_direct(t, z, n, ginv, s);
  if (t->d == 64) return inverse_metric_full(t, z, n, ginv, path, s, static_cast<float*>(work));
  if (t->symmetric && t->Mts_hi != nullptr && (work != nullptr || use_h16(t))) {
    // symmetric tables: accumulate the 136 packed entries; the split-fp16 kernel expands them to
    // [N,16,16] in its epilogue, the 3xTF32 kernel goes through the packed workspace
    float* packed = use_h16(t) ? nullptr : static_cast<float*>(work);
    return sym_forward(t, z, n, packed, nullptr, nullptr, 1.f, nullptr, nullptr, nullptr, s, ginv);
  }
  return launch_inverse_metric_tc(t, z, n, ginv, s);
}

int rlvae_inverse_metric_packed(const rlvae_tables_t* t, const float* z, int64_t n, float* packed,
                                void* stream) {
  RLVAE_REQUIRE(t != nullptr, "inverse_metric_packed: tables handle is NULL (metric not loaded)");
  RLVAE_REQUIRE(n >= 0, "inverse_metric_packed: negative batch");
  if (n == 0) return 0;
  RLVAE_REQUIRE(z != nullptr && packed != nullptr, "inverse_metric_packed: NULL pointer");
  RLVAE_REQUIRE(t->tensor_capable && t->symmetric && t->Mts_hi != nullptr,
                "inverse_metric_packed: needs latent_dim == 16 and symmetric metric matrices");
  return sym_forward(t, z, n, packed, nullptr, nullptr, 1.f, nullptr, nullptr, nullptr,
                     static_cast<cudaStream_t>(stream));
}

// full [N,d,d] G^{-1} for the non-symmetric / non-tensor cases
// `scratch` (d == 64 only): n * kSym64Cols floats for the packed tiles of the tensor kernel; without
// it AUTO falls back to the direct kernel (an explicit TENSOR request then fails).
static int inverse_metric_full(const rlvae_tables* t, const float* z, int64_t n, float* buf, int path,
                               cudaStream_t s, float* scratch) {
  bool use_tc;
  if (int rc = resolve_path(t, path, &use_tc)) return rc;
  if (use_tc && t->d == 64) {
    if (scratch != nullptr) return launch_inverse_metric_h64(t, z, n, buf, scratch, s);
    RLVAE_REQUIRE(path != RLVAE_PATH_TENSOR, "d = 64 tensor path needs a workspace");
    use_tc = false;
  }
  if (!use_tc) return launch_inverse_metric_direct(t, z, n, buf, s);
  return launch_inverse_metric_tc(t, z, n, buf, s);
}

int rlvae_batched_inverse(const float* a, int64_t n, int d, float* inv, float* logabsdet, float* sign,
                          float* diag_inv, int flags, void* stream) {
  RLVAE_REQUIRE(n >= 0, "batched_inverse: negative batch");
  if (n == 0) return 0;
  RLVAE_REQUIRE(a != nullptr, "batched_inverse: NULL input");
  return launch_batched_inverse(a, n, d, inv, logabsdet, sign, diag_inv, flags & 1,
                                static_cast<cudaStream_t>(stream));
}

int64_t rlvae_metric_grad_workspace(int64_t n, int d) {
  return d == 64 ? (int64_t)sizeof(float) * metric_grad_h64_scratch_floats(n) : 0;
}

int rlvae_metric_grad_ws(const rlvae_tables_t* t, const float* z, const float* u, int64_t n, float scale,
                         float* out, void* work, int path, void* stream) {
  RLVAE_REQUIRE(t != nullptr, "metric_grad: tables handle is NULL (metric not loaded)");
  RLVAE_REQUIRE(n >= 0, "metric_grad: negative batch");
  if (n == 0) return 0;
  RLVAE_REQUIRE(z != nullptr && u != nullptr && out != nullptr, "metric_grad: NULL pointer");
  bool use_tc;
  if (int rc = resolve_path(t, path, &use_tc)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (use_tc && t->d == 16) return launch_metric_grad_tc(t, z, u, n, scale, out, s);
  if (use_tc && t->d == 64 && work != nullptr && metric_grad_h64_available(t))
    return launch_metric_grad_h64(t, z, u, n, scale, out, static_cast<float*>(work), s);
  return launch_metric_grad_direct(t, z, u, n, scale, out, s);
}

int rlvae_metric_grad(const rlvae_tables_t* t, const float* z, const float* u, int64_t n, float scale,
                      float* out, int path, void* stream) {
  return rlvae_metric_grad_ws(t, z, u, n, scale, out, nullptr, path, stream);    // d = 64 without a workspace: direct kernel
}

int64_t rlvae_metric_grad_pythae_workspace(int64_t n, int d) {
  return (int64_t)sizeof(float) * n * (d * d + d);
}

static bool pythae_tensor_available(const rlvae_tables* t) {
  return t->d == 16 && t->symmetric && t->tensor_capable && use_h16(t) && t->Mnh_hi != nullptr && t->bt_hi != nullptr;
}

// variant C on the tensor path: out = (1/T^2) G^T (B - (A - lambda I)(z - shift)), A packed [N,144] G^{-1},
// B [N,16] = sum_k w_k b_k, G expanded [N,16,16] (g_is_packed == 0) or packed [N,144]; rows whose error bound is too
// large are appended to fail_ws (count zeroed here) for launch_pythae_exact
static int launch_pythae_finish_sym(const rlvae_tables* t, const float* z, const float* a_packed, const float* s_diag,
                                    const float* b, const float* g, int g_is_packed, int64_t n, int* fail_ws,
                                    float* out, cudaStream_t s) {
  if (n == 0) return 0;
  if (fail_ws != nullptr) RLVAE_CUDA_OK(cudaMemsetAsync(fail_ws, 0, sizeof(int), s));
  const int64_t threads = n * 16;
  const float floor_s = 1.8189894e-12f * t->m_fro_rms;      // 2^-39: absolute floor of the forward kernel's fp16 weights
  pythae_finish_sym_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(z, a_packed, s_diag, b, g, g_is_packed,
                                                                             t->cshift, n, 1.f / t->T2, floor_s, fail_ws,
                                                                             out);
  RLVAE_LAUNCH_OK();
  return 0;
}

// batches up to this size skip the table-contraction form: the per-centroid kernel with the centroids split over
// several CTAs per point is both faster (a single 128-point tile would walk the centroids serially) and as accurate
// as the reference
constexpr int64_t kPythaeSmallBatch = kPythaeSplitBatch;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int rlvae_metric_grad_pythae(const rlvae_tables_t* t, const float* z, const float* g, int64_t n,
                             float* out, void* work, int path, void* stream) {
  RLVAE_REQUIRE(t != nullptr, "metric_grad_pythae: tables handle is NULL");
  RLVAE_REQUIRE(n >= 0, "metric_grad_pythae: negative batch");
  if (n == 0) return 0;
  RLVAE_REQUIRE(z != nullptr && g != nullptr && out != nullptr, "metric_grad_pythae: NULL pointer");
  RLVAE_REQUIRE(work != nullptr, "metric_grad_pythae: workspace required");
  bool use_tc;
  if (int rc = resolve_path(t, path, &use_tc)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* w = static_cast<float*>(work);
  if (use_tc && n > kPythaeSmallBatch && pythae_tensor_available(t) && aligned16(z) && aligned16(w)) {
    // tensor path: packed G^{-1} from the forward kernel, sum_k w_k b_k from the gradient kernel's unit-weight
    // mode, then the two 16 x 16 products per point; the rows its error bound flags are redone per centroid.
    // Workspace: A [n,144] | B [n,16] | lambda-free diagonal [n,16] | fallback list (1 + n ints)
    float* a_packed = w;
    float* b = w + n * kSymCols;
    float* sd = b + n * 16;
    int* fail_ws = reinterpret_cast<int*>(sd + n * 16);
    if (int rc = sym_forward(t, z, n, a_packed, nullptr, nullptr, 1.f, nullptr, nullptr, nullptr, s, nullptr, nullptr, 1, sd))
      return rc;
    if (int rc = launch_metric_grad_h16(t, z, nullptr, n, 1.f, b, s, 2)) return rc;
    if (int rc = launch_pythae_finish_sym(t, z, a_packed, sd, b, g, 0, n, fail_ws, out, s)) return rc;
    return launch_pythae_exact(t, z, g, 0, fail_ws + 1, fail_ws, n, out, nullptr, s);
  }
  // (the workspace holds n * (d*d + d) floats >= n * kPythaeMaxSplits * d for every d >= 16; smaller d: no split)
  return launch_pythae_exact(t, z, g, 0, nullptr, nullptr, n, out, t->d >= kPythaeMaxSplits ? w : nullptr, s);
}

// variant C in one call (what one leapfrog step of the pythae sampler needs): log|det G^{-1}|, its sign and
// (1/T^2) G^T sum_k w_k M_k^T (c_k - z).  Tensor path, long batches: forward kernel (packed G^{-1}, packed G, log det)
// + unit-weight gradient kernel + finish (error bound) + the per-centroid kernel over the flagged rows; short
// batches: forward kernel + the per-centroid kernel; other tables: the CUDA-core kernels.
int64_t rlvae_pythae_eval_workspace(int64_t n, int d) {
  return (int64_t)sizeof(float) * (n * (3 * (int64_t)d * d + d) + 4);
}

int rlvae_pythae_eval(const rlvae_tables_t* t, const float* z, int64_t n, float* grad, float* logabsdet,
                      float* sign, void* work, int path, void* stream) {
  RLVAE_REQUIRE(t != nullptr, "pythae_eval: tables handle is NULL (metric not loaded)");
  RLVAE_REQUIRE(n >= 0, "pythae_eval: negative batch");
  if (n == 0) return 0;
  RLVAE_REQUIRE(z != nullptr && grad != nullptr, "pythae_eval: NULL pointer");
  RLVAE_REQUIRE(work != nullptr, "pythae_eval: workspace required");
  bool use_tc;
  if (int rc = resolve_path(t, path, &use_tc)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* w = static_cast<float*>(work);
  const int d = t->d;
  if (use_tc && pythae_tensor_available(t) && aligned16(z) && aligned16(w) && aligned16(grad)) {
    float* a_packed = w;
    float* g_packed = w + n * kSymCols;
    float* b = g_packed + n * kSymCols;
    float* sd = b + n * 16;
    int* fail_ws = reinterpret_cast<int*>(sd + n * 16);
    const bool split = n > kPythaeSmallBatch;
    if (int rc = sym_forward(t, z, n, a_packed, g_packed, logabsdet, 1.f, sign, nullptr, fail_ws, s, nullptr, nullptr, 1,
                             split ? sd : nullptr))
      return rc;
    if (!split) return launch_pythae_exact(t, z, g_packed, 1, nullptr, nullptr, n, grad, b /* 463 n floats free */, s);
    if (int rc = launch_metric_grad_h16(t, z, nullptr, n, 1.f, b, s, 2)) return rc;
    if (int rc = launch_pythae_finish_sym(t, z, a_packed, sd, b, g_packed, 1, n, fail_ws, grad, s)) return rc;
    return launch_pythae_exact(t, z, g_packed, 1, fail_ws + 1, fail_ws, n, grad, nullptr, s);
  }
  const int64_t mat = n * d * d;
  float* ginv = w;
  float* g = w + mat;
  float* scratch = w + 2 * mat;
  // (d = 64 tensor forward: its packed tiles live in the third slot of the workspace)
  if (int rc = inverse_metric_full(t, z, n, ginv, path, s, scratch)) return rc;
  if (d == 64 && t->symmetric) {       // SPD sweep; the third slot is free again and holds its fallback list
    if (int rc = launch_spd64(ginv, n, g, logabsdet, 1.f, reinterpret_cast<int*>(scratch), s, sign, nullptr)) return rc;
  } else if (int rc = launch_batched_inverse(ginv, n, d, g, logabsdet, sign, nullptr, 0, s)) {
    return rc;
  }
  // (the third slot, n * (d*d + d) floats, is free again: partial sums of the split-centroid launch)
  return launch_pythae_exact(t, z, g, 0, nullptr, nullptr, n, grad, d >= kPythaeMaxSplits ? scratch : nullptr, s);
}

int64_t rlvae_metric_eval_workspace(int64_t n, int d) {
  return (int64_t)sizeof(float) * (3 * n * d * d + n + 4);
}

int rlvae_metric_eval(const rlvae_tables_t* t, const float* z, int64_t n, float* ginv, float* g,
                      float* logdet_g, float* grad_logdet_g, void* work, int path, void* stream) {
  RLVAE_REQUIRE(t != nullptr, "metric_eval: tables handle is NULL (metric not loaded)");
  RLVAE_REQUIRE(n >= 0, "metric_eval: negative batch");
  if (n == 0) return 0;
  RLVAE_REQUIRE(z != nullptr, "metric_eval: z is NULL");
  RLVAE_REQUIRE(work != nullptr, "metric_eval: workspace required");
  const int d = t->d;
  const int64_t mat = n * d * d;
  float* w = static_cast<float*>(work);
  float* a_buf = w;                    // G^{-1}, full or packed (both fit in n*d*d floats)
  float* g_buf = g ? g : (w + mat);
  float* lad_buf = w + 2 * mat;
  float* gt_buf = w + 2 * mat + ((n + 3) & ~(int64_t)3);   // 16-byte aligned (also the d = 64 packed scratch)
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bool packed = false;
  if (int rc = sym_tensor_path(t, path, &packed)) return rc;
  if (packed) {
    // symmetric tables on the tensor path: ONE kernel produces packed G^{-1} [N,144], packed G and
    // -log|det G^{-1}| = log|det G| (per-thread Cholesky fused into its epilogue); the gradient
    // kernel contracts the packed G directly (G^T == G).  The spare tail of a_buf is the fallback list.
    float* g_packed = (g != nullptr || grad_logdet_g != nullptr) ? (w + mat) : nullptr;
    int* fail_ws = reinterpret_cast<int*>(a_buf + n * kSymCols);
    if (int rc = sym_forward(t, z, n, a_buf, g_packed, logdet_g, -1.f, nullptr, nullptr, fail_ws, s, ginv, g, 0))
      return rc;
    prof_mark(2, s);       // (0, 1 bracket the forward launch inside launch_inverse_metric_h16)
    int rc = 0;
    if (grad_logdet_g != nullptr)
      rc = launch_metric_grad_tc(t, z, g_packed, n, -2.f / t->T2, grad_logdet_g, s, 1);
    prof_mark(3, s);
    prof_next();
    return rc;
  }
  if (ginv != nullptr) a_buf = ginv;            // the caller's buffer is the working copy: no device-to-device copy
  if (int rc = inverse_metric_full(t, z, n, a_buf, path, s, gt_buf)) return rc;
  // the gradient contracts M_k with G^T (d log det A = tr(A^{-1} dA)); for symmetric tables
  // G^T == G up to rounding, otherwise a transposed copy is produced.
  const bool need_gt = (grad_logdet_g != nullptr) && !t->symmetric;
  const bool plain_g = (g != nullptr) || ((grad_logdet_g != nullptr) && t->symmetric);
  bool lad_done = false;
  if ((plain_g || logdet_g != nullptr) && d == 64 && t->symmetric) {
    // symmetric tables: G^{-1} is symmetric positive definite -> register-resident elimination / sweep without
    // pivoting or staging (spd64_logdet_kernel / spd64_inverse_kernel); the few matrices that are not PD re-run through the pivoting kernel.
    // (The G^T slot is free here: the packed tiles were consumed by the unpack, the gradient kernel comes later.)
    if (int rc = launch_spd64(a_buf, n, plain_g ? g_buf : nullptr, logdet_g, -1.f, reinterpret_cast<int*>(gt_buf), s))
      return rc;
    lad_done = true;
  } else if (plain_g || logdet_g != nullptr) {
    if (int rc = launch_batched_inverse(a_buf, n, d, plain_g ? g_buf : nullptr, logdet_g ? lad_buf : nullptr,
                                        nullptr, nullptr, 0, s))
      return rc;
  }
  if (need_gt) {
    if (int rc = launch_batched_inverse(a_buf, n, d, gt_buf, nullptr, nullptr, nullptr, 1, s)) return rc;
  }
  if (logdet_g != nullptr && !lad_done) {
    // log|det G| = -log|det G^{-1}|
    if (int rc = negate_copy(lad_buf, logdet_g, n, s)) return rc;
  }
  if (grad_logdet_g != nullptr) {
    // grad_z log det G = -(2/T^2) sum_k w_k tr(G M_k) (c_k - z)
    // (d = 64 tensor kernel: the G^T slot of the workspace is free for symmetric tables -- its partial tiles live there)
    if (int rc = rlvae_metric_grad_ws(t, z, need_gt ? gt_buf : g_buf, n, -2.f / t->T2, grad_logdet_g,
                                      need_gt ? nullptr : gt_buf, path, stream))
      return rc;
  }
  return 0;
}

int rlvae_sym_eigvalsh(const float* a, int64_t n, int d, int packed, float* eig, void* stream) {
  RLVAE_REQUIRE(n >= 0, "sym_eigvalsh: negative batch");
  if (n == 0) return 0;
  RLVAE_REQUIRE(a != nullptr && eig != nullptr, "sym_eigvalsh: NULL pointer");
  RLVAE_REQUIRE(d == 16, "sym_eigvalsh: latent_dim must be 16 (other sizes: use the framework's eigvalsh)");
  RLVAE_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(eig) & 15) == 0,
                "sym_eigvalsh: pointers must be 16-byte aligned");
  return launch_sym16_eigvalsh(a, n, packed ? 1 : 0, eig, static_cast<cudaStream_t>(stream));
}

int rlvae_metric_spectrum(const rlvae_tables_t* t, const float* z, int64_t n, float* eig_ginv,
                          float* logdet_g, void* work, int path, void* stream) {
  RLVAE_REQUIRE(t != nullptr, "metric_spectrum: tables handle is NULL (metric not loaded)");
  RLVAE_REQUIRE(n >= 0, "metric_spectrum: negative batch");
  if (n == 0) return 0;
  RLVAE_REQUIRE(z != nullptr && eig_ginv != nullptr && work != nullptr, "metric_spectrum: NULL pointer");
  RLVAE_REQUIRE(t->d == 16 && t->symmetric, "metric_spectrum: needs latent_dim == 16 and symmetric metric matrices");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* a_buf = static_cast<float*>(work);
  bool packed = false;
  if (int rc = sym_tensor_path(t, path, &packed)) return rc;
  if (packed) {
    int* fail_ws = reinterpret_cast<int*>(a_buf + n * kSymCols);
    if (int rc = sym_forward(t, z, n, a_buf, nullptr, logdet_g, -1.f, nullptr, nullptr, fail_ws, s)) return rc;
    return launch_sym16_eigvalsh(a_buf, n, 1, eig_ginv, s);
  }
  if (int rc = inverse_metric_full(t, z, n, a_buf, path, s, nullptr)) return rc;
  if (logdet_g != nullptr) {
    float* lad = a_buf + n * 256;
    if (int rc = launch_batched_inverse(a_buf, n, 16, nullptr, lad, nullptr, nullptr, 0, s)) return rc;
    if (int rc = negate_copy(lad, logdet_g, n, s)) return rc;
  }
  return launch_sym16_eigvalsh(a_buf, n, 0, eig_ginv, s);
}

int rlvae_local_covariance(const float* latents, int64_t n, const float* centroids, int n_centroids,
                           int latent_dim, float temperature, float* cov, void* stream) {
  RLVAE_REQUIRE(n >= 0 && n_centroids >= 0, "local_covariance: negative size");
  RLVAE_REQUIRE(latent_dim >= 1 && latent_dim <= kMaxLatentDim, "local_covariance: latent_dim must be in [1,64]");
  RLVAE_REQUIRE(temperature != 0.f, "local_covariance: temperature must be non-zero");
  if (n_centroids == 0) return 0;
  RLVAE_REQUIRE(centroids != nullptr && cov != nullptr && (latents != nullptr || n == 0),
                "local_covariance: NULL pointer");
  return launch_local_covariance(latents, n, centroids, n_centroids, latent_dim, temperature, cov,
                                 static_cast<cudaStream_t>(stream));
}

static int64_t hmc_core_floats(int64_t n, int d) {
  // ginv, g (exact mode), diag, rho_half, z_prev, grad, logabsdet, sign, h0
  return 2 * n * d * d + 4 * n * d + 3 * n;
}

int64_t rlvae_hmc_workspace(int64_t n, int d) {
  // + 64 ints: [0] = the fused trajectory kernel's rounding-failure counter (see rlvae_hmc_iteration)
  return (int64_t)sizeof(float) * (hmc_core_floats(n, d) + 64);
}

int64_t rlvae_hmc_run_workspace(int64_t n, int d, int n_iters, int n_lf) {
  return rlvae_hmc_workspace(n, d) + (int64_t)sizeof(float) * (((int64_t)n_iters * n_lf + 63) / 64 * 64);
}

// RLVAE_HMC_FUSED=0 keeps the per-step launches (A/B and debugging); read per call
static bool hmc_fused_enabled() {
  const char* e = getenv("RLVAE_HMC_FUSED");
  return !(e != nullptr && e[0] == '0');
}

static bool hmc_use_fused(const rlvae_tables* t, int grad_mode, int path) {
  if (grad_mode & RLVAE_HMC_NO_FUSION) return false;
  bool packed = false;
  if (sym_tensor_path(t, path, &packed) != 0 || !packed) return false;
  return grad_mode == RLVAE_GRAD_MODULAR && use_h16(t) && h16_hmc_available(t) && hmc_fused_enabled();
}

int rlvae_hmc_fused_available(const rlvae_tables_t* t, int grad_mode, int path) {
  return (t != nullptr && hmc_use_fused(t, grad_mode, path)) ? 1 : 0;
}

int rlvae_hmc_iteration(const rlvae_tables_t* t, float* z, const float* gamma, const float* acc,
                        int64_t n, int n_lf, float eps_lf, float beta_zero_sqrt, const float* h_scales,
                        int grad_mode, float* h0_out, float* h1, float* alpha, float* moves, void* work,
                        int path, void* stream) {
  RLVAE_REQUIRE(t != nullptr, "hmc_iteration: tables handle is NULL (metric not loaded)");
  RLVAE_REQUIRE(n >= 0 && n_lf >= 1, "hmc_iteration: need n >= 0 and n_lf >= 1");
  if (n == 0) return 0;
  RLVAE_REQUIRE(z && gamma && acc && h_scales && work, "hmc_iteration: NULL pointer");
  const int grad_mode_flags = grad_mode;
  grad_mode &= ~RLVAE_HMC_NO_FUSION;
  RLVAE_REQUIRE(grad_mode == RLVAE_GRAD_MODULAR || grad_mode == RLVAE_GRAD_EXACT,
                "hmc_iteration: unknown grad_mode");
  const int d = t->d;
  const int64_t mat = n * d * d, vec = n * d;
  float* w = static_cast<float*>(work);
  float* ginv = w;
  float* gfull = w + mat;
  float* diag = w + 2 * mat;
  float* rho_half = diag + vec;
  float* z_prev = rho_half + vec;
  float* gex = z_prev + vec;
  float* lad = gex + vec;
  float* sgn = lad + n;
  float* h0 = sgn + n;
  const bool exact = grad_mode == RLVAE_GRAD_EXACT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int* fused_fail = reinterpret_cast<int*>(w + hmc_core_floats(n, d));
  RLVAE_CUDA_OK(cudaMemsetAsync(fused_fail, 0, sizeof(int), s));

  bool packed = false;
  if (int rc = sym_tensor_path(t, path, &packed)) return rc;
  if (n_lf <= 64 && hmc_use_fused(t, grad_mode_flags, path)) {
    // ONE launch: every CTA pair keeps its 256 chains on chip for the n_lf + 1 metric evaluations
    return launch_hmc_trajectory_h16(t, z, gamma, acc, n, 1, n_lf, eps_lf, beta_zero_sqrt, nullptr, h_scales,
                                     h0_out, h1, alpha, moves, nullptr, fused_fail, s);
  }
  // one metric evaluation at the chain's current position: G^{-1}, then diag(G)/log|det|
  auto eval = [&](const float* zz) -> int {
    if (packed) {   // fused forward + per-thread Cholesky; the gradient contracts packed G
      int* fail_ws = reinterpret_cast<int*>(ginv + n * kSymCols);
      if (int rc = sym_forward(t, zz, n, ginv, exact ? gfull : nullptr, lad, 1.f, sgn, diag, fail_ws, s, nullptr,
                               nullptr, 0))
        return rc;
      if (exact) return launch_metric_grad_tc(t, zz, gfull, n, 1.f / t->T2, gex, s, 1);
      return 0;
    }
    if (int rc = inverse_metric_full(t, zz, n, ginv, path, s, gfull)) return rc;   // gfull doubles as the d = 64 scratch
    // exact mode wants G^T for the contraction (see rlvae_metric_eval)
    if (d == 64 && t->symmetric) {
      // symmetric tables: G^{-1} is SPD -> no-pivoting sweep (diag G, log det, sign = 1; G^T == G); the gradient
      // slot is free until the contraction below and holds the fallback list
      if (int rc = launch_spd64(ginv, n, exact ? gfull : nullptr, lad, 1.f, reinterpret_cast<int*>(gex), s, sgn, diag))
        return rc;
    } else if (int rc = launch_batched_inverse(ginv, n, d, exact ? gfull : nullptr, lad, sgn, diag, 1, s)) {
      return rc;
    }
    if (exact)  // grad_z 1/2 log det G^{-1} = (1/T^2) sum_k w_k tr(G M_k)(c_k - z)
      if (int rc = rlvae_metric_grad(t, zz, gfull, n, 1.f / t->T2, gex, path, stream)) return rc;
    return 0;
  };

  RLVAE_CUDA_OK(cudaMemcpyAsync(z_prev, z, sizeof(float) * vec, cudaMemcpyDeviceToDevice, s));
  if (int rc = eval(z)) return rc;
  if (int rc = launch_hmc_begin(z_prev, gamma, diag, lad, sgn, exact ? gex : nullptr, n, d,
                                beta_zero_sqrt, eps_lf, t->lambda, t->T2, grad_mode, rho_half, z, h0, s))
    return rc;
  for (int k = 1; k <= n_lf; ++k) {
    if (int rc = eval(z)) return rc;
    if (int rc = launch_hmc_step(diag, lad, sgn, exact ? gex : nullptr, n, d, eps_lf, t->lambda, t->T2,
                                 grad_mode, h_scales[k - 1], k == n_lf, rho_half, z, z_prev, acc, h0, h1,
                                 alpha, moves, z, s))
      return rc;
  }
  if (h0_out != nullptr)
    RLVAE_CUDA_OK(cudaMemcpyAsync(h0_out, h0, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  return 0;
}

int rlvae_hmc_run(const rlvae_tables_t* t, float* z, const float* gammas, const float* accs, int64_t n,
                  int n_iters, int n_lf, float eps_lf, float beta_zero_sqrt, const float* h_scales,
                  int grad_mode, float* h0, float* h1, float* alpha, float* moves, float* z_trace,
                  void* work, int path, void* stream) {
  RLVAE_REQUIRE(t != nullptr, "hmc_run: tables handle is NULL (metric not loaded)");
  RLVAE_REQUIRE(n >= 0 && n_lf >= 1 && n_iters >= 0, "hmc_run: need n >= 0, n_lf >= 1, n_iters >= 0");
  if (n == 0 || n_iters == 0) return 0;
  RLVAE_REQUIRE(z && gammas && accs && h_scales && work, "hmc_run: NULL pointer");
  const int d = t->d;
  float* w = static_cast<float*>(work);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (hmc_use_fused(t, grad_mode, path)) {
    int* fused_fail = reinterpret_cast<int*>(w + hmc_core_floats(n, d));
    float* scales_dev = w + hmc_core_floats(n, d) + 64;
    RLVAE_CUDA_OK(cudaMemsetAsync(fused_fail, 0, sizeof(int), s));
    RLVAE_CUDA_OK(cudaMemcpyAsync(scales_dev, h_scales, sizeof(float) * (size_t)n_iters * n_lf,
                                  cudaMemcpyHostToDevice, s));
    return launch_hmc_trajectory_h16(t, z, gammas, accs, n, n_iters, n_lf, eps_lf, beta_zero_sqrt, scales_dev,
                                     nullptr, h0, h1, alpha, moves, z_trace, fused_fail, s);
  }
  for (int i = 0; i < n_iters; ++i) {
    if (int rc = rlvae_hmc_iteration(t, z, gammas + (int64_t)i * n * d, accs + (int64_t)i * n, n, n_lf, eps_lf,
                                     beta_zero_sqrt, h_scales + (int64_t)i * n_lf, grad_mode,
                                     h0 ? h0 + (int64_t)i * n : nullptr, h1 ? h1 + (int64_t)i * n : nullptr,
                                     alpha ? alpha + (int64_t)i * n : nullptr,
                                     moves ? moves + (int64_t)i * n : nullptr, work, path, stream))
      return rc;
    if (z_trace != nullptr)
      RLVAE_CUDA_OK(cudaMemcpyAsync(z_trace + (int64_t)i * n * d, z, sizeof(float) * (size_t)n * d,
                                    cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

// ---- pythae RHVAESampler.hmc_sampling: the whole loop behind OfficialRHVAESampler.sample_prior ---------------
static int64_t round4(int64_t x) { return (x + 3) & ~(int64_t)3; }

int64_t rlvae_pythae_hmc_workspace(int64_t n, int d) {
  // evaluation workspace | grad, rho_half, z0, g0 [n,d] | lad, sgn, lp0, h0 [n]
  return rlvae_pythae_eval_workspace(n, d) + 16 + (int64_t)sizeof(float) * (4 * round4(n * d) + 4 * round4(n));
}

int rlvae_pythae_hmc_run(const rlvae_tables_t* t, float* z, const float* gammas, const float* accs, int64_t n,
                         int n_iters, int n_lf, float eps_lf, float beta_zero_sqrt, const float* h_scales, float* h0,
                         float* h1, float* alpha, float* moves, float* z_trace, void* work, int path, void* stream) {
  RLVAE_REQUIRE(t != nullptr, "pythae_hmc_run: tables handle is NULL (metric not loaded)");
  RLVAE_REQUIRE(n >= 0 && n_lf >= 1 && n_iters >= 0, "pythae_hmc_run: need n >= 0, n_lf >= 1, n_iters >= 0");
  if (n == 0 || n_iters == 0) return 0;
  RLVAE_REQUIRE(z && gammas && accs && h_scales && work, "pythae_hmc_run: NULL pointer");
  const int d = t->d;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* wb = static_cast<char*>(work);
  void* eval_ws = wb;
  float* f = reinterpret_cast<float*>(wb + ((rlvae_pythae_eval_workspace(n, d) + 15) & ~(int64_t)15));
  float* grad = f;                     f += round4(n * d);
  float* rho = f;                      f += round4(n * d);
  float* z0 = f;                       f += round4(n * d);
  float* g0 = f;                       f += round4(n * d);
  float* lad = f;                      f += round4(n);
  float* sgn = f;                      f += round4(n);
  float* lp0 = f;                      f += round4(n);
  float* h0buf = f;
  if (int rc = rlvae_pythae_eval(t, z, n, grad, lad, sgn, eval_ws, path, stream)) return rc;     // (log_pi, grad) at the start
  for (int i = 0; i < n_iters; ++i) {
    const int64_t on = (int64_t)i * n;
    if (int rc = launch_pythae_hmc_begin(n, d, eps_lf, beta_zero_sqrt, i == 0, lad, sgn, grad, gammas + on * d, z, z0,
                                         rho, g0, lp0, h0buf, h0 ? h0 + on : nullptr, s))
      return rc;
    for (int k = 0; k < n_lf; ++k) {
      if (int rc = rlvae_pythae_eval(t, z, n, grad, lad, sgn, eval_ws, path, stream)) return rc;
      const int last = (k == n_lf - 1);
      if (int rc = launch_pythae_hmc_step(n, d, eps_lf, h_scales[(int64_t)i * n_lf + k], last, lad, sgn, grad, accs + on,
                                          z, z0, rho, g0, lp0, h0buf, h1 ? h1 + on : nullptr,
                                          alpha ? alpha + on : nullptr, moves ? moves + on : nullptr,
                                          z_trace ? z_trace + on * d : nullptr, s))
        return rc;
    }
  }
  return 0;
}

int rlvae_hmc_refine(const rlvae_tables_t* t, float* z, int64_t n, int n_steps, float step_size,
                     void* work, int path, void* stream) {
  RLVAE_REQUIRE(t != nullptr, "hmc_refine: tables handle is NULL (metric not loaded)");
  RLVAE_REQUIRE(n >= 0 && n_steps >= 0, "hmc_refine: bad sizes");
  if (n == 0) return 0;
  RLVAE_REQUIRE(z && work, "hmc_refine: NULL pointer");
  const int d = t->d;
  float* w = static_cast<float*>(work);
  float* ginv = w;
  float* diag = w + 2 * n * d * d;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bool packed = false;
  if (int rc = sym_tensor_path(t, path, &packed)) return rc;
  for (int i = 0; i < n_steps; ++i) {
    if (packed) {
      int* fail_ws = reinterpret_cast<int*>(ginv + n * kSymCols);
      if (int rc = sym_forward(t, z, n, ginv, nullptr, nullptr, 1.f, nullptr, diag, fail_ws, s, nullptr, nullptr, 0))
        return rc;
    } else {
      if (int rc = inverse_metric_full(t, z, n, ginv, path, s, w + n * d * d)) return rc;
      if (int rc = launch_batched_inverse(ginv, n, d, nullptr, nullptr, nullptr, diag, 0, s)) return rc;
    }
    if (int rc = launch_axpy_grad_modular(z, diag, n, d, step_size, t->lambda, t->T2, s)) return rc;
  }
  return 0;
}

int rlvae_nearest2(const rlvae_tables_t* t, const float* mu, int64_t n, int64_t* idx, float* dist,
                   void* stream) {
  RLVAE_REQUIRE(t != nullptr, "nearest2: tables handle is NULL (metric not loaded)");
  RLVAE_REQUIRE(n >= 0, "nearest2: negative batch");
  if (n == 0) return 0;
  RLVAE_REQUIRE(mu && idx && dist, "nearest2: NULL pointer");
  // d == 16: distance GEMM on the tensor core as a pre-filter, exact decision among 8 candidates
  // (same indices and distances as the direct kernel); RLVAE_NEAREST=direct keeps the scalar scan
  static const int use_tc = [] {
    const char* e = getenv("RLVAE_NEAREST");
    return (e != nullptr && e[0] == 'd') ? 0 : 1;
  }();
  if (use_tc && t->d == 16 && t->cn_inf != nullptr && t->K >= 2 && (reinterpret_cast<uintptr_t>(mu) & 15) == 0)
    return launch_nearest2_tc(t, mu, n, idx, dist, static_cast<cudaStream_t>(stream));
  return launch_nearest2(t, mu, n, idx, dist, static_cast<cudaStream_t>(stream));
}

int rlvae_chol_apply(const float* a, const float* eps, int64_t n, int d, float jitter, float* out,
                     int32_t* status, void* stream) {
  RLVAE_REQUIRE(n >= 0, "chol_apply: negative batch");
  if (n == 0) return 0;
  RLVAE_REQUIRE(a && eps && out, "chol_apply: NULL pointer");
  return launch_chol_apply(a, eps, n, d, jitter, out, status, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

