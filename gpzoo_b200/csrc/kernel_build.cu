// K1: fused pairwise-distance + exp kernel-matrix build and its backward.
//
// Replaces  torch.cdist(X,Z)**2 -> expand(L) -> /l^2 -> exp -> *sigma^2   (kernels.py:114-130, 141-155)
// and the multi-group form with group-distance scaling                   (kernels.py:172-191, 204-228):
//
//   K[l,i,j] = sigma_l^2 * exp(-0.5 * |x1_i - x2_j|^2 / (ls_l^2 * den)) / den^(p/2),   den = a_l * r2[g1_i, g2_j] + 1
//
// (den = 1 without groups), and the Matern-3/2 covariance functor of kernels.py:6-30 (batched_Matern32):
//
//   K[l,i,j] = sigma_l^2 (1 + v) exp(-v),   v = sqrt(3) |x1_i - x2_j| / ls_l.
//  d^2 is a direct sum of squared differences (never |x|^2+|z|^2-2xz), which is
// what makes fp32 results track the fp64 reference.  Output layout L x n1 x n2, n2 contiguous.
// HBM-bound: one pass writes 4*L*n1*n2 bytes; each thread owns 4 consecutive j and streams float4 stores.
#include "common.cuh"
#include "gpzoo_b200.h"
#include "umma_gemm.h"

namespace gpz {

constexpr int KB_DMAX = 4;      // spatial dimension supported by the kernels (reference uses D = 1 or 2)
constexpr int KB_LMAX = 32;     // factors handled per launch by the backward (host loops over chunks)
constexpr int KB_ROWS = 16;     // rows (i) per CTA
constexpr int KB_THREADS = 256;
constexpr int KB_VEC = 4;       // consecutive j per thread
constexpr int KB_GMAX = 64;     // max number of groups (r2 table kept in shared memory)

template <typename T> struct KBArgs {
  const T* x1; const T* x2;         // n1 x D, n2 x D
  const T* sigma; const T* ls;      // L
  const T* a;                       // L  group-difference coefficient (already squared / abs'd by the host); MG only
  const T* r2;                      // ng x ng squared embedding distances; MG only
  const int64_t* g1; const int64_t* g2;
  int n1, n2, D, L, ng;
  T p_half;                         // input_dim / 2
  T jitter;                         // added to K[l,i,i] (add_jitter, utilities.py:407-418); 0 for cross matrices
  int kind;                         // 0: RBF (squared exponential); 1: Matern-3/2 (kernels.py:6-30), no groups
};

template <typename T> struct Vec4 { T v[4]; };

template <typename T, bool MG, bool ALIGNED, bool MAT = false>
__global__ void __launch_bounds__(KB_THREADS) kbuild_fwd_kernel(const KBArgs<T> a, T* __restrict__ out, T* __restrict__ out_lo) {
  __shared__ T s_c[KB_LMAX * 8], s_s2[KB_LMAX * 8], s_a[KB_LMAX * 8];
  __shared__ T s_r2[MG ? KB_GMAX * KB_GMAX : 1];
  const int tid = threadIdx.x;
  for (int l = tid; l < a.L; l += KB_THREADS) {
    const T ls = a.ls[l], sg = a.sigma[l];
    s_c[l] = MAT ? Num<T>::sqrt(T(3)) / ls : T(-0.5) / (ls * ls);
    s_s2[l] = sg * sg;
    if (MG) s_a[l] = a.a[l];
  }
  if (MG) for (int e = tid; e < a.ng * a.ng; e += KB_THREADS) s_r2[e] = a.r2[e];
  __syncthreads();

  const int64_t j0 = ((int64_t)blockIdx.x * KB_THREADS + tid) * KB_VEC;
  if (j0 >= a.n2) return;
  const int nv = (int)min((int64_t)KB_VEC, a.n2 - j0);
  T xj[KB_DMAX][KB_VEC];
  int gj[KB_VEC];
#pragma unroll
  for (int v = 0; v < KB_VEC; ++v) {
#pragma unroll
    for (int d = 0; d < KB_DMAX; ++d) xj[d][v] = (v < nv && d < a.D) ? a.x2[(j0 + v) * a.D + d] : T(0);
    gj[v] = (MG && v < nv) ? (int)a.g2[j0 + v] : 0;
  }
  const int i0 = blockIdx.y * KB_ROWS, i1 = min(i0 + KB_ROWS, a.n1);
  for (int i = i0; i < i1; ++i) {
    T d2[KB_VEC], r2[KB_VEC];
#pragma unroll
    for (int v = 0; v < KB_VEC; ++v) d2[v] = T(0);
#pragma unroll
    for (int d = 0; d < KB_DMAX; ++d) {
      if (d < a.D) {
        const T xi = a.x1[(int64_t)i * a.D + d];
#pragma unroll
        for (int v = 0; v < KB_VEC; ++v) { const T df = xi - xj[d][v]; d2[v] = fma(df, df, d2[v]); }
      }
    }
    if (MG) {
      const int gi = (int)a.g1[i];
#pragma unroll
      for (int v = 0; v < KB_VEC; ++v) r2[v] = s_r2[gi * a.ng + gj[v]];
    }
    for (int l = 0; l < a.L; ++l) {
      const T c = s_c[l], s2 = s_s2[l];
      Vec4<T> k;
#pragma unroll
      for (int v = 0; v < KB_VEC; ++v) {
        T val;
        if (MG) {
          // one reciprocal per entry (MUFU.RCP in fp32: the two IEEE divisions held the multi-group build at 20 % of HBM peak)
          const T den = fma(s_a[l], r2[v], T(1));
          const T idn = fast_div(T(1), den);
          const T sc = a.p_half == T(1) ? idn : Num<T>::pow(den, -a.p_half);
          val = s2 * Num<T>::exp(c * d2[v] * idn) * sc;
        } else if (MAT) {
          const T vv = c * Num<T>::sqrt(d2[v]);
          val = s2 * (T(1) + vv) * Num<T>::exp(-vv);
        } else {
          val = s2 * Num<T>::exp(c * d2[v]);
        }
        if (a.jitter != T(0) && (int64_t)i == j0 + v) val += a.jitter;
        k.v[v] = val;
      }
      T* o = out + ((int64_t)l * a.n1 + i) * a.n2 + j0;
      if (ALIGNED && nv == KB_VEC) {
        if (sizeof(T) == 4) {
          __stcs(reinterpret_cast<float4*>(o), make_float4((float)k.v[0], (float)k.v[1], (float)k.v[2], (float)k.v[3]));
        } else {
          __stcs(reinterpret_cast<double2*>(o), make_double2((double)k.v[0], (double)k.v[1]));
          __stcs(reinterpret_cast<double2*>(o) + 1, make_double2((double)k.v[2], (double)k.v[3]));
        }
      } else {
        for (int v = 0; v < nv; ++v) o[v] = k.v[v];
      }
      if (sizeof(T) == 4 && out_lo != nullptr) {          // lo = x - tf32_trunc(x): second operand of the split-TF32 GEMMs
        float lo[KB_VEC];
#pragma unroll
        for (int v = 0; v < KB_VEC; ++v) {
          const float x = (float)k.v[v];
          lo[v] = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        }
        float* ol = reinterpret_cast<float*>(out_lo) + ((int64_t)l * a.n1 + i) * a.n2 + j0;
        if (ALIGNED && nv == KB_VEC) __stcs(reinterpret_cast<float4*>(ol), make_float4(lo[0], lo[1], lo[2], lo[3]));
        else for (int v = 0; v < nv; ++v) ol[v] = lo[v];
      }
    }
  }
}

// fp32 build written directly as the fp16 operand planes of the split-FP16 tensor-core GEMMs (csrc/umma_gemm.cu):
// out_h + out_l ~= K * scale[l], scale[l] = the power of two that puts sigma_l^2 (+|jitter|), the largest possible entry,
// in (2^14, 2^15].  4 bytes per entry leave the SM (half of K + lo plane in fp32); each thread owns 8 consecutive j and
// streams one 16-byte store per plane, row and factor.
constexpr int KB_VEC8 = 8;
template <bool MG, bool MAT = false>
__global__ void __launch_bounds__(KB_THREADS) kbuild_fwd_h_kernel(const KBArgs<float> a, __half* __restrict__ out_h,
                                                                  __half* __restrict__ out_l, float* __restrict__ out_scale) {
  __shared__ float s_c[KB_LMAX * 8], s_s2[KB_LMAX * 8], s_a[KB_LMAX * 8], s_sc[KB_LMAX * 8];
  __shared__ float s_r2[MG ? KB_GMAX * KB_GMAX : 1];
  const int tid = threadIdx.x;
  for (int l = tid; l < a.L; l += KB_THREADS) {
    const float ls = a.ls[l], sg = a.sigma[l];
    s_c[l] = MAT ? sqrtf(3.f) / ls : -0.5f / (ls * ls);
    s_s2[l] = sg * sg;
    s_sc[l] = gpz_pow2_scale(sg * sg + fabsf(a.jitter));
    if (MG) s_a[l] = a.a[l];
    if (blockIdx.x == 0 && blockIdx.y == 0) out_scale[l] = s_sc[l];
  }
  if (MG) for (int e = tid; e < a.ng * a.ng; e += KB_THREADS) s_r2[e] = a.r2[e];
  __syncthreads();

  const int64_t j0 = ((int64_t)blockIdx.x * KB_THREADS + tid) * KB_VEC8;
  if (j0 >= a.n2) return;                        // n2 % 8 == 0 (checked by the host): whole vectors only
  float xj[KB_DMAX][KB_VEC8];
  int gj[KB_VEC8];
#pragma unroll
  for (int v = 0; v < KB_VEC8; ++v) {
#pragma unroll
    for (int d = 0; d < KB_DMAX; ++d) xj[d][v] = d < a.D ? a.x2[(j0 + v) * a.D + d] : 0.f;
    gj[v] = MG ? (int)a.g2[j0 + v] : 0;
  }
  const int i0 = blockIdx.y * KB_ROWS, i1 = min(i0 + KB_ROWS, a.n1);
  for (int i = i0; i < i1; ++i) {
    float d2[KB_VEC8], r2[KB_VEC8];
#pragma unroll
    for (int v = 0; v < KB_VEC8; ++v) d2[v] = 0.f;
#pragma unroll
    for (int d = 0; d < KB_DMAX; ++d) {
      if (d < a.D) {
        const float xi = a.x1[(int64_t)i * a.D + d];
#pragma unroll
        for (int v = 0; v < KB_VEC8; ++v) { const float df = xi - xj[d][v]; d2[v] = fmaf(df, df, d2[v]); }
      }
    }
    if (MG) {
      const int gi = (int)a.g1[i];
#pragma unroll
      for (int v = 0; v < KB_VEC8; ++v) r2[v] = s_r2[gi * a.ng + gj[v]];
    }
    for (int l = 0; l < a.L; ++l) {
      const float c = s_c[l], s2 = s_s2[l], sc = s_sc[l];
      float k[KB_VEC8];
#pragma unroll
      for (int v = 0; v < KB_VEC8; ++v) {
        float val;
        if (MG) {
          const float den = fmaf(s_a[l], r2[v], 1.f);
          const float idn = __fdividef(1.f, den);
          const float scd = a.p_half == 1.f ? idn : powf(den, -a.p_half);
          val = s2 * expf(c * d2[v] * idn) * scd;
        } else if (MAT) {
          const float vv = c * sqrtf(d2[v]);
          val = s2 * (1.f + vv) * expf(-vv);
        } else {
          val = s2 * expf(c * d2[v]);
        }
        if (a.jitter != 0.f && (int64_t)i == j0 + v) val += a.jitter;
        k[v] = val * sc;
      }
      uint4 h, lo;
      { const __half2 hh = __floats2half2_rn(k[0], k[1]); const float2 f = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(k[0] - f.x, k[1] - f.y);
        h.x = *reinterpret_cast<const uint32_t*>(&hh); lo.x = *reinterpret_cast<const uint32_t*>(&ll); }
      { const __half2 hh = __floats2half2_rn(k[2], k[3]); const float2 f = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(k[2] - f.x, k[3] - f.y);
        h.y = *reinterpret_cast<const uint32_t*>(&hh); lo.y = *reinterpret_cast<const uint32_t*>(&ll); }
      { const __half2 hh = __floats2half2_rn(k[4], k[5]); const float2 f = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(k[4] - f.x, k[5] - f.y);
        h.z = *reinterpret_cast<const uint32_t*>(&hh); lo.z = *reinterpret_cast<const uint32_t*>(&ll); }
      { const __half2 hh = __floats2half2_rn(k[6], k[7]); const float2 f = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(k[6] - f.x, k[7] - f.y);
        h.w = *reinterpret_cast<const uint32_t*>(&hh); lo.w = *reinterpret_cast<const uint32_t*>(&ll); }
      const int64_t o = ((int64_t)l * a.n1 + i) * a.n2 + j0;
      __stcs(reinterpret_cast<uint4*>(out_h + o), h);
      __stcs(reinterpret_cast<uint4*>(out_l + o), lo);
    }
  }
}

// Backward: given G = dLoss/dK (L x n1 x n2) recompute K and reduce
//   g_sigma[l] += (2/sigma_l) sum G K0                       (K0 without jitter)
//   g_ls[l]    += sum G K0 d2 / (ls^3 den)
//   g_a[l]     += sum G K0 (0.5 d2/(ls^2 den^2) - p_half/den) r2             (MG)
//   g_x1[i,:]  -= sum_{l,j} G K0 (x1_i - x2_j)/(ls^2 den) ;  g_x2[j,:] += same
// One CTA covers KB_ROWS rows x (KB_THREADS*4) columns; the L accumulators live in registers.  The recomputed K uses the MUFU
// exponential in fp32 (error ~1e-6 relative on entries that only enter sums with a 1e-4 parity tolerance; fp64 stays exact):
// the kernel is issue-bound, not DRAM-bound, and expf's range reduction was a third of its instructions.
// Matern-3/2 (MAT): with e = exp(-v), K0 = s2 (1 + v) e and u = G s2 (3 / ls^2) e the same three accumulations apply:
//   g_sigma += (2/sigma) sum G K0 ;  g_ls += sum u d2 / ls  (= G s2 v^2 e / ls) ;  dK/dx1 = -u (x1 - x2)  (finite at d = 0)
template <typename T, bool MG, bool ALIGNED, int LMAX, bool MAT = false>
__global__ void __launch_bounds__(KB_THREADS, (sizeof(T) == 4 && LMAX <= 12) ? 2 : 1) kbuild_bwd_kernel(const KBArgs<T> a, const T* __restrict__ G, int l0, int Lc,
                                                                 double* __restrict__ g_x1, T* __restrict__ g_x2,
                                                                 double* __restrict__ g_sigma, double* __restrict__ g_ls,
                                                                 double* __restrict__ g_a) {
  __shared__ T s_c[LMAX], s_s2[LMAX], s_a[LMAX];
  __shared__ T s_r2[MG ? KB_GMAX * KB_GMAX : 1];
  __shared__ double s_red[32];
  __shared__ T s_row[KB_ROWS][KB_DMAX][KB_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int l = tid; l < Lc; l += KB_THREADS) {
    const T ls = a.ls[l0 + l], sg = a.sigma[l0 + l];
    s_c[l] = MAT ? Num<T>::sqrt(T(3)) / ls : T(-0.5) / (ls * ls);
    s_s2[l] = sg * sg;
    if (MG) s_a[l] = a.a[l0 + l];
  }
  if (MG) for (int e = tid; e < a.ng * a.ng; e += KB_THREADS) s_r2[e] = a.r2[e];
  __syncthreads();

  const int64_t j0 = ((int64_t)blockIdx.x * KB_THREADS + tid) * KB_VEC;
  const int nv = j0 < a.n2 ? (int)min((int64_t)KB_VEC, a.n2 - j0) : 0;
  T xj[KB_DMAX][KB_VEC];
  int gj[KB_VEC];
#pragma unroll
  for (int v = 0; v < KB_VEC; ++v) {
#pragma unroll
    for (int d = 0; d < KB_DMAX; ++d) xj[d][v] = (v < nv && d < a.D) ? a.x2[(j0 + v) * a.D + d] : T(0);
    gj[v] = (MG && v < nv) ? (int)a.g2[j0 + v] : 0;
  }
  T acc_s[LMAX], acc_l[LMAX], acc_a[MG ? LMAX : 1];
#pragma unroll
  for (int l = 0; l < LMAX; ++l) { acc_s[l] = T(0); acc_l[l] = T(0); if (MG) acc_a[l] = T(0); }
  T gx2[KB_DMAX][KB_VEC];
#pragma unroll
  for (int d = 0; d < KB_DMAX; ++d)
#pragma unroll
    for (int v = 0; v < KB_VEC; ++v) gx2[d][v] = T(0);

  const int i0 = blockIdx.y * KB_ROWS, i1 = min(i0 + KB_ROWS, a.n1);
  for (int i = i0; i < i1; ++i) {
    T df[KB_DMAX][KB_VEC], d2[KB_VEC], r2[KB_VEC], w[KB_VEC];   // w = sum_l G K0 /(ls^2 den)
#pragma unroll
    for (int v = 0; v < KB_VEC; ++v) { d2[v] = T(0); w[v] = T(0); r2[v] = T(0); }
#pragma unroll
    for (int d = 0; d < KB_DMAX; ++d) {
      const T xi = d < a.D ? a.x1[(int64_t)i * a.D + d] : T(0);
#pragma unroll
      for (int v = 0; v < KB_VEC; ++v) { df[d][v] = xi - xj[d][v]; d2[v] = fma(df[d][v], df[d][v], d2[v]); }
    }
    if (MG) {
      const int gi = (int)a.g1[i];
#pragma unroll
      for (int v = 0; v < KB_VEC; ++v) r2[v] = s_r2[gi * a.ng + gj[v]];
    }
    if (nv > 0) {
      // phase 1: issue every load of this row (one 16-byte vector per factor) before any is used
      T g[LMAX][KB_VEC];
#pragma unroll
      for (int l = 0; l < LMAX; ++l) {
        const T* gp = G + ((int64_t)(l0 + (l < Lc ? l : 0)) * a.n1 + i) * a.n2 + j0;
        if (ALIGNED && nv == KB_VEC) {
          if (sizeof(T) == 4) {
            const float4 t = __ldcs(reinterpret_cast<const float4*>(gp));
            g[l][0] = t.x; g[l][1] = t.y; g[l][2] = t.z; g[l][3] = t.w;
          } else {
            const double2 t0 = __ldcs(reinterpret_cast<const double2*>(gp));
            const double2 t1 = __ldcs(reinterpret_cast<const double2*>(gp) + 1);
            g[l][0] = t0.x; g[l][1] = t0.y; g[l][2] = t1.x; g[l][3] = t1.y;
          }
        } else {
#pragma unroll
          for (int v = 0; v < KB_VEC; ++v) g[l][v] = v < nv ? gp[v] : T(0);
        }
      }
      // phase 2: recompute K and accumulate
#pragma unroll
      for (int l = 0; l < LMAX; ++l) {
        if (l < Lc) {
          const T c = s_c[l], s2 = s_s2[l];
#pragma unroll
          for (int v = 0; v < KB_VEC; ++v) {
            T k0, idn, ee = T(0);
            if (MG) {
              const T den = fma(s_a[l], r2[v], T(1));
              idn = fast_div(T(1), den);
              const T sc = a.p_half == T(1) ? idn : Num<T>::pow(den, -a.p_half);
              k0 = s2 * fast_exp(c * d2[v] * idn) * sc;
            } else if (MAT) {
              idn = T(1);
              const T vv = c * Num<T>::sqrt(d2[v]);
              ee = s2 * fast_exp(-vv);
              k0 = (T(1) + vv) * ee;
            } else {
              idn = T(1);
              k0 = s2 * fast_exp(c * d2[v]);
            }
            const T gk = g[l][v] * k0;
            const T u = MAT ? g[l][v] * ee * c * c       // G s2 exp(-v) 3 / ls^2
                            : gk * (T(-2) * c) * idn;    // G K0 / (ls^2 den)
            acc_s[l] += gk;
            acc_l[l] = fma(u, d2[v], acc_l[l]);          // later divided by ls
            if (MG) acc_a[l] = fma(gk * r2[v], (T(-1) * c * d2[v] * idn - a.p_half) * idn, acc_a[l]);
            w[v] += u;
          }
        }
      }
    }
    // row reduction for g_x1[i,:]
#pragma unroll
    for (int d = 0; d < KB_DMAX; ++d) {
      if (d < a.D) {
        T r = T(0);
#pragma unroll
        for (int v = 0; v < KB_VEC; ++v) { r = fma(w[v], df[d][v], r); gx2[d][v] = fma(w[v], df[d][v], gx2[d][v]); }
        r = warp_sum(r);
        if (lane == 0) s_row[i - i0][d][warp] = r;
      }
    }
  }
  __syncthreads();
  if (g_x1 != nullptr) {
    for (int e = tid; e < (i1 - i0) * a.D; e += KB_THREADS) {
      const int ii = e / a.D, d = e % a.D;
      double r = 0.0;
#pragma unroll
      for (int wv = 0; wv < KB_THREADS / 32; ++wv) r += (double)s_row[ii][d][wv];
      atomicAdd(g_x1 + (int64_t)(i0 + ii) * a.D + d, -r);
    }
  }
  if (g_x2 != nullptr) {
    for (int v = 0; v < nv; ++v)
#pragma unroll
      for (int d = 0; d < KB_DMAX; ++d)
        if (d < a.D) atomicAdd(g_x2 + (j0 + v) * a.D + d, gx2[d][v]);
  }
#pragma unroll
  for (int l = 0; l < LMAX; ++l) {
    if (l < Lc) {                                   // uniform across the block
      double v = block_sum<double>((double)acc_s[l], s_red);
      if (tid == 0) atomicAdd(g_sigma + l0 + l, v);
      v = block_sum<double>((double)acc_l[l], s_red);
      if (tid == 0) atomicAdd(g_ls + l0 + l, v);
      if (MG) {
        v = block_sum<double>((double)acc_a[l], s_red);
        if (tid == 0) atomicAdd(g_a + l0 + l, v);
      }
    }
  }
}

// double accumulators -> outputs:  g_sigma = 2 S/sigma, g_ls = Sl/ls, g_a = Sa, g_x1 = ws_x1
template <typename T>
__global__ void kbuild_bwd_finalize_kernel(const double* __restrict__ ws, const T* __restrict__ sigma, const T* __restrict__ ls, int L,
                                           int n1D, T* __restrict__ g_sigma, T* __restrict__ g_ls, T* __restrict__ g_a,
                                           T* __restrict__ g_x1) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < L) {
    g_sigma[e] = (T)(2.0 * ws[e] / (double)sigma[e]);
    g_ls[e] = (T)(ws[L + e] / (double)ls[e]);
    if (g_a != nullptr) g_a[e] = (T)ws[2 * L + e];
  }
  if (g_x1 != nullptr && e < n1D) g_x1[e] = (T)ws[3 * L + e];
}

// plain Euclidean distance matrix (n1 x n2) by direct differences: kernel(X, Z, return_distance=True)
// (kernels.py:118-124) without cdist's matmul path.
template <typename T>
__global__ void cdist_kernel(const T* __restrict__ x1, const T* __restrict__ x2, T* __restrict__ out, int n1, int n2, int D) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)n1 * n2) return;
  const int i = (int)(e / n2), j = (int)(e % n2);
  T d2 = T(0);
  for (int d = 0; d < D; ++d) { const T df = x1[(int64_t)i * D + d] - x2[(int64_t)j * D + d]; d2 = fma(df, df, d2); }
  out[e] = Num<T>::sqrt(d2);
}

template <typename T>
int kbuild_fwd(const KBArgs<T>& a, T* out, T* out_lo, cudaStream_t st) {
  if (a.D < 1 || a.D > KB_DMAX || a.L < 1 || a.L > KB_LMAX * 8) return GPZ_ERR_UNSUPPORTED;
  const bool mg = a.g1 != nullptr;
  if (mg && (a.ng < 1 || a.ng > KB_GMAX)) return GPZ_ERR_UNSUPPORTED;
  if (a.n1 == 0 || a.n2 == 0) return GPZ_OK;
  const bool al = (a.n2 % KB_VEC == 0) && ((reinterpret_cast<uintptr_t>(out) & 31) == 0) &&
                  ((reinterpret_cast<uintptr_t>(out_lo) & 15) == 0);
  dim3 grid((unsigned)cdiv(a.n2, (int64_t)KB_THREADS * KB_VEC), (unsigned)cdiv(a.n1, KB_ROWS));
  if (a.kind != 0 && (a.kind != 1 || mg)) return GPZ_ERR_UNSUPPORTED;       // Matern-3/2 has no multi-group form in the reference
  if (a.kind == 1) {
    if (al) kbuild_fwd_kernel<T, false, true, true><<<grid, KB_THREADS, 0, st>>>(a, out, out_lo);
    else kbuild_fwd_kernel<T, false, false, true><<<grid, KB_THREADS, 0, st>>>(a, out, out_lo);
  } else if (mg) {
    if (al) kbuild_fwd_kernel<T, true, true><<<grid, KB_THREADS, 0, st>>>(a, out, out_lo);
    else kbuild_fwd_kernel<T, true, false><<<grid, KB_THREADS, 0, st>>>(a, out, out_lo);
  } else {
    if (al) kbuild_fwd_kernel<T, false, true><<<grid, KB_THREADS, 0, st>>>(a, out, out_lo);
    else kbuild_fwd_kernel<T, false, false><<<grid, KB_THREADS, 0, st>>>(a, out, out_lo);
  }
  GPZ_CHECK_LAUNCH();
  return GPZ_OK;
}

template <typename T>
int kbuild_bwd(const KBArgs<T>& a, const T* G, T* g_x1, T* g_x2, T* g_sigma, T* g_ls, T* g_a, double* ws, cudaStream_t st) {
  if (a.D < 1 || a.D > KB_DMAX || a.L < 1) return GPZ_ERR_UNSUPPORTED;
  const bool mg = a.g1 != nullptr;
  if (mg && (a.ng < 1 || a.ng > KB_GMAX)) return GPZ_ERR_UNSUPPORTED;
  const int n1D = a.n1 * a.D;
  GPZ_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * (3 * (size_t)a.L + n1D), st));
  if (g_x2) GPZ_CUDA(cudaMemsetAsync(g_x2, 0, sizeof(T) * a.n2 * a.D, st));
  double* w_sigma = ws; double* w_ls = ws + a.L; double* w_a = ws + 2 * a.L; double* w_x1 = ws + 3 * a.L;
  if (a.n1 > 0 && a.n2 > 0) {
    const bool al = (a.n2 % KB_VEC == 0) && ((reinterpret_cast<uintptr_t>(G) & 31) == 0);
    dim3 grid((unsigned)cdiv(a.n2, (int64_t)KB_THREADS * KB_VEC), (unsigned)cdiv(a.n1, KB_ROWS));
    if (a.kind != 0 && (a.kind != 1 || mg)) return GPZ_ERR_UNSUPPORTED;
    const int lmax = a.kind == 1 ? KB_LMAX : (a.L <= 4 ? 4 : (a.L <= 8 ? 8 : (a.L <= 12 ? 12 : (a.L <= 16 ? 16 : KB_LMAX))));
    for (int l0 = 0; l0 < a.L; l0 += lmax) {
      const int Lc = min(lmax, a.L - l0);
      // g_x1/g_x2 accumulate over all l-chunks (atomics), so every chunk launch adds its share
#define GPZ_KB_LAUNCH(MGV, ALV, LM) \
  kbuild_bwd_kernel<T, MGV, ALV, LM><<<grid, KB_THREADS, 0, st>>>(a, G, l0, Lc, w_x1, g_x2, w_sigma, w_ls, w_a)
#define GPZ_KB_DISPATCH(LM)                                   \
  do {                                                        \
    if (mg) { if (al) GPZ_KB_LAUNCH(true, true, LM); else GPZ_KB_LAUNCH(true, false, LM); }     \
    else { if (al) GPZ_KB_LAUNCH(false, true, LM); else GPZ_KB_LAUNCH(false, false, LM); }      \
  } while (0)
      if (a.kind == 1) {
        if (al) kbuild_bwd_kernel<T, false, true, KB_LMAX, true><<<grid, KB_THREADS, 0, st>>>(a, G, l0, Lc, w_x1, g_x2, w_sigma, w_ls, w_a);
        else kbuild_bwd_kernel<T, false, false, KB_LMAX, true><<<grid, KB_THREADS, 0, st>>>(a, G, l0, Lc, w_x1, g_x2, w_sigma, w_ls, w_a);
      } else if (lmax == 4) GPZ_KB_DISPATCH(4);
      else if (lmax == 8) GPZ_KB_DISPATCH(8);
      else if (lmax == 12) GPZ_KB_DISPATCH(12);
      else if (lmax == 16) GPZ_KB_DISPATCH(16);
      else GPZ_KB_DISPATCH(KB_LMAX);
#undef GPZ_KB_DISPATCH
#undef GPZ_KB_LAUNCH
      GPZ_CHECK_LAUNCH();
    }
  }
  const int tot = max(a.L, n1D);
  kbuild_bwd_finalize_kernel<T><<<(unsigned)cdiv(tot, 256), 256, 0, st>>>(ws, a.sigma, a.ls, a.L, n1D, g_sigma, g_ls, g_a, g_x1);
  GPZ_CHECK_LAUNCH();
  return GPZ_OK;
}

}  // namespace gpz

using namespace gpz;

#define GPZ_KB_IMPL(SUF, T)                                                                                        \
  extern "C" int gpz_kernel_build_fwd_##SUF(const T* x1, const T* x2, const T* sigma, const T* ls, const T* a,     \
                                            const T* r2, const int64_t* g1, const int64_t* g2, int n1, int n2,     \
                                            int D, int L, int ng, int kind, T p_half, T jitter, T* out, T* out_lo, \
                                            void* stream) {                                                        \
    KBArgs<T> k{x1, x2, sigma, ls, a, r2, g1, g2, n1, n2, D, L, ng, p_half, jitter, kind};                         \
    return kbuild_fwd<T>(k, out, out_lo, (cudaStream_t)stream);                                                    \
  }                                                                                                                \
  extern "C" int gpz_kernel_build_bwd_##SUF(const T* x1, const T* x2, const T* sigma, const T* ls, const T* a,     \
                                            const T* r2, const int64_t* g1, const int64_t* g2, int n1, int n2,     \
                                            int D, int L, int ng, int kind, T p_half, const T* G, T* g_x1, T* g_x2, \
                                            T* g_sigma, T* g_ls, T* g_a, double* ws, void* stream) {               \
    KBArgs<T> k{x1, x2, sigma, ls, a, r2, g1, g2, n1, n2, D, L, ng, p_half, T(0), kind};                           \
    return kbuild_bwd<T>(k, G, g_x1, g_x2, g_sigma, g_ls, g_a, ws, (cudaStream_t)stream);                          \
  }                                                                                                                \
  extern "C" int gpz_cdist_##SUF(const T* x1, const T* x2, T* out, int n1, int n2, int D, void* stream) {          \
    const int64_t total = (int64_t)n1 * n2;                                                                        \
    if (total == 0) return GPZ_OK;                                                                                 \
    cdist_kernel<T><<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x1, x2, out, n1, n2, D);         \
    GPZ_CHECK_LAUNCH();                                                                                            \
    return GPZ_OK;                                                                                                 \
  }

GPZ_KB_IMPL(f32, float)
GPZ_KB_IMPL(f64, double)

// fp16-plane forward (fp32 arithmetic): see kbuild_fwd_h_kernel
extern "C" int gpz_kernel_build_fwd_h_f32(const float* x1, const float* x2, const float* sigma, const float* ls, const float* a,
                                          const float* r2, const int64_t* g1, const int64_t* g2, int n1, int n2, int D, int L,
                                          int ng, int kind, float p_half, float jitter, void* out_h, void* out_l,
                                          float* out_scale, void* stream) {
  KBArgs<float> k{x1, x2, sigma, ls, a, r2, g1, g2, n1, n2, D, L, ng, p_half, jitter, kind};
  if (D < 1 || D > KB_DMAX || L < 1 || L > KB_LMAX * 8) return GPZ_ERR_UNSUPPORTED;
  const bool mg = g1 != nullptr;
  if (mg && (ng < 1 || ng > KB_GMAX)) return GPZ_ERR_UNSUPPORTED;
  if (n2 % KB_VEC8 || (reinterpret_cast<uintptr_t>(out_h) & 15) || (reinterpret_cast<uintptr_t>(out_l) & 15)) return GPZ_ERR_UNSUPPORTED;
  if (n1 == 0 || n2 == 0) return GPZ_OK;
  dim3 grid((unsigned)cdiv(n2, (int64_t)KB_THREADS * KB_VEC8), (unsigned)cdiv(n1, KB_ROWS));
  if (kind != 0 && (kind != 1 || mg)) return GPZ_ERR_UNSUPPORTED;
  if (kind == 1) kbuild_fwd_h_kernel<false, true><<<grid, KB_THREADS, 0, (cudaStream_t)stream>>>(k, (__half*)out_h, (__half*)out_l, out_scale);
  else if (mg) kbuild_fwd_h_kernel<true><<<grid, KB_THREADS, 0, (cudaStream_t)stream>>>(k, (__half*)out_h, (__half*)out_l, out_scale);
  else kbuild_fwd_h_kernel<false><<<grid, KB_THREADS, 0, (cudaStream_t)stream>>>(k, (__half*)out_h, (__half*)out_l, out_scale);
  GPZ_CHECK_LAUNCH();
  return GPZ_OK;
}
