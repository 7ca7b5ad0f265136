// K6: nearest-neighbour variational GP (VNNGP.forward, gp.py:19-122).
//
// Reference: argsort of the full N x M cdist matrix (gp.py:64), an L x N x K x M gather of Cholesky rows (gp.py:67,
// 1.3 GB at config 3, the CUDA OOM of nnnsf_visium_anim_experiment.ipynb:468), little_Kzz = rows rows^T, a second jitter
// and torch.inverse (gp.py:72-77), gathers of Kxz / mu / Lu rows (gp.py:83-102) and svgp_forward on (L*N) K x K
// problems (gp.py:106).  Algebraically per point n and factor l, with nn = the K nearest inducing
// points of x_n (shared by all factors):
//     kzz = Kzz_j[l][nn,nn] + jitter I      (Kzz_j already carries the first jitter: "double jitter")
//     kxz = sigma_l^2 exp(-0.5 |x_n - z_nn|^2 / ls_l^2)
//     w = kzz^-1 kxz ;  mean = w . mu[l][nn] ;  var = Kxx + w^T S[l][nn,nn] w - w . kxz
// K6a: one thread per point scans the inducing points (staged through shared memory) and keeps the K smallest
//      distances in registers, ascending, ties to the lower index (what argsort gives on exact ties).
// K6b: one warp per point, one lane per factor: the K x K Cholesky factorisation and the two triangular solves live
//      entirely in registers; nothing of size L*N*K*M is ever formed.  The backward scatters the K^2 / K entries of
//      dL/dKzz, dL/dS, dL/dmu with atomics and differentiates kxz in place (App. A.7).
#include "common.cuh"
#include "gpzoo_b200.h"

namespace gpz {

constexpr int NN_DMAX = 4;
constexpr int NN_TILE = 256;

template <typename T, int KMAX>
__global__ void __launch_bounds__(128) knn_kernel(const T* __restrict__ X, const T* __restrict__ Z, int64_t* __restrict__ idx,
                                                  int N, int M, int D, int K) {
  __shared__ T zs[NN_TILE][NN_DMAX];
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  T x[NN_DMAX];
#pragma unroll
  for (int d = 0; d < NN_DMAX; ++d) x[d] = (n < N && d < D) ? X[(int64_t)n * D + d] : T(0);
  T bd[KMAX];
  int bi[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) { bd[k] = T(INFINITY); bi[k] = 0x7fffffff; }
  for (int m0 = 0; m0 < M; m0 += NN_TILE) {
    __syncthreads();
    for (int e = threadIdx.x; e < NN_TILE * NN_DMAX; e += blockDim.x) {
      const int r = e / NN_DMAX, d = e % NN_DMAX;
      zs[r][d] = (m0 + r < M && d < D) ? Z[(int64_t)(m0 + r) * D + d] : T(0);
    }
    __syncthreads();
    const int mt = min(NN_TILE, M - m0);
    for (int r = 0; r < mt; ++r) {
      T d2 = T(0);
#pragma unroll
      for (int d = 0; d < NN_DMAX; ++d) { const T df = x[d] - zs[r][d]; d2 = fma(df, df, d2); }   // zs[r][d>=D] = x[d] = 0
      const T dist = Num<T>::sqrt(d2);          // the reference sorts cdist (sqrt) values
      // insert (dist, m0+r) keeping ascending order; strict '<' keeps the earlier (lower) index on ties,
      // after the insertion point everything shifts down by one
      T cd = dist;
      int ci = m0 + r;
      bool ins = false;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k < K && (ins || cd < bd[k])) {
          const T td = bd[k]; const int ti = bi[k];
          bd[k] = cd; bi[k] = ci;
          cd = td; ci = ti;
          ins = true;
        }
      }
    }
  }
  if (n < N)
    for (int k = 0; k < K; ++k) idx[(int64_t)n * K + k] = bi[k];
}

template <typename T> struct VnnArgs {
  const T* X; const T* Z; const T* sigma; const T* ls;
  const T* Kzz; const T* S; const T* mu; const T* kxx;
  const int64_t* nn;
  int N, M, D, L, K;
  T jitter;
};

// Cholesky of the K x K system (lower, in place) and solves; everything stays in registers (loops fully unrolled).
template <typename T, int KMAX> __device__ __forceinline__ void chol_lower(T (&a)[KMAX][KMAX], int K) {
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    if (j < K) {
      T d = a[j][j];
#pragma unroll
      for (int t = 0; t < j; ++t) d -= a[j][t] * a[j][t];
      d = Num<T>::sqrt(d);
      a[j][j] = d;
      const T inv = T(1) / d;
#pragma unroll
      for (int i = j + 1; i < KMAX; ++i) {
        if (i < K) {
          T v = a[i][j];
#pragma unroll
          for (int t = 0; t < j; ++t) v -= a[i][t] * a[j][t];
          a[i][j] = v * inv;
        }
      }
    }
  }
}
template <typename T, int KMAX> __device__ __forceinline__ void chol_solve(const T (&c)[KMAX][KMAX], T (&b)[KMAX], int K) {
#pragma unroll
  for (int i = 0; i < KMAX; ++i) {          // C y = b
    if (i < K) {
      T v = b[i];
#pragma unroll
      for (int t = 0; t < i; ++t) v -= c[i][t] * b[t];
      b[i] = v / c[i][i];
    }
  }
#pragma unroll
  for (int i = KMAX - 1; i >= 0; --i) {     // C^T w = y
    if (i < K) {
      T v = b[i];
#pragma unroll
      for (int t = i + 1; t < KMAX; ++t) if (t < K) v -= c[t][i] * b[t];
      b[i] = v / c[i][i];
    }
  }
}

// BWD = false: mean, var.   BWD = true: gradients (scatter-add), given gm, gv.
template <typename T, int KMAX, bool BWD>
__global__ void __launch_bounds__(128) vnngp_kernel(const VnnArgs<T> a, T* __restrict__ mean, T* __restrict__ var,
                                                    const T* __restrict__ gm, const T* __restrict__ gv, T* __restrict__ gKzz,
                                                    T* __restrict__ gS, T* __restrict__ gmu, T* __restrict__ gZ,
                                                    double* __restrict__ gsl /* [2L]: sigma, ls */) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n = warp;
  if (n >= a.N) return;
  const int K = a.K;
  int nn[KMAX];
  T d2[KMAX], dz[KMAX][NN_DMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    nn[k] = k < K ? (int)a.nn[(int64_t)n * K + k] : 0;
    d2[k] = T(0);
#pragma unroll
    for (int d = 0; d < NN_DMAX; ++d) {
      const T df = (k < K && d < a.D) ? a.Z[(int64_t)nn[k] * a.D + d] - a.X[(int64_t)n * a.D + d] : T(0);
      dz[k][d] = df;                       // z - x
      d2[k] = fma(df, df, d2[k]);
    }
  }
  for (int l = lane; l < a.L; l += 32) {
    const int64_t sMM = (int64_t)a.M * a.M;
    const T sg = a.sigma[l], ls = a.ls[l];
    const T c = T(-0.5) / (ls * ls), s2 = sg * sg;
    T C[KMAX][KMAX], kxz[KMAX], w[KMAX];
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
      kxz[i] = i < K ? s2 * Num<T>::exp(c * d2[i]) : T(0);
      w[i] = kxz[i];
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        C[i][j] = T(0);
        if (j <= i && i < K) C[i][j] = a.Kzz[l * sMM + (int64_t)nn[i] * a.M + nn[j]] + (i == j ? a.jitter : T(0));
      }
    }
    chol_lower<T, KMAX>(C, K);
    chol_solve<T, KMAX>(C, w, K);
    // Sw = S[nn,nn] w
    T Sw[KMAX];
    T mval = T(0), wSw = T(0), wk = T(0);
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
      Sw[i] = T(0);
      if (i < K) {
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
          if (j < K) Sw[i] = fma(a.S[l * sMM + (int64_t)nn[i] * a.M + nn[j]], w[j], Sw[i]);
        wSw = fma(w[i], Sw[i], wSw);
        wk = fma(w[i], kxz[i], wk);
        mval = fma(w[i], a.mu[(int64_t)l * a.M + nn[i]], mval);
      }
    }
    if (!BWD) {
      mean[(int64_t)l * a.N + n] = mval;
      var[(int64_t)l * a.N + n] = a.kxx[(int64_t)l * a.N + n] + wSw - wk;
    } else {
      const T gmn = gm[(int64_t)l * a.N + n], gvn = gv[(int64_t)l * a.N + n];
      T u[KMAX];
#pragma unroll
      for (int i = 0; i < KMAX; ++i)
        u[i] = i < K ? gmn * a.mu[(int64_t)l * a.M + nn[i]] + T(2) * gvn * (Sw[i] - kxz[i]) : T(0);     // dL/dw
      chol_solve<T, KMAX>(C, u, K);                                                                     // u = kzz^-1 dL/dw = dL/dkxz
      double acc_s = 0.0, acc_l = 0.0;
#pragma unroll
      for (int i = 0; i < KMAX; ++i) {
        if (i < K) {
          atomicAdd(gmu + (int64_t)l * a.M + nn[i], gmn * w[i]);
#pragma unroll
          for (int j = 0; j < KMAX; ++j) {
            if (j < K) {
              const int64_t o = l * sMM + (int64_t)nn[i] * a.M + nn[j];
              atomicAdd(gS + o, gvn * w[i] * w[j]);
              atomicAdd(gKzz + o, -gvn * w[i] * w[j] - u[i] * w[j]);
            }
          }
          // kxz_i = s2 exp(c d2_i):  d/dz = kxz * 2c (z - x);  d/dls = kxz d2 / ls^3;  d/dsigma = 2 kxz / sigma
          const T gk = u[i] * kxz[i];
          acc_s += (double)gk;
          acc_l += (double)(gk * d2[i]);
#pragma unroll
          for (int d = 0; d < NN_DMAX; ++d)
            if (d < a.D) atomicAdd(gZ + (int64_t)nn[i] * a.D + d, gk * T(2) * c * dz[i][d]);
        }
      }
      atomicAdd(gsl + l, 2.0 * acc_s / (double)sg);
      atomicAdd(gsl + a.L + l, acc_l / ((double)ls * (double)ls * (double)ls));
    }
  }
}

}  // namespace gpz

using namespace gpz;

#define GPZ_VNN_DISPATCH(K, CALL4, CALL8, CALL16) \
  do { if ((K) <= 4) { CALL4; } else if ((K) <= 8) { CALL8; } else { CALL16; } } while (0)

#define GPZ_VNN_IMPL(SUF, T)                                                                                              \
  extern "C" int gpz_vnngp_neighbors_##SUF(const T* X, const T* Z, int64_t* idx, int N, int M, int D, int K, void* stream) { \
    if (K < 1 || K > 16 || K > M || D < 1 || D > NN_DMAX) return GPZ_ERR_UNSUPPORTED;                                     \
    if (N == 0) return GPZ_OK;                                                                                            \
    cudaStream_t st = (cudaStream_t)stream;                                                                               \
    dim3 grid((unsigned)cdiv(N, 128));                                                                                    \
    GPZ_VNN_DISPATCH(K, (knn_kernel<T, 4><<<grid, 128, 0, st>>>(X, Z, idx, N, M, D, K)),                                  \
                     (knn_kernel<T, 8><<<grid, 128, 0, st>>>(X, Z, idx, N, M, D, K)),                                     \
                     (knn_kernel<T, 16><<<grid, 128, 0, st>>>(X, Z, idx, N, M, D, K)));                                   \
    GPZ_CHECK_LAUNCH();                                                                                                   \
    return GPZ_OK;                                                                                                        \
  }                                                                                                                       \
  extern "C" int gpz_vnngp_fwd_##SUF(const T* X, const T* Z, const T* sigma, const T* ls, const T* Kzz, const T* S,       \
                                     const T* mu, const T* kxx, const int64_t* nn, int N, int M, int D, int L, int K,     \
                                     T jitter, T* mean, T* var, void* stream) {                                           \
    if (K < 1 || K > 16 || D < 1 || D > NN_DMAX) return GPZ_ERR_UNSUPPORTED;                                              \
    if (N == 0) return GPZ_OK;                                                                                            \
    VnnArgs<T> a{X, Z, sigma, ls, Kzz, S, mu, kxx, nn, N, M, D, L, K, jitter};                                            \
    cudaStream_t st = (cudaStream_t)stream;                                                                               \
    dim3 grid((unsigned)cdiv((int64_t)N * 32, 128));                                                                      \
    GPZ_VNN_DISPATCH(K, (vnngp_kernel<T, 4, false><<<grid, 128, 0, st>>>(a, mean, var, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr)), \
                     (vnngp_kernel<T, 8, false><<<grid, 128, 0, st>>>(a, mean, var, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr)),    \
                     (vnngp_kernel<T, 16, false><<<grid, 128, 0, st>>>(a, mean, var, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr)));  \
    GPZ_CHECK_LAUNCH();                                                                                                   \
    return GPZ_OK;                                                                                                        \
  }                                                                                                                       \
  extern "C" int gpz_vnngp_bwd_##SUF(const T* X, const T* Z, const T* sigma, const T* ls, const T* Kzz, const T* S,       \
                                     const T* mu, const T* kxx, const int64_t* nn, int N, int M, int D, int L, int K,     \
                                     T jitter, const T* gm, const T* gv, T* gKzz, T* gS, T* gmu, T* gZ, double* gsl,      \
                                     void* stream) {                                                                      \
    if (K < 1 || K > 16 || D < 1 || D > NN_DMAX) return GPZ_ERR_UNSUPPORTED;                                              \
    cudaStream_t st = (cudaStream_t)stream;                                                                               \
    GPZ_CUDA(cudaMemsetAsync(gKzz, 0, sizeof(T) * (size_t)L * M * M, st));                                                \
    GPZ_CUDA(cudaMemsetAsync(gS, 0, sizeof(T) * (size_t)L * M * M, st));                                                  \
    GPZ_CUDA(cudaMemsetAsync(gmu, 0, sizeof(T) * (size_t)L * M, st));                                                     \
    GPZ_CUDA(cudaMemsetAsync(gZ, 0, sizeof(T) * (size_t)M * D, st));                                                      \
    GPZ_CUDA(cudaMemsetAsync(gsl, 0, sizeof(double) * 2 * (size_t)L, st));                                                \
    if (N == 0) return GPZ_OK;                                                                                            \
    VnnArgs<T> a{X, Z, sigma, ls, Kzz, S, mu, kxx, nn, N, M, D, L, K, jitter};                                            \
    dim3 grid((unsigned)cdiv((int64_t)N * 32, 128));                                                                      \
    GPZ_VNN_DISPATCH(K, (vnngp_kernel<T, 4, true><<<grid, 128, 0, st>>>(a, nullptr, nullptr, gm, gv, gKzz, gS, gmu, gZ, gsl)),  \
                     (vnngp_kernel<T, 8, true><<<grid, 128, 0, st>>>(a, nullptr, nullptr, gm, gv, gKzz, gS, gmu, gZ, gsl)),     \
                     (vnngp_kernel<T, 16, true><<<grid, 128, 0, st>>>(a, nullptr, nullptr, gm, gv, gKzz, gS, gmu, gZ, gsl)));   \
    GPZ_CHECK_LAUNCH();                                                                                                   \
    return GPZ_OK;                                                                                                        \
  }

GPZ_VNN_IMPL(f32, float)
GPZ_VNN_IMPL(f64, double)
