// K3/K4: whitened triangular products and the fused SVGP predictive mean / variance.
//
// Reference (gp.py:218-225, utilities.py:382-397):  W^T = Kzz^-1 Kzx via cholesky_solve (two TRSMs),
// mean = W mu, cov = Kxx + sum((W (S - Kzz)) o W).  With Lc = chol(Kzz), Linv = Lc^-1, T = Linv Lu,
// q = Linv mu the same quantities are (exact algebra):
//     A = Linv Kzx          (lower-triangular product  == TRSM  Lc A = Kzx)
//     C = T^T A             (upper-triangular product  == Lu^T Lc^-T A)
//     mean[n] = sum_m q[m] A[m,n]
//     var[n]  = Kxx[n] - sum_m A[m,n]^2 + sum_m C[m,n]^2
// which needs 2 M^2 N flops instead of 6 M^2 N and never forms S - Kzz (no cancellation).
//
// Backward (App. A.5), given gm = dL/dmean, gv = dL/dvar (L x N):
//     gC = 2 C diag(gv)                    gT   = tril(A gC^T)           gq = A gm
//     gA = T gC - 2 A diag(gv) + q gm^T    gKzx = Linv^T gA              gLinv = tril(gA Kzx^T)
//     gKxx = gv
#include "gemm_simt.cuh"
#include "gpzoo_b200.h"

namespace gpz {

// mean / var column reductions over m.  One thread per column n, rows streamed coalesced.
template <typename T>
__global__ void __launch_bounds__(256) predict_reduce_kernel(const T* __restrict__ A, const T* __restrict__ C,
                                                              const T* __restrict__ q, const T* __restrict__ kxx,
                                                              T* __restrict__ mean, T* __restrict__ var, int M, int N) {
  extern __shared__ unsigned char smem_raw[];
  T* sq = reinterpret_cast<T*>(smem_raw);
  const int l = blockIdx.y;
  for (int m = threadIdx.x; m < M; m += blockDim.x) sq[m] = q[(int64_t)l * M + m];
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const T* a = A + (int64_t)l * M * N + n;
  const T* c = C + (int64_t)l * M * N + n;
  T mu = T(0), s_a = T(0), s_c = T(0);
#pragma unroll 4
  for (int m = 0; m < M; ++m) {
    const T av = a[(int64_t)m * N], cv = c[(int64_t)m * N];
    mu = fma(sq[m], av, mu);
    s_a = fma(av, av, s_a);
    s_c = fma(cv, cv, s_c);
  }
  mean[(int64_t)l * N + n] = mu;
  var[(int64_t)l * N + n] = kxx[(int64_t)l * N + n] - s_a + s_c;
}

// backward prologue, one pass over A and C:  C <- gC = 2 C gv ;  gA0 = -2 A gv + q gm^T
template <typename T>
__global__ void __launch_bounds__(256) predict_bwd_prep_kernel(const T* __restrict__ A, T* __restrict__ C, const T* __restrict__ q,
                                                                const T* __restrict__ gm, const T* __restrict__ gv,
                                                                T* __restrict__ gA, int M, int N, int rows_per_cta) {
  const int l = blockIdx.z;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const T gmn = gm[(int64_t)l * N + n], gvn = gv[(int64_t)l * N + n];
  const int m0 = blockIdx.y * rows_per_cta, m1 = min(m0 + rows_per_cta, M);
  for (int m = m0; m < m1; ++m) {
    const int64_t e = ((int64_t)l * M + m) * N + n;
    const T av = A[e], cv = C[e];
    C[e] = T(2) * cv * gvn;
    gA[e] = fma(q[(int64_t)l * M + m], gmn, T(-2) * av * gvn);
  }
}

// out[l,m] = sum_n A[l,m,n] v[l,n]      (gq = A gm)
template <typename T>
__global__ void __launch_bounds__(256) rowdot_kernel(const T* __restrict__ A, const T* __restrict__ v, T* __restrict__ out, int M, int N) {
  __shared__ T red[32];
  const int m = blockIdx.x, l = blockIdx.y;
  const T* a = A + ((int64_t)l * M + m) * N;
  const T* vv = v + (int64_t)l * N;
  T acc = T(0);
  for (int n = threadIdx.x; n < N; n += blockDim.x) acc = fma(a[n], vv[n], acc);
  acc = block_sum<T>(acc, red);
  if (threadIdx.x == 0) out[(int64_t)l * M + m] = acc;
}

template <typename T>
int predict_fwd(const T* Kzx, const T* Linv, const T* Tm, const T* q, const T* kxx, T* A, T* C, T* mean, T* var,
                int M, int N, int L, cudaStream_t st) {
  const int64_t sMM = (int64_t)M * M, sMN = (int64_t)M * N;
  int rc = gemm<T>(st, false, false, M, N, M, T(1), Linv, M, sMM, Kzx, N, sMN, T(0), A, N, sMN, L, 1, 0, 0);
  if (rc) return rc;
  rc = gemm<T>(st, true, false, M, N, M, T(1), Tm, M, sMM, A, N, sMN, T(0), C, N, sMN, L, 2, 0, 0);
  if (rc) return rc;
  predict_reduce_kernel<T><<<dim3((unsigned)cdiv(N, 256), L), 256, sizeof(T) * M, st>>>(A, C, q, kxx, mean, var, M, N);
  GPZ_CHECK_LAUNCH();
  return GPZ_OK;
}

// gT and gLinv must be zero-initialised by the caller (only their lower triangles are written, split-K atomics).
template <typename T>
int predict_bwd(const T* Kzx, const T* Linv, const T* Tm, const T* q, const T* A, T* C, const T* gm, const T* gv,
                T* gA, T* gKzx, T* gLinv, T* gT, T* gq, int M, int N, int L, cudaStream_t st) {
  const int64_t sMM = (int64_t)M * M, sMN = (int64_t)M * N;
  const int rows = 64;
  predict_bwd_prep_kernel<T><<<dim3((unsigned)cdiv(N, 256), (unsigned)cdiv(M, rows), L), 256, 0, st>>>(A, C, q, gm, gv, gA, M, N, rows);
  GPZ_CHECK_LAUNCH();
  rowdot_kernel<T><<<dim3(M, L), 256, 0, st>>>(A, gm, gq, M, N);
  GPZ_CHECK_LAUNCH();
  // split the long reduction over the N spots so the small M x M outputs still fill the GPU
  int splitk = 1;
  {
    const int64_t tiles = cdiv(M, GemmCfg<T>::BM) * cdiv(M, GemmCfg<T>::BN) * L / 2 + 1;
    while (splitk < 16 && tiles * splitk < 148 * 4 && N / (splitk * 2) >= 512) splitk *= 2;
  }
  // gT = tril(A gC^T)
  int rc = gemm<T>(st, false, true, M, M, N, T(1), A, N, sMN, C, N, sMN, T(0), gT, M, sMM, L, 0, 0, 1, splitk > 1 ? splitk : 1);
  if (rc) return rc;
  // gA += T gC
  rc = gemm<T>(st, false, false, M, N, M, T(1), Tm, M, sMM, C, N, sMN, T(1), gA, N, sMN, L, 1, 0, 0);
  if (rc) return rc;
  // gKzx = Linv^T gA
  rc = gemm<T>(st, true, false, M, N, M, T(1), Linv, M, sMM, gA, N, sMN, T(0), gKzx, N, sMN, L, 2, 0, 0);
  if (rc) return rc;
  // gLinv = tril(gA Kzx^T)
  rc = gemm<T>(st, false, true, M, M, N, T(1), gA, N, sMN, Kzx, N, sMN, T(0), gLinv, M, sMM, L, 0, 0, 1, splitk > 1 ? splitk : 1);
  return rc;
}

}  // namespace gpz

using namespace gpz;

#define GPZ_PREDICT_IMPL(SUF, T)                                                                                      \
  extern "C" int gpz_svgp_predict_fwd_##SUF(const T* Kzx, const T* Linv, const T* Tm, const T* q, const T* kxx, T* A, \
                                            T* C, T* mean, T* var, int M, int N, int L, void* stream) {               \
    if (M <= 0 || L <= 0 || N < 0) return GPZ_ERR_BADARG;                                                             \
    if (N == 0) return GPZ_OK;                                                                                        \
    return predict_fwd<T>(Kzx, Linv, Tm, q, kxx, A, C, mean, var, M, N, L, (cudaStream_t)stream);                     \
  }                                                                                                                   \
  extern "C" int gpz_svgp_predict_bwd_##SUF(const T* Kzx, const T* Linv, const T* Tm, const T* q, const T* A, T* C,   \
                                            const T* gm, const T* gv, T* gA, T* gKzx, T* gLinv, T* gT, T* gq, int M,  \
                                            int N, int L, void* stream) {                                             \
    if (M <= 0 || L <= 0 || N < 0) return GPZ_ERR_BADARG;                                                             \
    if (N == 0) return GPZ_OK;                                                                                        \
    return predict_bwd<T>(Kzx, Linv, Tm, q, A, C, gm, gv, gA, gKzx, gLinv, gT, gq, M, N, L, (cudaStream_t)stream);    \
  }

GPZ_PREDICT_IMPL(f32, float)
GPZ_PREDICT_IMPL(f64, double)
