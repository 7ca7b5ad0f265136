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
#include "umma_gemm.h"

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
                                                                T* __restrict__ gA, T* __restrict__ Clo, int M, int N,
                                                                int rows_per_cta) {
  const int l = blockIdx.z;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const T gmn = gm[(int64_t)l * N + n], gvn = gv[(int64_t)l * N + n];
  const int m0 = blockIdx.y * rows_per_cta, m1 = min(m0 + rows_per_cta, M);
  for (int m = m0; m < m1; ++m) {
    const int64_t e = ((int64_t)l * M + m) * N + n;
    const T av = A[e], cv = C[e];
    const T gc = T(2) * cv * gvn;
    C[e] = gc;
    if (sizeof(T) == 4 && Clo != nullptr) {
      const float x = (float)gc;
      reinterpret_cast<float*>(Clo)[e] = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    }
    gA[e] = fma(q[(int64_t)l * M + m], gmn, T(-2) * av * gvn);
  }
}

// out[l,m] = sum_n A[l,m,n] v[l,n]      (gq = A gm); one CTA per row, 4 independent accumulators per thread
template <typename T>
__global__ void __launch_bounds__(256) rowdot_kernel(const T* __restrict__ A, const T* __restrict__ v, T* __restrict__ out, int M, int N) {
  __shared__ T red[32];
  const int m = blockIdx.x, l = blockIdx.y;
  const T* a = A + ((int64_t)l * M + m) * N;
  const T* vv = v + (int64_t)l * N;
  T acc0 = T(0), acc1 = T(0), acc2 = T(0), acc3 = T(0);
  int n = threadIdx.x;
  for (; n + 3 * 256 < N; n += 4 * 256) {
    const T a0 = a[n], a1 = a[n + 256], a2 = a[n + 512], a3 = a[n + 768];
    acc0 = fma(a0, vv[n], acc0);
    acc1 = fma(a1, vv[n + 256], acc1);
    acc2 = fma(a2, vv[n + 512], acc2);
    acc3 = fma(a3, vv[n + 768], acc3);
  }
  for (; n < N; n += 256) acc0 = fma(a[n], vv[n], acc0);
  T acc = block_sum<T>((acc0 + acc1) + (acc2 + acc3), red);
  if (threadIdx.x == 0) out[(int64_t)l * M + m] = acc;
}

// out[l,j] = sum_k A[l,k,j] v[l,k]      (A^T v for row-major A: one thread per column, rows streamed coalesced)
template <typename T>
__global__ void __launch_bounds__(256) coldot_kernel(const T* __restrict__ A, const T* __restrict__ v, T* __restrict__ out, int K, int N) {
  const int l = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const T* a = A + (int64_t)l * K * N + j;
  const T* vv = v + (int64_t)l * K;
  T acc0 = T(0), acc1 = T(0);
  int k = 0;
  for (; k + 1 < K; k += 2) {
    acc0 = fma(a[(int64_t)k * N], vv[k], acc0);
    acc1 = fma(a[(int64_t)(k + 1) * N], vv[k + 1], acc1);
  }
  if (k < K) acc0 = fma(a[(int64_t)k * N], vv[k], acc0);
  out[(int64_t)l * N + j] = acc0 + acc1;
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
  predict_bwd_prep_kernel<T><<<dim3((unsigned)cdiv(N, 256), (unsigned)cdiv(M, rows), L), 256, 0, st>>>(A, C, q, gm, gv, gA, (T*)nullptr, M, N, rows);
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

// ---- tensor-core (tcgen05, split-TF32) variant of the same two entry points, fp32 only ----------------------------
// ws layout (floats): [Linv_lo | TT | TT_lo | T_lo | LinvT | LinvT_lo] (each L*M*M) then [sumA2 | sumC2] (each L*N).
// The column reductions of the predictive mean / variance and the row reduction gq = A gm are fused into the GEMM
// epilogues (umma_gemm.cu epi_mode 1..3), so A and C are never re-read for them.
static int tc_gemm(cudaStream_t st, int bk, int m, int n, int k, const float* A, const float* Alo, int64_t lda, int64_t sA,
                   const float* B, const float* Blo, int64_t ldb, int64_t sB, float* D, float* Dlo, int64_t ldd, int64_t sD,
                   int batch, int a_tri, int d_tri, int splitk, const UmmaEpilogue* epi = nullptr) {
  return umma_gemm_ex(bk, m, n, k, 1.0f, A, Alo, lda, sA, B, Blo, ldb, sB, nullptr, D, Dlo, ldd, sD, batch, a_tri, 0, d_tri, splitk,
                      3, epi, (void*)st);
}

__global__ void predict_var_kernel(const float* __restrict__ kxx, const float* __restrict__ sA2, const float* __restrict__ sC2,
                                   float* __restrict__ var, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) var[i] = kxx[i] - sA2[i] + sC2[i];
}

// C <- gC = 2 C gv  and its lo part
__global__ void __launch_bounds__(256) predict_scale_kernel(float* __restrict__ C, float* __restrict__ Clo, const float* __restrict__ gv,
                                                             int M, int N) {
  const int l = blockIdx.z;
  const int n4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (n4 >= N) return;
  const float4 g = *reinterpret_cast<const float4*>(gv + (int64_t)l * N + n4);
  const int m0 = blockIdx.y * 16, m1 = min(m0 + 16, M);
  for (int m = m0; m < m1; ++m) {
    const int64_t e = ((int64_t)l * M + m) * N + n4;
    float4 c = *reinterpret_cast<const float4*>(C + e);
    c.x *= 2.f * g.x; c.y *= 2.f * g.y; c.z *= 2.f * g.z; c.w *= 2.f * g.w;
    *reinterpret_cast<float4*>(C + e) = c;
    float4 lo;
    lo.x = c.x - __uint_as_float(__float_as_uint(c.x) & 0xFFFFE000u);
    lo.y = c.y - __uint_as_float(__float_as_uint(c.y) & 0xFFFFE000u);
    lo.z = c.z - __uint_as_float(__float_as_uint(c.z) & 0xFFFFE000u);
    lo.w = c.w - __uint_as_float(__float_as_uint(c.w) & 0xFFFFE000u);
    *reinterpret_cast<float4*>(Clo + e) = lo;
  }
}

static int predict_fwd_tc(const float* Kzx, const float* Kzx_lo, const float* Linv, const float* Tm, const float* q, const float* kxx,
                          float* A, float* A_lo, float* C, float* mean, float* var, float* ws, int M, int N, int L, cudaStream_t st) {
  const int64_t sMM = (int64_t)M * M, sMN = (int64_t)M * N, W = sMM * L, LN = (int64_t)L * N;
  float* Linv_lo = ws; float* TT = ws + W; float* TT_lo = ws + 2 * W;
  float* sA2 = ws + 6 * W; float* sC2 = sA2 + LN;
  int rc = gpz_tf32_lo_f32(Linv, Linv_lo, W, (void*)st);
  if (rc) return rc;
  rc = gpz_transpose_lo_f32(Tm, TT, TT_lo, M, L, (void*)st);
  if (rc) return rc;
  GPZ_CUDA(cudaMemsetAsync(sA2, 0, sizeof(float) * 2 * LN, st));
  GPZ_CUDA(cudaMemsetAsync(mean, 0, sizeof(float) * LN, st));
  // A = Linv Kzx  (lower-triangular product == TRSM Lc A = Kzx); epilogue: sum_m A^2 and mean = sum_m q_m A
  UmmaEpilogue e1{1, nullptr, q, nullptr, nullptr, sA2, mean, nullptr};
  rc = tc_gemm(st, 0, M, N, M, Linv, Linv_lo, M, sMM, Kzx, Kzx_lo, N, sMN, A, A_lo, N, sMN, L, 1, 0, 1, &e1);
  if (rc) return rc;
  // C = T^T A     (upper-triangular product); epilogue: sum_m C^2
  UmmaEpilogue e2{2, nullptr, nullptr, nullptr, nullptr, sC2, nullptr, nullptr};
  rc = tc_gemm(st, 0, M, N, M, TT, TT_lo, M, sMM, A, A_lo, N, sMN, C, nullptr, N, sMN, L, 2, 0, 1, &e2);
  if (rc) return rc;
  predict_var_kernel<<<(unsigned)cdiv(LN, 256), 256, 0, st>>>(kxx, sA2, sC2, var, LN);
  GPZ_CHECK_LAUNCH();
  return GPZ_OK;
}

static int predict_bwd_tc(const float* Kzx, const float* Kzx_lo, const float* Linv, const float* Tm, const float* q, const float* A,
                          const float* A_lo, float* C, float* C_lo, const float* gm, const float* gv, float* gA, float* gA_lo,
                          float* gKzx, float* gLinv, float* gT, float* gq, float* ws, int M, int N, int L, cudaStream_t st) {
  const int64_t sMM = (int64_t)M * M, sMN = (int64_t)M * N, W = sMM * L;
  float* T_lo = ws + 3 * W; float* LinvT = ws + 4 * W; float* LinvT_lo = ws + 5 * W;
  int rc = gpz_tf32_lo_f32(Tm, T_lo, W, (void*)st);
  if (rc) return rc;
  rc = gpz_transpose_lo_f32(Linv, LinvT, LinvT_lo, M, L, (void*)st);
  if (rc) return rc;
  // C <- gC = 2 C gv (+ lo part)
  predict_scale_kernel<<<dim3((unsigned)cdiv(N, 1024), (unsigned)cdiv(M, 16), L), 256, 0, st>>>(C, C_lo, gv, M, N);
  GPZ_CHECK_LAUNCH();
  GPZ_CUDA(cudaMemsetAsync(gq, 0, sizeof(float) * (size_t)L * M, st));
  int splitk = 1;
  {
    const int64_t tiles = cdiv(M, 128) * cdiv(M, 256) * L / 2 + 1;
    while (splitk < 32 && tiles * splitk < 148 * 2 && N / (splitk * 2) >= 1024) splitk *= 2;
  }
  // gT = tril(A gC^T)   (reduction over the N spots, both operands K-major)
  rc = tc_gemm(st, 1, M, M, N, A, A_lo, N, sMN, C, C_lo, N, sMN, gT, nullptr, M, sMM, L, 0, 1, splitk);
  if (rc) return rc;
  // gA = T gC - 2 A gv + q gm^T ;  gq = A gm     (both in the epilogue, which reads the A tile once)
  UmmaEpilogue e3{3, A, q, gv, gm, nullptr, nullptr, gq};
  rc = tc_gemm(st, 0, M, N, M, Tm, T_lo, M, sMM, C, C_lo, N, sMN, gA, gA_lo, N, sMN, L, 1, 0, 1, &e3);
  if (rc) return rc;
  // gKzx = Linv^T gA
  rc = tc_gemm(st, 0, M, N, M, LinvT, LinvT_lo, M, sMM, gA, gA_lo, N, sMN, gKzx, nullptr, N, sMN, L, 2, 0, 1);
  if (rc) return rc;
  // gLinv = tril(gA Kzx^T)
  return tc_gemm(st, 1, M, M, N, gA, gA_lo, N, sMN, Kzx, Kzx_lo, N, sMN, gLinv, nullptr, M, sMM, L, 0, 1, splitk);
}

}  // namespace gpz

using namespace gpz;

extern "C" int gpz_svgp_predict_tc_supported(int M, int N) { return (M % 4 == 0 && N % 4 == 0 && M >= 64 && N >= 256) ? 1 : 0; }

extern "C" int gpz_svgp_predict_fwd_tc_f32(const float* Kzx, const float* Kzx_lo, const float* Linv, const float* T, const float* q,
                                           const float* kxx, float* A, float* A_lo, float* C, float* mean, float* var, float* ws,
                                           int M, int N, int L, void* stream) {
  if (!gpz_svgp_predict_tc_supported(M, N) || L <= 0) return GPZ_ERR_UNSUPPORTED;
  return predict_fwd_tc(Kzx, Kzx_lo, Linv, T, q, kxx, A, A_lo, C, mean, var, ws, M, N, L, (cudaStream_t)stream);
}
extern "C" int gpz_svgp_predict_bwd_tc_f32(const float* Kzx, const float* Kzx_lo, const float* Linv, const float* T, const float* q,
                                           const float* A, const float* A_lo, float* C, float* C_lo, const float* gm, const float* gv,
                                           float* gA, float* gA_lo, float* gKzx, float* gLinv, float* gT, float* gq, float* ws, int M,
                                           int N, int L, void* stream) {
  if (!gpz_svgp_predict_tc_supported(M, N) || L <= 0) return GPZ_ERR_UNSUPPORTED;
  return predict_bwd_tc(Kzx, Kzx_lo, Linv, T, q, A, A_lo, C, C_lo, gm, gv, gA, gA_lo, gKzx, gLinv, gT, gq, ws, M, N, L,
                        (cudaStream_t)stream);
}

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

// batched matrix-vector products: trans == 0: out[l,i] = sum_k A[l,i,k] v[l,k] (A: rows x cols);  trans != 0: out[l,j] = sum_k A[l,k,j] v[l,k]
#define GPZ_GEMV_IMPL(SUF, T)                                                                                         \
  extern "C" int gpz_gemv_##SUF(int trans, const T* A, const T* v, T* out, int rows, int cols, int L, void* stream) { \
    if (rows <= 0 || cols <= 0 || L <= 0) return GPZ_OK;                                                              \
    cudaStream_t st = (cudaStream_t)stream;                                                                           \
    if (!trans) rowdot_kernel<T><<<dim3(rows, L), 256, 0, st>>>(A, v, out, rows, cols);                               \
    else coldot_kernel<T><<<dim3((unsigned)cdiv(cols, 256), L), 256, 0, st>>>(A, v, out, rows, cols);                 \
    GPZ_CHECK_LAUNCH();                                                                                               \
    return GPZ_OK;                                                                                                    \
  }
GPZ_GEMV_IMPL(f32, float)
GPZ_GEMV_IMPL(f64, double)
