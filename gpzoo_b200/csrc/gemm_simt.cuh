// Batched register-tiled GEMM on CUDA cores (fp32 and fp64), with triangular-structure skipping and
// split-K.  This is the exact-arithmetic workhorse: it serves every fp64 (1e-10 parity) contraction,
// all O(M^3) glue (Cholesky trailing updates, triangular inverse, whitening, Cholesky backward) and is
// the fp32 fallback for shapes the tcgen05 path (umma_gemm.cu) does not take.
//
//   D[b] = alpha * op(A[b]) (m x k) * op(B[b]) (k x n) + beta * D[b]        all row-major
//   op(A) = A (stored m x k) or A^T (stored k x m);  op(B) = B (stored k x n) or B^T (stored n x k)
//
// a_tri / b_tri declare op(A) / op(B) lower(1) or upper(2) triangular so whole k-blocks are skipped;
// d_tri = 1/2 computes only the lower/upper triangle of D (other elements keep beta*D).
#pragma once
#include "common.cuh"

namespace gpz {

template <typename T> struct GemmParams {
  const T* A; const T* B; T* D;
  int m, n, k;
  int64_t lda, ldb, ldd;
  int64_t sAo, sAi, sBo, sBi, sDo, sDi;   // batch strides: outer (batch / batch_inner), inner (batch % batch_inner)
  int batch, batch_inner;
  T alpha, beta;
  int a_tri, b_tri, d_tri;
  int splitk;                             // >1: D += alpha*acc with atomics (caller pre-initialises D; beta ignored)
};

template <typename T> struct GemmCfg;
template <> struct GemmCfg<float>  { static constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8; };
template <> struct GemmCfg<double> { static constexpr int BM = 64,  BN = 64,  BK = 16, TM = 4, TN = 4; };

template <int TMN, int BMN> __device__ __forceinline__ int frag_pos(int t, int r) {
  if (TMN == 8) return (r >> 2) * (BMN / 2) + t * 4 + (r & 3);
  return t * TMN + r;
}

template <typename T, bool TA, bool TB>
__global__ void __launch_bounds__(256, 2) gemm_kernel(const GemmParams<T> p) {
  using C = GemmCfg<T>;
  constexpr int BM = C::BM, BN = C::BN, BK = C::BK, TM = C::TM, TN = C::TN;
  constexpr int NT = 256, PAD = 4;
  constexpr int LA = BM * BK / NT, LB = BN * BK / NT;
  __shared__ T As[2][BK][BM + PAD];
  __shared__ T Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int z = blockIdx.z, split = z % p.splitk, b = z / p.splitk;
  const int bo = b / p.batch_inner, bi = b % p.batch_inner;
  const T* __restrict__ A = p.A + bo * p.sAo + bi * p.sAi;
  const T* __restrict__ B = p.B + bo * p.sBo + bi * p.sBi;
  T* __restrict__ D = p.D + bo * p.sDo + bi * p.sDi;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  const int i1 = min(i0 + BM, p.m), j1 = min(j0 + BN, p.n);

  if ((p.d_tri == 1 && j0 >= i1) || (p.d_tri == 2 && i0 >= j1)) return;   // tile entirely in the unused triangle

  int k_lo = 0, k_hi = p.k;
  if (p.a_tri == 1) k_hi = min(k_hi, i1);
  if (p.a_tri == 2) k_lo = max(k_lo, i0);
  if (p.b_tri == 1) k_lo = max(k_lo, j0);
  if (p.b_tri == 2) k_hi = min(k_hi, j1);
  const int kb0 = k_lo / BK;
  int nkt = k_hi > kb0 * BK ? (k_hi - kb0 * BK + BK - 1) / BK : 0;
  int kt_begin = 0, kt_end = nkt;
  if (p.splitk > 1) {
    const int chunk = (nkt + p.splitk - 1) / p.splitk;
    kt_begin = split * chunk;
    kt_end = min(nkt, kt_begin + chunk);
    if (kt_begin >= kt_end) return;
  }

  T acc[TM][TN];
#pragma unroll
  for (int r = 0; r < TM; ++r)
#pragma unroll
    for (int c = 0; c < TN; ++c) acc[r][c] = T(0);

  T ra[LA], rb[LB];
  auto gload = [&](int kt) {
    const int kbase = (kb0 + kt) * BK;
#pragma unroll
    for (int s = 0; s < LA; ++s) {
      const int e = tid + s * NT;
      int i, kk;
      if (TA) { kk = e / BM; i = e % BM; } else { i = e / BK; kk = e % BK; }
      const int gi = i0 + i, gk = kbase + kk;
      T v = T(0);
      if (gi < p.m && gk >= k_lo && gk < k_hi) v = TA ? A[(int64_t)gk * p.lda + gi] : A[(int64_t)gi * p.lda + gk];
      ra[s] = v;
    }
#pragma unroll
    for (int s = 0; s < LB; ++s) {
      const int e = tid + s * NT;
      int j, kk;
      if (TB) { j = e / BK; kk = e % BK; } else { kk = e / BN; j = e % BN; }
      const int gj = j0 + j, gk = kbase + kk;
      T v = T(0);
      if (gj < p.n && gk >= k_lo && gk < k_hi) v = TB ? B[(int64_t)gj * p.ldb + gk] : B[(int64_t)gk * p.ldb + gj];
      rb[s] = v;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int s = 0; s < LA; ++s) {
      const int e = tid + s * NT;
      int i, kk;
      if (TA) { kk = e / BM; i = e % BM; } else { i = e / BK; kk = e % BK; }
      As[buf][kk][i] = ra[s];
    }
#pragma unroll
    for (int s = 0; s < LB; ++s) {
      const int e = tid + s * NT;
      int j, kk;
      if (TB) { j = e / BK; kk = e % BK; } else { kk = e / BN; j = e % BN; }
      Bs[buf][kk][j] = rb[s];
    }
  };

  gload(kt_begin);
  sstore(0);
  __syncthreads();
  for (int kt = kt_begin; kt < kt_end; ++kt) {
    const int buf = (kt - kt_begin) & 1;
    if (kt + 1 < kt_end) gload(kt + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      T a[TM], bb[TN];
#pragma unroll
      for (int r = 0; r < TM; ++r) a[r] = As[buf][kk][frag_pos<TM, BM>(ty, r)];
#pragma unroll
      for (int c = 0; c < TN; ++c) bb[c] = Bs[buf][kk][frag_pos<TN, BN>(tx, c)];
#pragma unroll
      for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int c = 0; c < TN; ++c) acc[r][c] = fma(a[r], bb[c], acc[r][c]);
    }
    if (kt + 1 < kt_end) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int r = 0; r < TM; ++r) {
    const int gi = i0 + frag_pos<TM, BM>(ty, r);
    if (gi >= p.m) continue;
#pragma unroll
    for (int c = 0; c < TN; ++c) {
      const int gj = j0 + frag_pos<TN, BN>(tx, c);
      if (gj >= p.n) continue;
      if ((p.d_tri == 1 && gj > gi) || (p.d_tri == 2 && gj < gi)) continue;
      T* d = D + (int64_t)gi * p.ldd + gj;
      if (p.splitk > 1) atomicAdd(d, p.alpha * acc[r][c]);
      else *d = p.beta == T(0) ? p.alpha * acc[r][c] : p.alpha * acc[r][c] + p.beta * (*d);
    }
  }
}

template <typename T>
int gemm_launch(const GemmParams<T>& p, bool ta, bool tb, cudaStream_t st) {
  using C = GemmCfg<T>;
  if (p.m <= 0 || p.n <= 0 || p.batch <= 0) return GPZ_OK;
  GemmParams<T> q = p;
  if (q.splitk < 1) q.splitk = 1;
  if (q.batch_inner < 1) q.batch_inner = 1;
  dim3 grid((unsigned)cdiv(p.n, C::BN), (unsigned)cdiv(p.m, C::BM), (unsigned)(p.batch * q.splitk));
  if (grid.y > 65535u || grid.z > 65535u) return GPZ_ERR_UNSUPPORTED;
  if (ta && tb) gemm_kernel<T, true, true><<<grid, 256, 0, st>>>(q);
  else if (ta) gemm_kernel<T, true, false><<<grid, 256, 0, st>>>(q);
  else if (tb) gemm_kernel<T, false, true><<<grid, 256, 0, st>>>(q);
  else gemm_kernel<T, false, false><<<grid, 256, 0, st>>>(q);
  GPZ_CHECK_LAUNCH();
  return GPZ_OK;
}

// Convenience: simple strided batch.
template <typename T>
int gemm(cudaStream_t st, bool ta, bool tb, int m, int n, int k, T alpha, const T* A, int64_t lda, int64_t sA,
         const T* B, int64_t ldb, int64_t sB, T beta, T* D, int64_t ldd, int64_t sD, int batch,
         int a_tri = 0, int b_tri = 0, int d_tri = 0, int splitk = 1) {
  GemmParams<T> p;
  p.A = A; p.B = B; p.D = D; p.m = m; p.n = n; p.k = k; p.lda = lda; p.ldb = ldb; p.ldd = ldd;
  p.sAo = sA; p.sBo = sB; p.sDo = sD; p.sAi = p.sBi = p.sDi = 0; p.batch = batch; p.batch_inner = 1;
  p.alpha = alpha; p.beta = beta; p.a_tri = a_tri; p.b_tri = b_tri; p.d_tri = d_tri; p.splitk = splitk;
  return gemm_launch(p, ta, tb, st);
}

}  // namespace gpz
