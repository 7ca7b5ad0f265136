// K7: fused Poisson log-likelihood with the factor-loading contraction, forward + backward in one pass.
//
// Reference chain (likelihoods.py:49-53, 80-97, 110-145; utilities.py:611-614; torch poisson.py log_prob):
//   F = mean + eps*sd            (Normal.rsample, E samples)          E x F x B
//   rate = softplus(V) * (softplus(W) @ exp(F))                       E x G x B   (materialised by the reference)
//   ll = mean_e sum_{g,n} [ y log rate - rate (- lgamma(y+1)) ]
// and autograd's backward through all of it.  Here nothing of size G x B is ever written: one pass over
// y (G x B floats, the algorithmic HBM traffic) produces ll and the gradients w.r.t. mean, spread, W, V.
//
// Mapping: a CTA owns 128 spots (thread = spot, so y rows are read coalesced) and walks the genes in
// chunks of 32.  Phase A (thread = spot): zr, log-lik, t = d ll/d zr, accumulates d/dF in registers and
// parks t in shared memory.  Phase B (lane = gene, warp = quarter of the spots): the gW[g,f] partial
// sum_n t[g,n] exp(F)[f,n] from shared memory.  Per-CTA gW partials go to a workspace slice and are
// summed by a second tiny kernel (deterministic, no atomics on the G x F output).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "gpzoo_b200.h"

namespace gpz {

constexpr int PZ_SPOTS = 128;
constexpr int PZ_GCH = 32;

template <typename T> struct PoissonArgs {
  const T* y; int64_t y_ld;           // G x Ntot, row stride
  const int64_t* idx;                  // B or null (minibatch gather y[:, idx], V[idx])
  const T* W; int w_softplus;          // G x F raw loadings; softplus (NSF2/PNMF) or raw (Hybrid_NSF)
  const T* V;                          // Ntot raw
  const T* mean; const T* spread;      // F x B ; spread = variance for f < n_var (clamped at clamp_min), sd otherwise
  const T* eps;                        // E x F x B
  int G, F, B, E, n_var;
  T clamp_min;
  int with_lgamma;
  // outputs
  double* ll_part;                     // gridDim.x * gridDim.y
  T* gW_part;                          // gridDim.x x G x F
  T* gV;                               // B   (d ll / d V[idx[n]])
  T* gmean; T* gspread;                // F x B
  int genes_per_cta;                   // multiple of PZ_GCH
  int atomic_out;                      // gridDim.y > 1
};

// 16-byte vector loads of an FMAX-long shared-memory row (FMAX is a multiple of 4, rows are 16-byte aligned)
template <int FMAX> __device__ __forceinline__ void load_row(const float* p, float (&o)[FMAX]) {
#pragma unroll
  for (int i = 0; i < FMAX / 4; ++i) {
    const float4 v = reinterpret_cast<const float4*>(p)[i];
    o[4 * i] = v.x; o[4 * i + 1] = v.y; o[4 * i + 2] = v.z; o[4 * i + 3] = v.w;
  }
}
template <int FMAX> __device__ __forceinline__ void load_row(const double* p, double (&o)[FMAX]) {
#pragma unroll
  for (int i = 0; i < FMAX / 2; ++i) {
    const double2 v = reinterpret_cast<const double2*>(p)[i];
    o[2 * i] = v.x; o[2 * i + 1] = v.y;
  }
}

template <typename T, int FMAX>
__global__ void __launch_bounds__(PZ_SPOTS, sizeof(T) == 4 ? 3 : 1) poisson_kernel(const PoissonArgs<T> a) {
  extern __shared__ __align__(16) unsigned char pz_smem[];
  typedef T RowT[PZ_SPOTS + 1];
  typedef T RowE[FMAX];
  typedef T RowW[FMAX];
  typedef T AccW[PZ_GCH][FMAX];
  RowT* tS = reinterpret_cast<RowT*>(pz_smem);                         // [PZ_GCH][PZ_SPOTS+1]
  RowE* efS = reinterpret_cast<RowE*>(tS + PZ_GCH);                    // [PZ_SPOTS][FMAX]  (spot-major)
  RowW* sW = reinterpret_cast<RowW*>(efS + PZ_SPOTS);                  // [PZ_GCH][FMAX]
  AccW* sAcc = reinterpret_cast<AccW*>(sW + PZ_GCH);                   // [4][PZ_GCH][FMAX]
  __shared__ double red[32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.x * PZ_SPOTS + tid;
  const bool active = n < a.B;
  const bool block_full = (blockIdx.x + 1) * PZ_SPOTS <= a.B;      // uniform: every thread of the CTA has a spot
  const int64_t col = active ? (a.idx ? a.idx[n] : (int64_t)n) : 0;
  const T invE = T(1) / T(a.E);

  T mu[FMAX], sd[FMAX], gm[FMAX], gs[FMAX];
#pragma unroll
  for (int f = 0; f < FMAX; ++f) {
    gm[f] = T(0); gs[f] = T(0); mu[f] = T(0); sd[f] = T(0);
    if (f < a.F && active) {
      mu[f] = a.mean[(int64_t)f * a.B + n];
      const T s = a.spread[(int64_t)f * a.B + n];
      sd[f] = f < a.n_var ? Num<T>::sqrt(s > a.clamp_min ? s : a.clamp_min) : s;
    }
  }
  const T vraw = active ? a.V[col] : T(0);
  const T spV = active ? softplus(vraw) : T(0);
  T gVacc = T(0);
  double ll = 0.0;

  const int g_begin = blockIdx.y * a.genes_per_cta;
  const int g_end = min(a.G, g_begin + a.genes_per_cta);
  // software pipeline over groups of 8 genes: the y values of the NEXT group are in flight while the current group is
  // being processed (the sequence of groups is chunk -> sample e -> group, so with E > 1 a chunk is re-read E times)
  T ynext[8];
#pragma unroll
  for (int u = 0; u < 8; ++u)
    ynext[u] = (active && g_begin + u < g_end) ? __ldcs(a.y + (int64_t)(g_begin + u) * a.y_ld + col) : T(0);
  for (int g0 = g_begin; g0 < g_end; g0 += PZ_GCH) {
    T acc[FMAX];
#pragma unroll
    for (int f = 0; f < FMAX; ++f) acc[f] = T(0);
    __syncthreads();
    for (int e = tid; e < PZ_GCH * FMAX; e += PZ_SPOTS) {
      const int gi = e / FMAX, f = e % FMAX;
      T w = T(0);
      if (g0 + gi < g_end && f < a.F) {
        w = a.W[(int64_t)(g0 + gi) * a.F + f];
        if (a.w_softplus) w = softplus(w);
      }
      sW[gi][f] = w;
    }
    for (int e = 0; e < a.E; ++e) {
      T ef[FMAX], pg[FMAX];
#pragma unroll
      for (int f = 0; f < FMAX; ++f) {
        pg[f] = T(0);
        T v = T(0);
        if (f < a.F && active) v = Num<T>::exp(fma(a.eps[((int64_t)e * a.F + f) * a.B + n], sd[f], mu[f]));
        ef[f] = v;
      }
      __syncthreads();               // previous phase B finished with tS / efS; sW visible
#pragma unroll
      for (int f = 0; f < FMAX; ++f) efS[tid][f] = ef[f];
      // ---- phase A: thread = spot ----
      T llc = T(0);
#pragma unroll 1
      for (int gb = 0; gb < PZ_GCH; gb += 8) {
        T yv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) yv[u] = ynext[u];
        {
          // first gene of the group after this one
          int gn = g0 + gb + 8;
          if (gb + 8 == PZ_GCH) gn = (e + 1 < a.E) ? g0 : g0 + PZ_GCH;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            ynext[u] = (active && gn + u < g_end) ? __ldcs(a.y + (int64_t)(gn + u) * a.y_ld + col) : T(0);
        }
        // per-gene body; GUARD=false is the interior case (all 128 spots and all 32 genes of the chunk valid): no branches
        auto body = [&](auto guard, int u) {
          constexpr bool GUARD = decltype(guard)::value;
          const int gi = gb + u;
          T t = T(0);
          if (!GUARD || (active && g0 + gi < g_end)) {
            const T y = yv[u];
            T wrow[FMAX];
            load_row<FMAX>(&sW[gi][0], wrow);
            T zr = T(0);
#pragma unroll
            for (int f = 0; f < FMAX; ++f) zr = fma(wrow[f], ef[f], zr);
            const T r = spV * zr;
            const T ylog = y * fast_log(r);
            T lp = (y == T(0) ? T(0) : ylog) - r;
            if (a.with_lgamma && y > T(1.5)) lp -= Num<T>::lgamma(y + T(1));   // lgamma(1) = lgamma(2) = 0
            llc += lp;
            t = (fast_div(y, zr) - spV) * invE;
            gVacc += y - r;
#pragma unroll
            for (int f = 0; f < FMAX; ++f) pg[f] = fma(wrow[f], t, pg[f]);
          }
          tS[gi][tid] = t;
        };
        if (block_full && g0 + PZ_GCH <= g_end) {
#pragma unroll
          for (int u = 0; u < 8; ++u) body(std::false_type{}, u);
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) body(std::true_type{}, u);
        }
      }
      ll += (double)(llc * invE);
#pragma unroll
      for (int f = 0; f < FMAX; ++f) {
        if (f < a.F && active) {
          const T gF = ef[f] * pg[f];
          gm[f] += gF;
          gs[f] = fma(gF, a.eps[((int64_t)e * a.F + f) * a.B + n], gs[f]);
        }
      }
      __syncthreads();
      // ---- phase B: lane = gene, warp = quarter of the spots ----
      const int nb = warp * 32;
#pragma unroll 4
      for (int k = 0; k < 32; ++k) {
        const T t = tS[lane][nb + k];
        T erow[FMAX];
        load_row<FMAX>(&efS[nb + k][0], erow);
#pragma unroll
        for (int f = 0; f < FMAX; ++f) acc[f] = fma(t, erow[f], acc[f]);
      }
    }
#pragma unroll
    for (int f = 0; f < FMAX; ++f) sAcc[warp][lane][f] = acc[f];
    __syncthreads();
    for (int e = tid; e < PZ_GCH * a.F; e += PZ_SPOTS) {
      const int gi = e / a.F, f = e % a.F;
      if (g0 + gi < g_end)
        a.gW_part[((int64_t)blockIdx.x * a.G + g0 + gi) * a.F + f] =
            sAcc[0][gi][f] + sAcc[1][gi][f] + sAcc[2][gi][f] + sAcc[3][gi][f];
    }
  }

  if (active) {
    const T gv = softplus_grad(vraw) * gVacc * invE / spV;
    if (a.atomic_out) atomicAdd(a.gV + n, gv); else a.gV[n] = gv;
#pragma unroll
    for (int f = 0; f < FMAX; ++f) {
      if (f < a.F) {
        T gsp = gs[f];
        if (f < a.n_var) {                        // spread is a variance: d sd/d var = 1/(2 sd) where not clamped
          const T s = a.spread[(int64_t)f * a.B + n];
          gsp = s >= a.clamp_min ? gsp / (T(2) * sd[f]) : T(0);
        }
        const int64_t o = (int64_t)f * a.B + n;
        if (a.atomic_out) { atomicAdd(a.gmean + o, gm[f]); atomicAdd(a.gspread + o, gsp); }
        else { a.gmean[o] = gm[f]; a.gspread[o] = gsp; }
      }
    }
  }
  const double tot = block_sum<double>(ll, red);
  if (tid == 0) a.ll_part[blockIdx.y * gridDim.x + blockIdx.x] = tot;
}

// ---- v2: two spots per thread (256 spots per CTA), sample loop outermost ---------------------------------------------
// Same two-phase scheme, restructured for instruction count and occupancy (the kernel is FMA-issue bound, not HBM bound):
// each softplus(W) row fetched from shared memory serves two spots; per-thread state is only exp(F) and the d/dF
// accumulator of the current sample (the E samples are an outer loop; y is re-read per sample, from L2 when it fits).
constexpr int P2_THREADS = 128, P2_SPOTS = 256;
constexpr int PZ_GRP = 4;        // genes per prefetch group

template <typename T, int FMAX>
__global__ void __launch_bounds__(P2_THREADS, sizeof(T) == 4 ? 4 : 1) poisson_kernel2(const PoissonArgs<T> a) {
  extern __shared__ __align__(16) unsigned char pz_smem[];
  typedef T RowT[P2_SPOTS + 1];
  typedef T RowF[FMAX];
  typedef T AccW[PZ_GCH][FMAX];
  RowT* tS = reinterpret_cast<RowT*>(pz_smem);                         // [PZ_GCH][P2_SPOTS+1]   t = d ll / d zr
  RowF* efS = reinterpret_cast<RowF*>(reinterpret_cast<unsigned char*>(pz_smem) +
                                      ((sizeof(T) * PZ_GCH * (P2_SPOTS + 1) + 15) / 16) * 16);   // [P2_SPOTS][FMAX]
  RowF* sW = efS + P2_SPOTS;                                           // [PZ_GCH][FMAX]
  AccW* sAcc = reinterpret_cast<AccW*>(sW + PZ_GCH);                   // [4][PZ_GCH][FMAX]
  __shared__ double red[32];
  __shared__ T lfact[64];                     // log(y!) for integer counts y < 64 (lgamma(y+1) of torch poisson.py log_prob)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 64) lfact[tid] = Num<T>::lgamma(T(tid + 1));
  const int n0 = blockIdx.x * P2_SPOTS;
  int nn[2]; bool act[2]; int64_t col[2]; T spV[2], vraw[2], gVacc[2];
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2) {
    nn[s2] = n0 + s2 * P2_THREADS + tid;
    act[s2] = nn[s2] < a.B;
    col[s2] = act[s2] ? (a.idx ? a.idx[nn[s2]] : (int64_t)nn[s2]) : 0;
    vraw[s2] = act[s2] ? a.V[col[s2]] : T(0);
    spV[s2] = act[s2] ? softplus(vraw[s2]) : T(0);
    gVacc[s2] = T(0);
  }
  const bool block_full = n0 + P2_SPOTS <= a.B;
  const T invE = T(1) / T(a.E);
  const int g_begin = blockIdx.y * a.genes_per_cta;
  const int g_end = min(a.G, g_begin + a.genes_per_cta);
  double ll = 0.0;

  T ynext[PZ_GRP][2];
  auto prefetch = [&](int gfirst) {
#pragma unroll
    for (int u = 0; u < PZ_GRP; ++u)
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2)
        ynext[u][s2] = (act[s2] && gfirst + u < g_end) ? __ldcs(a.y + (int64_t)(gfirst + u) * a.y_ld + col[s2]) : T(0);
  };
  prefetch(g_begin);

  for (int e = 0; e < a.E; ++e) {
    T ef[2][FMAX], pg[2][FMAX];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
      for (int f = 0; f < FMAX; ++f) {
        pg[s2][f] = T(0);
        T v = T(0);
        if (f < a.F && act[s2]) {
          const int64_t o = (int64_t)f * a.B + nn[s2];
          const T sp = a.spread[o];
          const T sd = f < a.n_var ? Num<T>::sqrt(sp > a.clamp_min ? sp : a.clamp_min) : sp;
          v = Num<T>::exp(fma(a.eps[((int64_t)e * a.F + f) * a.B + nn[s2]], sd, a.mean[o]));
        }
        ef[s2][f] = v;
      }
    __syncthreads();                               // phase B of the previous sample finished with efS
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
      for (int f = 0; f < FMAX; ++f) efS[s2 * P2_THREADS + tid][f] = ef[s2][f];

    for (int g0 = g_begin; g0 < g_end; g0 += PZ_GCH) {
      __syncthreads();                             // previous chunk's phase B / flush finished with sW, tS, sAcc
      for (int i = tid; i < PZ_GCH * FMAX; i += P2_THREADS) {
        const int gi = i / FMAX, f = i % FMAX;
        T w = T(0);
        if (g0 + gi < g_end && f < a.F) {
          w = a.W[(int64_t)(g0 + gi) * a.F + f];
          if (a.w_softplus) w = softplus(w);
        }
        sW[gi][f] = w;
      }
      __syncthreads();
      // ---- phase A: thread = two spots ----
      T llc = T(0);
#pragma unroll 1
      for (int gb = 0; gb < PZ_GCH; gb += PZ_GRP) {
        T yv[PZ_GRP][2];
#pragma unroll
        for (int u = 0; u < PZ_GRP; ++u) { yv[u][0] = ynext[u][0]; yv[u][1] = ynext[u][1]; }
        {
          int gn = g0 + gb + PZ_GRP;               // first gene of the next group in (sample, chunk, group) order
          if (gb + PZ_GRP == PZ_GCH && g0 + PZ_GCH >= g_end) gn = (e + 1 < a.E) ? g_begin : g_end;
          prefetch(gn);
        }
        auto body = [&](auto guard, int u) {
          constexpr bool GUARD = decltype(guard)::value;
          const int gi = gb + u;
          T wrow[FMAX];
          load_row<FMAX>(&sW[gi][0], wrow);
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            T t = T(0);
            if (!GUARD || (act[s2] && g0 + gi < g_end)) {
              const T y = yv[u][s2];
              T z0 = T(0), z1 = T(0);
#pragma unroll
              for (int f = 0; f < FMAX; f += 2) { z0 = fma(wrow[f], ef[s2][f], z0); z1 = fma(wrow[f + 1], ef[s2][f + 1], z1); }
              const T zr = z0 + z1;
              const T r = spV[s2] * zr;
              const T ylog = y * fast_log(r);
              T lp = (y == T(0) ? T(0) : ylog) - r;
              if (a.with_lgamma) {
                const int yi = (int)y;
                // counts: table lookup (no divergent lgamma call); anything else: the real thing
                lp -= (yi >= 0 && yi < 64 && T(yi) == y) ? lfact[yi] : Num<T>::lgamma(y + T(1));
              }
              llc += lp;
              t = (fast_div(y, zr) - spV[s2]) * invE;
              gVacc[s2] += y - r;
#pragma unroll
              for (int f = 0; f < FMAX; ++f) pg[s2][f] = fma(wrow[f], t, pg[s2][f]);
            }
            tS[gi][s2 * P2_THREADS + tid] = t;
          }
        };
        if (block_full && g0 + PZ_GCH <= g_end) {
#pragma unroll
          for (int u = 0; u < PZ_GRP; ++u) body(std::false_type{}, u);
        } else {
#pragma unroll 1
          for (int u = 0; u < PZ_GRP; ++u) body(std::true_type{}, u);
        }
      }
      ll += (double)(llc * invE);
      __syncthreads();
      // ---- phase B: lane = gene, warp = a quarter (64) of the spots ----
      T acc[FMAX];
#pragma unroll
      for (int f = 0; f < FMAX; ++f) acc[f] = T(0);
      const int nb = warp * 64;
#pragma unroll 4
      for (int k = 0; k < 64; ++k) {
        const T t = tS[lane][nb + k];
        T erow[FMAX];
        load_row<FMAX>(&efS[nb + k][0], erow);
#pragma unroll
        for (int f = 0; f < FMAX; ++f) acc[f] = fma(t, erow[f], acc[f]);
      }
#pragma unroll
      for (int f = 0; f < FMAX; ++f) sAcc[warp][lane][f] = acc[f];
      __syncthreads();
      for (int i = tid; i < PZ_GCH * a.F; i += P2_THREADS) {
        const int gi = i / a.F, f = i % a.F;
        if (g0 + gi < g_end) {
          T* dst = a.gW_part + ((int64_t)blockIdx.x * a.G + g0 + gi) * a.F + f;
          const T v = sAcc[0][gi][f] + sAcc[1][gi][f] + sAcc[2][gi][f] + sAcc[3][gi][f];
          *dst = e == 0 ? v : *dst + v;
        }
      }
    }
    // ---- per-sample epilogue: d ll / d mean, d ll / d spread ----
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      if (!act[s2]) continue;
#pragma unroll
      for (int f = 0; f < FMAX; ++f) {
        if (f < a.F) {
          const int64_t o = (int64_t)f * a.B + nn[s2];
          const T gF = ef[s2][f] * pg[s2][f];
          T gsp = gF * a.eps[((int64_t)e * a.F + f) * a.B + nn[s2]];
          if (f < a.n_var) {                       // spread is a variance: d sd / d var = 1 / (2 sd) where not clamped
            const T sp = a.spread[o];
            gsp = sp >= a.clamp_min ? gsp / (T(2) * Num<T>::sqrt(sp)) : T(0);
          }
          if (a.atomic_out) { atomicAdd(a.gmean + o, gF); atomicAdd(a.gspread + o, gsp); }
          else if (e == 0) { a.gmean[o] = gF; a.gspread[o] = gsp; }
          else { a.gmean[o] += gF; a.gspread[o] += gsp; }
        }
      }
    }
  }
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2) {
    if (act[s2]) {
      const T gv = softplus_grad(vraw[s2]) * gVacc[s2] * invE / spV[s2];
      if (a.atomic_out) atomicAdd(a.gV + nn[s2], gv); else a.gV[nn[s2]] = gv;
    }
  }
  const double tot = block_sum<double>(ll, red);
  if (tid == 0) a.ll_part[blockIdx.y * gridDim.x + blockIdx.x] = tot;
}

// ---- v3 (fp32): the v2 scheme on packed fp32x2 FMAs -------------------------------------------------------------------------
// The kernel is bound by instruction issue, not by HBM (v2: 113 instructions per (gene, spot) element, 232 M warp instructions
// per launch at config 2).  Blackwell issues one 3-register FFMA per two cycles and scheduler, but FFMA2 (fma.rn.f32x2) does
// two fp32 FMAs per instruction, so the three F-long contractions per element (rate, d/dF, d/dW) run on pairs of factors:
// F = 10 -> 5 FFMA2 each, no padding to 12.  On top of that
//   * ef' = softplus(V_n) exp(F) is formed once per spot, so the rate is the contraction itself and t' = (y / r - 1) / E;
//   * sum_g r (for the log-likelihood and for d/dV) is sum_f ef'_f * (sum_g w_gf): one short contraction per 32-gene chunk
//     instead of two adds per element;
//   * log(y!) comes from a 64-entry table for integer counts (anything else takes the exact lgamma).
// Same outputs, workspace layout and grid as v2.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

template <int FP>       // FP = pairs of factors held per thread (F <= 2 FP)
__global__ void __launch_bounds__(P2_THREADS, 4) poisson_kernel3(const PoissonArgs<float> a) {
  constexpr int FPP = (FP + 1) & ~1;      // pairs per shared-memory row, padded to 16-byte multiples
  extern __shared__ __align__(16) unsigned char pz_smem[];
  typedef float RowT[P2_SPOTS + 1];
  typedef float2 RowP[FPP];
  RowT* tS = reinterpret_cast<RowT*>(pz_smem);                                           // [PZ_GCH][P2_SPOTS+1]  t' = (y/r - 1)/E
  RowP* efS = reinterpret_cast<RowP*>(pz_smem + ((sizeof(float) * PZ_GCH * (P2_SPOTS + 1) + 15) / 16) * 16);   // [P2_SPOTS][FPP]
  RowP* sW = efS + P2_SPOTS;                                                             // [PZ_GCH][FPP]
  RowP* sAcc = sW + PZ_GCH;                                                              // [4 * PZ_GCH][FPP]
  float2* sWsum = reinterpret_cast<float2*>(sAcc + 4 * PZ_GCH);                          // [FPP]  sum over the chunk's genes
  __shared__ double red[32];
  __shared__ float lfact[64];                   // log(y!) for integer counts y < 64

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 64) lfact[tid] = lgammaf((float)(tid + 1));
  const int n0 = blockIdx.x * P2_SPOTS;
  int nn[2]; bool act[2]; int64_t col[2]; float spV[2], vraw[2], ysum[2], rsum[2];
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2) {
    nn[s2] = n0 + s2 * P2_THREADS + tid;
    act[s2] = nn[s2] < a.B;
    col[s2] = act[s2] ? (a.idx ? a.idx[nn[s2]] : (int64_t)nn[s2]) : 0;
    vraw[s2] = act[s2] ? a.V[col[s2]] : 0.f;
    spV[s2] = act[s2] ? softplus(vraw[s2]) : 0.f;
    ysum[s2] = 0.f; rsum[s2] = 0.f;
  }
  const bool block_full = n0 + P2_SPOTS <= a.B;
  const float invE = 1.f / (float)a.E;
  const int g_begin = blockIdx.y * a.genes_per_cta;
  const int g_end = min(a.G, g_begin + a.genes_per_cta);
  double ll = 0.0;
  constexpr float LN2 = 0.69314718055994530942f;

  float ynext[PZ_GRP][2];
  auto prefetch = [&](int gfirst) {
#pragma unroll
    for (int u = 0; u < PZ_GRP; ++u)
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2)
        ynext[u][s2] = (act[s2] && gfirst + u < g_end) ? __ldcs(a.y + (int64_t)(gfirst + u) * a.y_ld + col[s2]) : 0.f;
  };
  prefetch(g_begin);

  for (int e = 0; e < a.E; ++e) {
    float2 ef[2][FP], pg[2][FP];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
      for (int fp = 0; fp < FP; ++fp) {
        pg[s2][fp] = make_float2(0.f, 0.f);
        float v[2] = {0.f, 0.f};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int f = 2 * fp + h;
          if (f < a.F && act[s2]) {
            const int64_t o = (int64_t)f * a.B + nn[s2];
            const float sp = a.spread[o];
            const float sd = f < a.n_var ? sqrtf(sp > a.clamp_min ? sp : a.clamp_min) : sp;
            v[h] = spV[s2] * expf(fmaf(a.eps[((int64_t)e * a.F + f) * a.B + nn[s2]], sd, a.mean[o]));
          }
        }
        ef[s2][fp] = make_float2(v[0], v[1]);
      }
    __syncthreads();                               // phase B of the previous sample finished with efS
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
#pragma unroll
      for (int fp = 0; fp < FP; ++fp) efS[s2 * P2_THREADS + tid][fp] = ef[s2][fp];
      if (FPP > FP) efS[s2 * P2_THREADS + tid][FPP - 1] = make_float2(0.f, 0.f);
    }

    for (int g0 = g_begin; g0 < g_end; g0 += PZ_GCH) {
      __syncthreads();                             // previous chunk's phase B / flush finished with sW, tS, sAcc
      for (int i = tid; i < PZ_GCH * FPP; i += P2_THREADS) {
        const int gi = i / FPP, fp = i % FPP;
        float w[2] = {0.f, 0.f};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int f = 2 * fp + h;
          if (g0 + gi < g_end && f < a.F) {
            w[h] = a.W[(int64_t)(g0 + gi) * a.F + f];
            if (a.w_softplus) w[h] = softplus(w[h]);
          }
        }
        sW[gi][fp] = make_float2(w[0], w[1]);
      }
      __syncthreads();
      if (tid < FPP) {                             // column sums of the chunk's loadings
        float2 sacc = make_float2(0.f, 0.f);
        for (int gi = 0; gi < PZ_GCH; ++gi) { const float2 w = sW[gi][tid]; sacc.x += w.x; sacc.y += w.y; }
        sWsum[tid] = sacc;
      }
      // ---- phase A: thread = two spots ----
      float llc = 0.f;
#pragma unroll 1
      for (int gb = 0; gb < PZ_GCH; gb += PZ_GRP) {
        float yv[PZ_GRP][2];
#pragma unroll
        for (int u = 0; u < PZ_GRP; ++u) { yv[u][0] = ynext[u][0]; yv[u][1] = ynext[u][1]; }
        {
          int gn = g0 + gb + PZ_GRP;               // first gene of the next group in (sample, chunk, group) order
          if (gb + PZ_GRP == PZ_GCH && g0 + PZ_GCH >= g_end) gn = (e + 1 < a.E) ? g_begin : g_end;
          prefetch(gn);
        }
        auto body = [&](auto guard, int u) {
          constexpr bool GUARD = decltype(guard)::value;
          const int gi = gb + u;
          float2 wrow[FPP];
#pragma unroll
          for (int i = 0; i < FPP / 2; ++i) {
            const float4 v = reinterpret_cast<const float4*>(&sW[gi][0])[i];
            wrow[2 * i] = make_float2(v.x, v.y); wrow[2 * i + 1] = make_float2(v.z, v.w);
          }
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            float t = 0.f;
            if (!GUARD || (act[s2] && g0 + gi < g_end)) {
              const float y = yv[u][s2];
              float2 z = make_float2(0.f, 0.f);
#pragma unroll
              for (int fp = 0; fp < FP; ++fp) z = ffma2(wrow[fp], ef[s2][fp], z);
              const float r = z.x + z.y;                                  // rate (softplus(V) is inside ef)
              const float ylog = (y * LN2) * __log2f(r);
              float lp = y == 0.f ? 0.f : ylog;
              if (a.with_lgamma) {
                const int yi = (int)y;
                lp -= (yi >= 0 && yi < 64 && (float)yi == y) ? lfact[yi] : lgammaf(y + 1.f);
              }
              llc += lp;
              ysum[s2] += y;
              t = fmaf(y * invE, __frcp_rn(r), -invE);                    // (y / r - 1) / E
              const float2 t2 = make_float2(t, t);
#pragma unroll
              for (int fp = 0; fp < FP; ++fp) pg[s2][fp] = ffma2(wrow[fp], t2, pg[s2][fp]);
            }
            tS[gi][s2 * P2_THREADS + tid] = t;
          }
        };
        if (block_full && g0 + PZ_GCH <= g_end) {
#pragma unroll
          for (int u = 0; u < PZ_GRP; ++u) body(std::false_type{}, u);
        } else {
#pragma unroll 1
          for (int u = 0; u < PZ_GRP; ++u) body(std::true_type{}, u);
        }
      }
      __syncthreads();
      // sum over the chunk's genes of the rate, per spot: sum_f ef'_f * Wsum_f
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        float2 z = make_float2(0.f, 0.f);
#pragma unroll
        for (int fp = 0; fp < FP; ++fp) z = ffma2(sWsum[fp], ef[s2][fp], z);
        rsum[s2] += z.x + z.y;
        llc -= act[s2] ? (z.x + z.y) : 0.f;
      }
      ll += (double)(llc * invE);
      // ---- phase B: lane = gene, warp = a quarter (64) of the spots ----
      float2 acc[FP];
#pragma unroll
      for (int fp = 0; fp < FP; ++fp) acc[fp] = make_float2(0.f, 0.f);
      const int nb = warp * 64;
#pragma unroll 4
      for (int k = 0; k < 64; ++k) {
        const float t = tS[lane][nb + k];
        const float2 t2 = make_float2(t, t);
        float2 erow[FPP];
#pragma unroll
        for (int i = 0; i < FPP / 2; ++i) {
          const float4 v = reinterpret_cast<const float4*>(&efS[nb + k][0])[i];
          erow[2 * i] = make_float2(v.x, v.y); erow[2 * i + 1] = make_float2(v.z, v.w);
        }
#pragma unroll
        for (int fp = 0; fp < FP; ++fp) acc[fp] = ffma2(t2, erow[fp], acc[fp]);
      }
#pragma unroll
      for (int fp = 0; fp < FP; ++fp) sAcc[warp * PZ_GCH + lane][fp] = acc[fp];
      __syncthreads();
      for (int i = tid; i < PZ_GCH * a.F; i += P2_THREADS) {
        const int gi = i / a.F, f = i % a.F;
        if (g0 + gi < g_end) {
          float* dst = a.gW_part + ((int64_t)blockIdx.x * a.G + g0 + gi) * a.F + f;
          float v = 0.f;
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) {
            const float2 pr = sAcc[w4 * PZ_GCH + gi][f >> 1];
            v += (f & 1) ? pr.y : pr.x;
          }
          *dst = e == 0 ? v : *dst + v;
        }
      }
    }
    // ---- per-sample epilogue: d ll / d mean, d ll / d spread ----
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      if (!act[s2]) continue;
#pragma unroll
      for (int fp = 0; fp < FP; ++fp) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int f = 2 * fp + h;
          if (f < a.F) {
            const int64_t o = (int64_t)f * a.B + nn[s2];
            const float gF = h ? ef[s2][fp].y * pg[s2][fp].y : ef[s2][fp].x * pg[s2][fp].x;
            float gsp = gF * a.eps[((int64_t)e * a.F + f) * a.B + nn[s2]];
            if (f < a.n_var) {                       // spread is a variance: d sd / d var = 1 / (2 sd) where not clamped
              const float sp = a.spread[o];
              gsp = sp >= a.clamp_min ? gsp / (2.f * sqrtf(sp)) : 0.f;
            }
            if (a.atomic_out) { atomicAdd(a.gmean + o, gF); atomicAdd(a.gspread + o, gsp); }
            else if (e == 0) { a.gmean[o] = gF; a.gspread[o] = gsp; }
            else { a.gmean[o] += gF; a.gspread[o] += gsp; }
          }
        }
      }
    }
  }
#pragma unroll
  for (int s2 = 0; s2 < 2; ++s2) {
    if (act[s2]) {
      // d ll / d V = softplus'(V) / softplus(V) * mean_e sum_g (y - r);  ysum was accumulated once per sample
      const float gv = softplus_grad(vraw[s2]) * (ysum[s2] - rsum[s2]) * invE / spV[s2];
      if (a.atomic_out) atomicAdd(a.gV + nn[s2], gv); else a.gV[nn[s2]] = gv;
    }
  }
  const double tot = block_sum<double>(ll, red);
  if (tid == 0) a.ll_part[blockIdx.y * gridDim.x + blockIdx.x] = tot;
}

template <int FP> constexpr size_t poisson_smem3() {
  constexpr int FPP = (FP + 1) & ~1;
  return ((sizeof(float) * PZ_GCH * (P2_SPOTS + 1) + 15) / 16) * 16 + sizeof(float2) * FPP * (P2_SPOTS + PZ_GCH + 4 * PZ_GCH + 1);
}
template <int FP> int poisson_launch3(const PoissonArgs<float>& a, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = poisson_smem3<FP>();
  GPZ_CUDA(cudaFuncSetAttribute(poisson_kernel3<FP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  poisson_kernel3<FP><<<grid, P2_THREADS, smem, st>>>(a);
  GPZ_CHECK_LAUNCH();
  return GPZ_OK;
}


// ---- v4 (fp32, F <= 16): the three F-long contractions on the tensor cores ------------------------------------------------
// v2 / v3 are bound by instruction issue (3.3 warp instructions per (gene, spot) element, 36 of the 105 thread instructions being
// the FMAs of the contractions); HBM sits at 12 %.  Here the contractions are warp-level MMAs (mma.sync m16n8k16, bf16 operands,
// fp32 accumulation) chained in registers the way a fused attention kernel chains QK^T -> P -> PV:
//   C1[n, g] = EF'^T[n, :] Wm[g, :]^T                      rate tile (16 spots x 8 genes), EF' = softplus(V) exp(F)
//   t'       = (y / C1 - 1) / E,  ll += y log2 C1          element-wise on the accumulator fragment, y read straight into that layout
//   D2[n, l] += t'^T[n, g] Wm[g, l]                        the packed C1 fragments of two gene tiles ARE the A fragment (d ll / d EF')
//   D3[l, g] += EF'[l, n] t'^T[n, g]                       B fragment = movmatrix-transposed packed t' (d ll / d Wm, over the spots)
// Every operand is split x = hi + lo into two bf16 (|x - hi - lo| <= 2^-18 |x|) and a product is hi*hi + hi*lo + lo*hi, so a
// contraction carries ~1e-5 relative error at worst (1e-6 measured on the gradients) - bf16 rather than fp16 because exp(F) and y / r
// have fp32 range.  Three launches:
//   prep  per gene: softplus(W) as bf16 hi / lo planes [G][16] + fp32 column sums per gene range; per spot and sample: EF' as bf16
//         planes [E][B][16]; zeroes the D2 accumulator.  Everything that is per gene or per spot is done ONCE here, not once per CTA.
//   main  grid (256-spot blocks) x (gene ranges).  A warp owns 32 spots (two 16-spot sub-tiles, D2 resident in registers over the
//         CTA's gene range) and walks the range in 16-gene blocks; D3 is summed over the sub-tiles in registers, over the warps through
//         shared memory once per 64-gene chunk, and leaves as per-CTA partials (same workspace layout as v2).  D2 leaves by vector
//         atomics into [E][B][16].  ~0.6 warp instructions per element instead of 3.3.
//   post  gW = softplus'(W) * sum of the partials; per spot: gF = EF' D2 -> d/dmean, d/dspread, d/dV, and the terms that never needed
//         the G x B elements: sum_g r = sum_l EF'_l (sum_g Wm_gl), d ll / d V = softplus'(V) / softplus(V) sum_l EF'_l D2_l; the last
//         block to finish adds up the log-likelihood partials.
constexpr int P4_WARPS = 8, P4_THREADS = 32 * P4_WARPS, P4_SUB = 2, P4_SPOTS = P4_WARPS * P4_SUB * 16;
constexpr int P4_GCH = 64;            // genes per staged chunk of loadings (four 16-gene blocks)
constexpr int P4_WLD = 24;            // bf16 per shared-memory row of loadings: 48-byte rows keep ldmatrix conflict-free
constexpr int P4_ALD = 72;            // floats per row of the D3 exchange buffer: float2 stores of a half-warp hit 32 distinct banks

__device__ __forceinline__ uint32_t pack_bf16x2(float e0, float e1) {      // e0 -> bits 0..15, e1 -> bits 16..31
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(e1), "f"(e0));
  return d;
}
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(x0, x1);
  lo = pack_bf16x2(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xffff0000u));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movmatrix_t(uint32_t a) {              // 8x8 b16 transpose across the warp
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&d)[4], const void* p) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_t(uint32_t (&d)[4], const void* p) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(addr));
}

// workspace of the v4 path behind the ll / gW partials
struct Poisson4Ws {
  uint16_t* wh; uint16_t* wl;          // [G][16]   softplus(W), bf16 hi / lo (factors >= F are zero)
  float* wsum;                         // [nby][16] column sums of softplus(W) over the gene range of grid row y
  uint16_t* efh; uint16_t* efl;        // [E][B][16] EF' = softplus(V) exp(mean + eps sd), bf16 hi / lo
  float* d2;                           // [E][B][16] sum_g t'[g, n] Wm[g, l]
  double* rs_part;                     // [post spot blocks] sum over the block's spots of sum_l EF'_l Wsum_l
  unsigned int* ticket;                // last-block-done counter of the post kernel
};

// EF'[l, n] of sample e for the four factors lq .. lq + 3 of spot n (exact fp32; used by prep and by post).  All twelve loads are issued
// before anything is computed (factor indices clamped instead of predicated): one memory latency, not twelve.
__device__ __forceinline__ void poisson_ef4(const PoissonArgs<float>& a, int e, int n, int lq, float spV, float (&ev)[4], float (&epsv)[4],
                                            float (&sdv)[4]) {
  float sp[4], mn[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int f = min(lq + j, a.F - 1);
    const int64_t o = (int64_t)f * a.B + n;
    sp[j] = a.spread[o];
    mn[j] = a.mean[o];
    epsv[j] = a.eps[((int64_t)e * a.F + f) * a.B + n];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int f = lq + j;
    sdv[j] = f < a.n_var ? sqrtf(sp[j] > a.clamp_min ? sp[j] : a.clamp_min) : sp[j];
    ev[j] = f < a.F ? spV * expf(fmaf(epsv[j], sdv[j], mn[j])) : 0.f;
  }
}

// blocks [0, nby): loadings of one gene range; blocks [nby, ...): 64 spots of one sample each (thread <-> spot tid / 4, factors 4 (tid & 3)..+3)
__global__ void __launch_bounds__(256) poisson_prep4_kernel(const PoissonArgs<float> a, const Poisson4Ws w, int nby, int spot_blocks) {
  __shared__ float part[8][16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lq = (tid & 3) * 4;
  if ((int)blockIdx.x < nby) {
    if (blockIdx.x == 0 && tid == 0) *w.ticket = 0u;
    const int g_begin = blockIdx.x * a.genes_per_cta, g_end = min(a.G, g_begin + a.genes_per_cta);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int g = g_begin + (tid >> 2); g < g_end; g += 64) {
      float wv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        wv[j] = 0.f;
        if (lq + j < a.F) {
          wv[j] = a.W[(int64_t)g * a.F + lq + j];
          if (a.w_softplus) wv[j] = softplus(wv[j]);
        }
        acc[j] += wv[j];
      }
      uint32_t h0, l0, h1, l1;
      split_bf16x2(wv[0], wv[1], h0, l0);
      split_bf16x2(wv[2], wv[3], h1, l1);
      *reinterpret_cast<uint2*>(w.wh + (int64_t)g * 16 + lq) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(w.wl + (int64_t)g * 16 + lq) = make_uint2(l0, l1);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[j];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 4) part[warp][lq + j] = v;
    }
    __syncthreads();
    if (tid < 16) {
      float v = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) v += part[w8][tid];
      w.wsum[blockIdx.x * 16 + tid] = v;
    }
    return;
  }
  const int sb = blockIdx.x - nby, e = sb / spot_blocks, n = (sb - e * spot_blocks) * 64 + (tid >> 2);
  if (n >= a.B) return;
  const float spV = softplus(a.V[a.idx ? a.idx[n] : (int64_t)n]);
  float ev[4], epsv[4], sdv[4];
  poisson_ef4(a, e, n, lq, spV, ev, epsv, sdv);
  uint32_t h0, l0, h1, l1;
  split_bf16x2(ev[0], ev[1], h0, l0);
  split_bf16x2(ev[2], ev[3], h1, l1);
  const int64_t o = ((int64_t)e * a.B + n) * 16 + lq;
  *reinterpret_cast<uint2*>(w.efh + o) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(w.efl + o) = make_uint2(l0, l1);
  *reinterpret_cast<float4*>(w.d2 + o) = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Counts as stored by the caller: fp32 (the reference's format), or uint8 / int16 / int32 (lossless for counts, 4x / 2x less HBM
// traffic and upload).  y_load2 fetches the two neighbouring spots of a fragment row pair with one load.
// Small unsigned integers become floats on the ALU / FMA pipes (2^23 + k has k in its low mantissa bits), not through I2F, which
// shares the quarter-rate XU pipe with the kernel's log2 / reciprocal.
__device__ __forceinline__ float u16_to_float(unsigned k) { return __uint_as_float(0x4B000000u | k) - 8388608.f; }
template <typename YT> __device__ __forceinline__ float y_load1(const YT* p) { return (float)__ldcs(p); }
template <> __device__ __forceinline__ float y_load1<uint8_t>(const uint8_t* p) { return u16_to_float(__ldcs(p)); }
template <> __device__ __forceinline__ float y_load1<int16_t>(const int16_t* p) {
  return u16_to_float(__ldcs(reinterpret_cast<const unsigned short*>(p)));              // counts: non-negative by contract
}
template <typename YT> __device__ __forceinline__ void y_load2(const YT* p, float& a, float& b);
template <> __device__ __forceinline__ void y_load2<float>(const float* p, float& a, float& b) {
  const float2 v = __ldcs(reinterpret_cast<const float2*>(p)); a = v.x; b = v.y;
}
template <> __device__ __forceinline__ void y_load2<uint8_t>(const uint8_t* p, float& a, float& b) {
  const unsigned v = __ldcs(reinterpret_cast<const unsigned short*>(p)); a = u16_to_float(v & 0xffu); b = u16_to_float(v >> 8);
}
template <> __device__ __forceinline__ void y_load2<int16_t>(const int16_t* p, float& a, float& b) {
  const unsigned v = __ldcs(reinterpret_cast<const unsigned*>(p)); a = u16_to_float(v & 0xffffu); b = u16_to_float(v >> 16);
}
template <> __device__ __forceinline__ void y_load2<int32_t>(const int32_t* p, float& a, float& b) {
  const int2 v = __ldcs(reinterpret_cast<const int2*>(p)); a = (float)v.x; b = (float)v.y;
}

__device__ __noinline__ float lfact_slow(float y) { return lgammaf(y + 1.f); }     // counts >= 64 or non-integer y: rare, kept out of line

template <bool HAS_IDX, typename YT>
__global__ void __launch_bounds__(P4_THREADS, 2) poisson_kernel4(const PoissonArgs<float> a, const Poisson4Ws w) {
  __shared__ __align__(16) uint16_t sWh[P4_GCH][P4_WLD];      // softplus(W) chunk, bf16 hi / lo, [gene][factor]
  __shared__ __align__(16) uint16_t sWl[P4_GCH][P4_WLD];
  __shared__ __align__(16) float sAcc[P4_WARPS][16][P4_ALD];  // per-warp D3[l][gene of the chunk]
  __shared__ double red[32];
  __shared__ float lfact[64];                                 // log(y!) for integer counts y < 64

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = lane >> 2, c = lane & 3;                      // fragment coordinates (groupID, threadID_in_group)
  if (a.with_lgamma && tid < 64) lfact[tid] = lgammaf((float)(tid + 1));
  const int nw = blockIdx.x * P4_SPOTS + warp * (16 * P4_SUB);          // first spot of the warp
  const bool block_full = (blockIdx.x + 1) * P4_SPOTS <= a.B;
  // fragment row r (+8) of sub-tile s <-> spot nw + 16 s + 2 r (+1): a thread's two spots of a sub-tile are neighbours, so one 8-byte load
  // per gene row fetches both (the row <-> spot assignment inside a 16-spot tile is free as long as EF' and D2 use the same one)
  const bool fast_cols = !HAS_IDX && block_full && (a.y_ld & 1) == 0 && (reinterpret_cast<uintptr_t>(a.y) & (2 * sizeof(YT) - 1)) == 0;
  const float invE = 1.f / (float)a.E;
  const int g_begin = blockIdx.y * a.genes_per_cta;
  const int g_end = min(a.G, g_begin + a.genes_per_cta);
  constexpr float LN2 = 0.69314718055994530942f;

  // the thread's spots: sub-tile s, half h -> n = nw + 16 s + 2 r + h ; invalid spots read a valid column of y and have EF' = 0
  int col[P4_SUB][2];                                         // columns of y: 32-bit (Ntot < 2^31)
  if (!fast_cols) {
#pragma unroll
    for (int s = 0; s < P4_SUB; ++s)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int nc = min(nw + 16 * s + 2 * r + h, a.B - 1);
        col[s][h] = HAS_IDX ? (int)a.idx[nc] : nc;
      }
  }

  // y in accumulator-fragment layout: register q = 4 tt + i of sub-tile s <-> gene gblk + 8 tt + 2c + (i & 1), spot half i >> 1.
  // Rows past the gene range re-read the range's last row (masked when consumed: only the last block of a range is ragged), so the
  // loads carry no predicate and no select and really are in flight for a whole block.
  float yreg[P4_SUB][8];
  const YT* ybase = reinterpret_cast<const YT*>(a.y);      // a.y carries the caller's pointer whatever its element type
  const int ld = (int)a.y_ld;                          // a row of y is shorter than 2^31 entries
  const int ld8 = 8 * ld;
  auto load_y = [&](int s, int gblk) {
    if (fast_cols && gblk + 16 <= g_end) {
      // interior block: the four rows 2c, 2c + 1, 2c + 8, 2c + 9 from one base pointer (one 32 x 32 -> 64-bit multiply-add, three adds)
      const YT* p0 = ybase + (int64_t)(gblk + 2 * c) * ld + (nw + 16 * s + 2 * r);
      const YT* p1 = p0 + ld;
      const YT* p2 = p0 + ld8;
      const YT* p3 = p2 + ld;
      y_load2<YT>(p0, yreg[s][0], yreg[s][2]); y_load2<YT>(p1, yreg[s][1], yreg[s][3]);
      y_load2<YT>(p2, yreg[s][4], yreg[s][6]); y_load2<YT>(p3, yreg[s][5], yreg[s][7]);
      return;
    }
    // ragged block / gathered columns: four clamped row pointers
    const YT* rowp[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = min(gblk + 8 * (j >> 1) + 2 * c + (j & 1), g_end - 1);
      rowp[j] = ybase + (int64_t)g * ld;
    }
    if (!HAS_IDX && block_full) {
#pragma unroll
      for (int q = 0; q < 8; ++q) yreg[s][q] = y_load1<YT>(rowp[2 * (q >> 2) + (q & 1)] + (nw + 16 * s + 2 * r) + ((q >> 1) & 1));
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) yreg[s][q] = y_load1<YT>(rowp[2 * (q >> 2) + (q & 1)] + col[s][(q >> 1) & 1]);
    }
  };
#pragma unroll
  for (int s = 0; s < P4_SUB; ++s) load_y(s, g_begin);

  double ll = 0.0;
  for (int e = 0; e < a.E; ++e) {
    // operand fragments of the sample: A1 = EF'^T [16 spots x 16 factors] (rate), A3 = EF' [16 factors x 16 spots] (d/dW);
    // register k of A1 <-> spot half k & 1, factors 2c + 8 (k >> 1) + {0, 1}
    uint32_t a1h[P4_SUB][4], a1l[P4_SUB][4], a3h[P4_SUB][4], a3l[P4_SUB][4];
    float d2[P4_SUB][2][4];
#pragma unroll
    for (int s = 0; s < P4_SUB; ++s) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int n = nw + 16 * s + 2 * r + (k & 1);
        a1h[s][k] = 0u; a1l[s][k] = 0u;
        if (n < a.B) {
          const int64_t o = (((int64_t)e * a.B + n) * 16 + 2 * c + 8 * (k >> 1)) >> 1;
          a1h[s][k] = reinterpret_cast<const uint32_t*>(w.efh)[o];
          a1l[s][k] = reinterpret_cast<const uint32_t*>(w.efl)[o];
        }
      }
    }
#pragma unroll
    for (int s = 0; s < P4_SUB; ++s) {
      a3h[s][0] = movmatrix_t(a1h[s][0]); a3h[s][1] = movmatrix_t(a1h[s][2]);
      a3h[s][2] = movmatrix_t(a1h[s][1]); a3h[s][3] = movmatrix_t(a1h[s][3]);
      a3l[s][0] = movmatrix_t(a1l[s][0]); a3l[s][1] = movmatrix_t(a1l[s][2]);
      a3l[s][2] = movmatrix_t(a1l[s][1]); a3l[s][3] = movmatrix_t(a1l[s][3]);
#pragma unroll
      for (int lt = 0; lt < 2; ++lt)
#pragma unroll
        for (int i = 0; i < 4; ++i) d2[s][lt][i] = 0.f;
    }

    for (int g0 = g_begin; g0 < g_end; g0 += P4_GCH) {
      __syncthreads();                              // the previous chunk's reduction is done with sAcc; every warp is done with sW
      {                                             // stage the chunk's loadings: thread <-> (gene tid / 4, factors 4 (tid & 3) ..+3)
        const int gi = tid >> 2, lq = (tid & 3) * 4;
        uint2 vh = make_uint2(0u, 0u), vl = make_uint2(0u, 0u);
        if (g0 + gi < g_end) {
          vh = *reinterpret_cast<const uint2*>(w.wh + (int64_t)(g0 + gi) * 16 + lq);
          vl = *reinterpret_cast<const uint2*>(w.wl + (int64_t)(g0 + gi) * 16 + lq);
        }
        *reinterpret_cast<uint2*>(&sWh[gi][lq]) = vh;
        *reinterpret_cast<uint2*>(&sWl[gi][lq]) = vl;
      }
      __syncthreads();
      const int nblk = min(P4_GCH / 16, (g_end - g0 + 15) >> 4);
      float llc = 0.f, llg = 0.f;
#pragma unroll 1
      for (int jb = 0; jb < nblk; ++jb) {
        const int gblk = g0 + 16 * jb;
        // next block in (sample, gene) order, for the y prefetch
        int gnext = gblk + 16;
        if (gnext >= g_end) gnext = (e + 1 < a.E) ? g_begin : -1;
        const bool ragged = gblk + 16 > g_end || !block_full;          // uniform: some of the 16 x 32 elements are not there
        // ldmatrix row addresses.  rate: B[k = factor][n = gene], matrices (genes 0-7, l 0-7), (genes 0-7, l 8-15), (genes 8-15, l 0-7),
        // (genes 8-15, l 8-15); d/dEF': B[k = gene][n = factor], transposed on load: (genes 0-7, l 0-7), (genes 8-15, l 0-7), (genes 0-7,
        // l 8-15), (genes 8-15, l 8-15).  The fragments are (re)loaded per sub-tile right before their MMAs to keep them short-lived.
        const int mi = lane >> 3, row = lane & 7;
        const int o1 = (16 * jb + (mi >> 1) * 8 + row) * P4_WLD + (mi & 1) * 8;
        const int o2 = (16 * jb + (mi & 1) * 8 + row) * P4_WLD + (mi >> 1) * 8;
        float d3[2][4];
#pragma unroll
        for (int tt = 0; tt < 2; ++tt)
#pragma unroll
          for (int i = 0; i < 4; ++i) d3[tt][i] = 0.f;
#pragma unroll
        for (int s = 0; s < P4_SUB; ++s) {
          float c1[2][4];
          {
            uint32_t b1h[4], b1l[4];
            ldmatrix_x4(b1h, &sWh[0][0] + o1);
            ldmatrix_x4(b1l, &sWl[0][0] + o1);
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
#pragma unroll
              for (int i = 0; i < 4; ++i) c1[tt][i] = 0.f;
              mma_bf16(c1[tt], a1l[s], b1h[2 * tt], b1h[2 * tt + 1]);
              mma_bf16(c1[tt], a1h[s], b1l[2 * tt], b1l[2 * tt + 1]);
              mma_bf16(c1[tt], a1h[s], b1h[2 * tt], b1h[2 * tt + 1]);
            }
          }
          if (ragged) {                             // genes past the range / spots past the minibatch carry no count
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (gblk + 8 * (q >> 2) + 2 * c + (q & 1) >= g_end || nw + 16 * s + 2 * r + ((q >> 1) & 1) >= a.B) yreg[s][q] = 0.f;
          }
          // element-wise on the rate fragment
          float t[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float y = yreg[s][q];
            const float rate = fmaxf(c1[q >> 2][q & 3], 1e-30f);
            float rinv;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(rate));
            llc = fmaf(y, __log2f(rate), llc);
            t[q] = fmaf(y * invE, rinv, -invE);
          }
          if (a.with_lgamma) {
            // log(y!) from the table, branch-free: the entry of min(trunc(y), 63), exact whenever y is an integer count below 64
            // (lfact[0] = lfact[1] = 0); anything else (y >= 64, non-integer y) is detected by comparing back and redone exactly.
            bool exact = true;
            float lg8 = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float y = yreg[s][q];
              const int yi = (int)min((unsigned)__float2int_rz(y), 63u);       // negative or huge -> 63, which fails the comparison
              exact = exact && ((float)yi == y);
              lg8 += lfact[yi];
            }
            if (!exact) {
              lg8 = 0.f;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float y = yreg[s][q];
                const int yi = (int)min((unsigned)__float2int_rz(y), 63u);
                lg8 += ((float)yi == y) ? lfact[yi] : lfact_slow(y);
              }
            }
            llg += lg8;
          }
          if (gnext >= 0) load_y(s, gnext);         // the sub-tile's y registers are free: fetch the next block's values into them
          // t' as the A fragment [16 spots x 16 genes]: a0 = tile 0 rows r, a1 = tile 0 rows r + 8, a2 / a3 = tile 1
          uint32_t th[4], tl[4];
          split_bf16x2(t[0], t[1], th[0], tl[0]);
          split_bf16x2(t[2], t[3], th[1], tl[1]);
          split_bf16x2(t[4], t[5], th[2], tl[2]);
          split_bf16x2(t[6], t[7], th[3], tl[3]);
          {
            uint32_t b2h[4], b2l[4];
            ldmatrix_x4_t(b2h, &sWh[0][0] + o2);
            ldmatrix_x4_t(b2l, &sWl[0][0] + o2);
#pragma unroll
            for (int lt = 0; lt < 2; ++lt) {
              mma_bf16(d2[s][lt], tl, b2h[2 * lt], b2h[2 * lt + 1]);
              mma_bf16(d2[s][lt], th, b2l[2 * lt], b2l[2 * lt + 1]);
              mma_bf16(d2[s][lt], th, b2h[2 * lt], b2h[2 * lt + 1]);
            }
          }
          // t'^T as B fragments [16 spots x 8 genes] of the two gene tiles
#pragma unroll
          for (int tt = 0; tt < 2; ++tt) {
            const uint32_t bh0 = movmatrix_t(th[2 * tt]), bh1 = movmatrix_t(th[2 * tt + 1]);
            const uint32_t bl0 = movmatrix_t(tl[2 * tt]), bl1 = movmatrix_t(tl[2 * tt + 1]);
            mma_bf16(d3[tt], a3l[s], bh0, bh1);
            mma_bf16(d3[tt], a3h[s], bl0, bl1);
            mma_bf16(d3[tt], a3h[s], bh0, bh1);
          }
        }
        // D3[l = r (+8)][gene = 16 jb + 8 tt + 2c (+1)] of this warp's 32 spots
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          *reinterpret_cast<float2*>(&sAcc[warp][r][16 * jb + 8 * tt + 2 * c]) = make_float2(d3[tt][0], d3[tt][1]);
          *reinterpret_cast<float2*>(&sAcc[warp][r + 8][16 * jb + 8 * tt + 2 * c]) = make_float2(d3[tt][2], d3[tt][3]);
        }
      }
      ll += (double)((llc * LN2 - llg) * invE);
      __syncthreads();
      const int nvalid = min(P4_GCH, g_end - g0) * a.F;
      for (int i = tid; i < nvalid; i += P4_THREADS) {
        const int gi = i / a.F, f = i - gi * a.F;
        float v = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < P4_WARPS; ++w8) v += sAcc[w8][f][gi];
        float* dst = a.gW_part + ((int64_t)blockIdx.x * a.G + g0 + gi) * a.F + f;
        *dst = e == 0 ? v : *dst + v;
      }
    }
    // D2[n][l] of the sample leaves the registers: (n = r + 8 (i >> 1), l = 8 lt + 2c + (i & 1)) pairs as float2
#pragma unroll
    for (int s = 0; s < P4_SUB; ++s)
#pragma unroll
      for (int lt = 0; lt < 2; ++lt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int n = nw + 16 * s + 2 * r + h;
          if (n < a.B) {
            float2* dst = reinterpret_cast<float2*>(w.d2 + ((int64_t)e * a.B + n) * 16 + 8 * lt + 2 * c);
            const float2 v = make_float2(d2[s][lt][2 * h], d2[s][lt][2 * h + 1]);
            if (a.atomic_out) atomicAdd(dst, v); else *dst = v;
          }
        }
  }
  const double tot = block_sum<double>(ll, red);
  if (tid == 0) a.ll_part[blockIdx.y * gridDim.x + blockIdx.x] = tot;
}

// blocks [0, gw_blocks): gW[g, f] = (softplus'(W[g, f]) or 1) * sum_b gW_part[b, g, f]; blocks [gw_blocks, ...): 64 spots each, all samples:
// gF = EF' D2 -> d/dmean, d/dspread, d/dV and sum_g r.  The last block to finish sums the log-likelihood partials (fixed order).
__global__ void __launch_bounds__(256) poisson_post4_kernel(const PoissonArgs<float> a, const Poisson4Ws w, float* __restrict__ gW, int nbx,
                                                            int nby, int gw_blocks, int spot_blocks, double* __restrict__ ll_out) {
  __shared__ double red[32];
  __shared__ float wsum[16];
  __shared__ bool last;
  const int tid = threadIdx.x;
  if ((int)blockIdx.x < gw_blocks) {
    // 64 outputs per block; thread (o, p) adds the partials of the CTAs b = p, p + 4, ... (independent loads, 256-byte rows)
    __shared__ float part[4][64];
    const int64_t GF = (int64_t)a.G * a.F, i = (int64_t)blockIdx.x * 64 + (tid & 63);
    const int p = tid >> 6;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (i < GF) {
      int b = p;
      for (; b + 12 < nbx; b += 16) {
        s0 += a.gW_part[(int64_t)b * GF + i];
        s1 += a.gW_part[(int64_t)(b + 4) * GF + i];
        s2 += a.gW_part[(int64_t)(b + 8) * GF + i];
        s3 += a.gW_part[(int64_t)(b + 12) * GF + i];
      }
      for (; b < nbx; b += 4) s0 += a.gW_part[(int64_t)b * GF + i];
    }
    part[p][tid & 63] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (tid < 64 && i < GF) {
      const float s = (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
      gW[i] = a.w_softplus ? s * softplus_grad(a.W[i]) : s;
    }
  } else {
    if (tid < 16) {
      float v = 0.f;
      for (int by = 0; by < nby; ++by) v += w.wsum[by * 16 + tid];
      wsum[tid] = v;
    }
    __syncthreads();
    const int sb = blockIdx.x - gw_blocks, n = sb * 64 + (tid >> 2), lq = (tid & 3) * 4;
    float rs = 0.f, gsum = 0.f, vraw = 0.f, spV = 1.f;
    if (n < a.B) {
      vraw = a.V[a.idx ? a.idx[n] : (int64_t)n];
      spV = softplus(vraw);
      float gm[4] = {0.f, 0.f, 0.f, 0.f}, gs[4] = {0.f, 0.f, 0.f, 0.f}, sdv[4];
      for (int e = 0; e < a.E; ++e) {
        float ev[4], epsv[4];
        poisson_ef4(a, e, n, lq, spV, ev, epsv, sdv);
        const float4 d = *reinterpret_cast<const float4*>(w.d2 + ((int64_t)e * a.B + n) * 16 + lq);
        const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float gF = ev[j] * dv[j];
          gm[j] += gF;
          gs[j] = fmaf(gF, epsv[j], gs[j]);
          gsum += gF;
          rs = fmaf(ev[j], wsum[lq + j], rs);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int f = lq + j;
        if (f < a.F) {
          const int64_t o = (int64_t)f * a.B + n;
          float gsp = gs[j];
          if (f < a.n_var) gsp = a.spread[o] >= a.clamp_min ? gsp / (2.f * sdv[j]) : 0.f;   // d sd / d var = 1 / (2 sd) where not clamped
          a.gmean[o] = gm[j];
          a.gspread[o] = gsp;
        }
      }
    }
    // d ll / d V[n] = softplus'(V) / softplus(V) * sum_e sum_l gF_e[l, n]   (sum_g (y - r) = E sum_l EF'_l D2_l; 1/E is inside t')
    gsum += __shfl_xor_sync(0xffffffffu, gsum, 1);
    gsum += __shfl_xor_sync(0xffffffffu, gsum, 2);
    if (n < a.B && (tid & 3) == 0) a.gV[n] = softplus_grad(vraw) * gsum / spV;
    const double tot = block_sum<double>((double)rs, red);
    if (tid == 0) w.rs_part[sb] = tot;
  }
  // last block done: ll = sum(main partials) - sum(rate sums) / E
  __threadfence();
  __syncthreads();
  if (tid == 0) last = atomicAdd(w.ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  double s = 0.0;
  for (int i = tid; i < nbx * nby; i += 256) s += __ldcg(a.ll_part + i);
  double rsum = 0.0;
  for (int i = tid; i < spot_blocks; i += 256) rsum += __ldcg(w.rs_part + i);
  s -= rsum / (double)a.E;
  s = block_sum<double>(s, red);
  if (tid == 0) ll_out[0] = s;
}

// gW[g,f] = (softplus'(W[g,f]) or 1) * sum_b gW_part[b,g,f]
template <typename T>
__global__ void poisson_gw_reduce_kernel(const T* __restrict__ part, const T* __restrict__ W, T* __restrict__ gW, int64_t GF, int nb,
                                         int w_softplus) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= GF) return;
  T s = T(0);
  for (int b = 0; b < nb; ++b) s += part[(int64_t)b * GF + e];
  gW[e] = w_softplus ? s * softplus_grad(W[e]) : s;
}

__global__ void sum_double_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  s = block_sum<double>(s, red);
  if (threadIdx.x == 0) out[0] = s;
}

template <typename T, int FMAX> constexpr size_t poisson_smem() {
  return sizeof(T) * (PZ_GCH * (PZ_SPOTS + 1) + FMAX * PZ_SPOTS + PZ_GCH * FMAX + 4 * PZ_GCH * FMAX);
}
template <typename T, int FMAX> constexpr size_t poisson_smem2() {
  return ((sizeof(T) * PZ_GCH * (P2_SPOTS + 1) + 15) / 16) * 16 + sizeof(T) * (FMAX * P2_SPOTS + PZ_GCH * FMAX + 4 * PZ_GCH * FMAX);
}
static int poisson_version() {          // GPZ_POISSON_V: 1 = thread-per-spot kernel, 2 = two spots per thread, 3 = 2 on packed fp32x2 FMAs (fp32 only; measured equal to 2: both are bound by bookkeeping instructions), 4 (default) = tensor-core contractions (fp32, F <= 16; anything else runs 2)
  static int v = -1;
  if (v < 0) { const char* e = getenv("GPZ_POISSON_V"); v = e ? atoi(e) : 4; }
  return v;
}
static bool poisson_use_v4(int F, bool is_f32) { return is_f32 && F <= 16 && poisson_version() >= 4; }
static bool poisson_v2_forced() { return poisson_version() == 2; }
template <typename T, int FMAX> int poisson_launch(const PoissonArgs<T>& a, dim3 grid, cudaStream_t st) {
  if (poisson_version() >= 2) {
    constexpr size_t smem = poisson_smem2<T, FMAX>();
    GPZ_CUDA(cudaFuncSetAttribute(poisson_kernel2<T, FMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    poisson_kernel2<T, FMAX><<<grid, P2_THREADS, smem, st>>>(a);
  } else {
    constexpr size_t smem = poisson_smem<T, FMAX>();
    GPZ_CUDA(cudaFuncSetAttribute(poisson_kernel<T, FMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    poisson_kernel<T, FMAX><<<grid, PZ_SPOTS, smem, st>>>(a);
  }
  GPZ_CHECK_LAUNCH();
  return GPZ_OK;
}

static void poisson_grid(int G, int F, int B, bool is_f32, int* nbx, int* nby, int* gpc) {
  if (poisson_use_v4(F, is_f32)) {        // 256 spots per CTA, gene ranges in whole 16-gene blocks, about four waves of 2 CTAs per SM
    *nbx = (int)cdiv(B, P4_SPOTS);
    const int blocks = (int)cdiv(G, 16);
    int by = 1;
    if (*nbx < 1184) by = (int)min((int64_t)blocks, (int64_t)max(1, 1184 / (*nbx > 0 ? *nbx : 1)));
    const int bpc = (int)cdiv(blocks, by);
    *gpc = bpc * 16;
    *nby = (int)cdiv(G, *gpc);
    return;
  }
  *nbx = (int)cdiv(B, poisson_version() >= 2 ? P2_SPOTS : PZ_SPOTS);
  const int chunks = (int)cdiv(G, PZ_GCH);
  int by = 1;
  if (*nbx < 1184) by = (int)min((int64_t)chunks, cdiv(1184, *nbx > 0 ? *nbx : 1));
  const int cpc = (int)cdiv(chunks, by);
  *gpc = cpc * PZ_GCH;
  *nby = (int)cdiv(G, *gpc);
}

// workspace: ll partials (double) | gW partials [nbx][G][F] | (v4) the Poisson4Ws arrays
static size_t poisson_ll_slots(int nbx, int nby) { return ((size_t)nbx * nby + 2 + 1) & ~(size_t)1; }   // keeps what follows 16-byte aligned
static size_t poisson_ws_bytes(int G, int F, int B, int E, bool is_f32, size_t elt, int nbx, int nby) {
  size_t b = sizeof(double) * poisson_ll_slots(nbx, nby) + (elt * (size_t)nbx * G * F + 15) / 16 * 16;
  if (poisson_use_v4(F, is_f32)) {
    const size_t EB16 = 16 * (size_t)(E > 0 ? E : 1) * B;
    b += 2 * sizeof(uint16_t) * 16 * (size_t)G + 2 * sizeof(uint16_t) * EB16 + sizeof(float) * EB16 + sizeof(float) * 16 * (size_t)nby +
         sizeof(double) * (size_t)cdiv(B, 64) + 16;
  }
  return b;
}

// y_kind: element type of y — 0 = T (the reference's format), 1 = uint8, 2 = int16, 3 = int32 (fp32 tensor-core kernel only)
template <typename T>
int poisson_fwdbwd(PoissonArgs<T> a, T* gW, double* ll_out, void* ws, size_t ws_bytes, cudaStream_t st, int y_kind = 0) {
  if (a.F < 1 || a.F > 32 || a.E < 1 || a.G < 1) return GPZ_ERR_UNSUPPORTED;
  if (y_kind < 0 || y_kind > 3 || (y_kind != 0 && !poisson_use_v4(a.F, std::is_same<T, float>::value))) return GPZ_ERR_UNSUPPORTED;
  if (a.B == 0) {                       // empty minibatch: ll = 0, no gradient
    GPZ_CUDA(cudaMemsetAsync(ll_out, 0, sizeof(double), st));
    GPZ_CUDA(cudaMemsetAsync(gW, 0, sizeof(T) * (size_t)a.G * a.F, st));
    return GPZ_OK;
  }
  int nbx, nby, gpc;
  poisson_grid(a.G, a.F, a.B, std::is_same<T, float>::value, &nbx, &nby, &gpc);
  const size_t need = poisson_ws_bytes(a.G, a.F, a.B, a.E, std::is_same<T, float>::value, sizeof(T), nbx, nby);
  if (ws_bytes < need) return GPZ_ERR_BADARG;
  a.ll_part = reinterpret_cast<double*>(ws);
  a.gW_part = reinterpret_cast<T*>(a.ll_part + poisson_ll_slots(nbx, nby));
  a.genes_per_cta = gpc;
  a.atomic_out = nby > 1;
  if (a.atomic_out && !poisson_use_v4(a.F, std::is_same<T, float>::value)) {
    GPZ_CUDA(cudaMemsetAsync(a.gV, 0, sizeof(T) * a.B, st));
    GPZ_CUDA(cudaMemsetAsync(a.gmean, 0, sizeof(T) * a.B * a.F, st));
    GPZ_CUDA(cudaMemsetAsync(a.gspread, 0, sizeof(T) * a.B * a.F, st));
  }
  dim3 grid(nbx, nby);
  int rc;
  if constexpr (std::is_same<T, float>::value) {
    if (poisson_use_v4(a.F, true)) {       // fp32, F <= 16: contractions on the tensor cores
      unsigned char* p4 = reinterpret_cast<unsigned char*>(a.gW_part) + (sizeof(T) * (size_t)nbx * a.G * a.F + 15) / 16 * 16;
      const size_t EB16 = 16 * (size_t)a.E * a.B;
      const int spot_blocks = (int)cdiv(a.B, 64), gw_blocks = (int)cdiv((int64_t)a.G * a.F, 64);
      Poisson4Ws w;
      w.d2 = reinterpret_cast<float*>(p4); p4 += sizeof(float) * EB16;
      w.wsum = reinterpret_cast<float*>(p4); p4 += sizeof(float) * 16 * (size_t)nby;
      w.rs_part = reinterpret_cast<double*>(p4); p4 += sizeof(double) * (size_t)spot_blocks;       // 8-byte aligned: every block above is a multiple of 64 B
      w.efh = reinterpret_cast<uint16_t*>(p4); p4 += sizeof(uint16_t) * EB16;
      w.efl = reinterpret_cast<uint16_t*>(p4); p4 += sizeof(uint16_t) * EB16;
      w.wh = reinterpret_cast<uint16_t*>(p4); p4 += sizeof(uint16_t) * 16 * (size_t)a.G;
      w.wl = reinterpret_cast<uint16_t*>(p4); p4 += sizeof(uint16_t) * 16 * (size_t)a.G;
      w.ticket = reinterpret_cast<unsigned int*>(p4);
      poisson_prep4_kernel<<<nby + a.E * spot_blocks, 256, 0, st>>>(a, w, nby, spot_blocks);
      GPZ_CHECK_LAUNCH();
#define GPZ_P4_LAUNCH(YT)                                                              \
  do {                                                                                 \
    if (a.idx) poisson_kernel4<true, YT><<<grid, P4_THREADS, 0, st>>>(a, w);           \
    else poisson_kernel4<false, YT><<<grid, P4_THREADS, 0, st>>>(a, w);                \
  } while (0)
      if (y_kind == 1) GPZ_P4_LAUNCH(uint8_t);
      else if (y_kind == 2) GPZ_P4_LAUNCH(int16_t);
      else if (y_kind == 3) GPZ_P4_LAUNCH(int32_t);
      else GPZ_P4_LAUNCH(float);
#undef GPZ_P4_LAUNCH
      GPZ_CHECK_LAUNCH();
      poisson_post4_kernel<<<gw_blocks + spot_blocks, 256, 0, st>>>(a, w, gW, nbx, nby, gw_blocks, spot_blocks, ll_out);
      GPZ_CHECK_LAUNCH();
      return GPZ_OK;
    }
    if (poisson_version() == 3) {          // fp32: the packed-FMA kernel
      if (a.F <= 4) rc = poisson_launch3<2>(a, grid, st);
      else if (a.F <= 10) rc = poisson_launch3<5>(a, grid, st);
      else if (a.F <= 12) rc = poisson_launch3<6>(a, grid, st);
      else if (a.F <= 20) rc = poisson_launch3<10>(a, grid, st);
      else rc = poisson_launch3<16>(a, grid, st);
      if (rc) return rc;
      const int64_t GF3 = (int64_t)a.G * a.F;
      poisson_gw_reduce_kernel<T><<<(unsigned)cdiv(GF3, 256), 256, 0, st>>>(a.gW_part, a.W, gW, GF3, nbx, a.w_softplus);
      GPZ_CHECK_LAUNCH();
      sum_double_kernel<<<1, 256, 0, st>>>(a.ll_part, nbx * nby, ll_out);
      GPZ_CHECK_LAUNCH();
      return GPZ_OK;
    }
  }
  if (a.F <= 4) rc = poisson_launch<T, 4>(a, grid, st);
  else if (a.F <= 12) rc = poisson_launch<T, 12>(a, grid, st);
  else if (a.F <= 20) rc = poisson_launch<T, 20>(a, grid, st);
  else rc = poisson_launch<T, 32>(a, grid, st);
  if (rc) return rc;
  const int64_t GF = (int64_t)a.G * a.F;
  poisson_gw_reduce_kernel<T><<<(unsigned)cdiv(GF, 256), 256, 0, st>>>(a.gW_part, a.W, gW, GF, nbx, a.w_softplus);
  GPZ_CHECK_LAUNCH();
  sum_double_kernel<<<1, 256, 0, st>>>(a.ll_part, nbx * nby, ll_out);
  GPZ_CHECK_LAUNCH();
  return GPZ_OK;
}

// rate[e,g,n] = softplus(V[col]) * sum_f w(W[g,f]) exp(F[e,f,n]) -- the compatibility path that materialises
// pY.rate for callers of model.forward() (likelihoods.py:83-85); not used by the fused ELBO.
template <typename T>
__global__ void poisson_rate_kernel(const T* __restrict__ W, int w_softplus, const T* __restrict__ V, const int64_t* __restrict__ idx,
                                    const T* __restrict__ Fs, T* __restrict__ rate, int G, int F, int B, int E) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = blockIdx.z;
  if (n >= B) return;
  const T spV = softplus(V[idx ? idx[n] : (int64_t)n]);
  const int g0 = blockIdx.y * 32, g1 = min(G, g0 + 32);
  for (int g = g0; g < g1; ++g) {
    T zr = T(0);
    for (int f = 0; f < F; ++f) {
      T w = W[(int64_t)g * F + f];
      if (w_softplus) w = softplus(w);
      zr = fma(w, Num<T>::exp(Fs[((int64_t)e * F + f) * B + n]), zr);
    }
    rate[((int64_t)e * G + g) * B + n] = spV * zr;
  }
}

}  // namespace gpz

using namespace gpz;

#define GPZ_POISSON_IMPL(SUF, T)                                                                                       \
  extern "C" int64_t gpz_poisson_workspace_bytes_##SUF(int G, int F, int B, int E) {                                   \
    int nbx, nby, gpc;                                                                                                 \
    poisson_grid(G, F, B, sizeof(T) == 4, &nbx, &nby, &gpc);                                                           \
    return (int64_t)poisson_ws_bytes(G, F, B, E, sizeof(T) == 4, sizeof(T), nbx, nby);                                 \
  }                                                                                                                    \
  extern "C" int gpz_poisson_fwdbwd_##SUF(const T* y, int64_t y_ld, const int64_t* idx, const T* W, int w_softplus,    \
                                          const T* V, const T* mean, const T* spread, const T* eps, int G, int F,      \
                                          int B, int E, int n_var, T clamp_min, int with_lgamma, double* ll, T* gW,    \
                                          T* gV, T* gmean, T* gspread, void* ws, int64_t ws_bytes, void* stream) {     \
    PoissonArgs<T> a;                                                                                                  \
    a.y = y; a.y_ld = y_ld; a.idx = idx; a.W = W; a.w_softplus = w_softplus; a.V = V; a.mean = mean;                   \
    a.spread = spread; a.eps = eps; a.G = G; a.F = F; a.B = B; a.E = E; a.n_var = n_var; a.clamp_min = clamp_min;      \
    a.with_lgamma = with_lgamma; a.gV = gV; a.gmean = gmean; a.gspread = gspread;                                      \
    return poisson_fwdbwd<T>(a, gW, ll, ws, (size_t)ws_bytes, (cudaStream_t)stream);                                   \
  }                                                                                                                    \
  extern "C" int gpz_poisson_rate_##SUF(const T* W, int w_softplus, const T* V, const int64_t* idx, const T* Fs,       \
                                        T* rate, int G, int F, int B, int E, void* stream) {                           \
    if (B == 0 || G == 0 || E == 0) return GPZ_OK;                                                                     \
    poisson_rate_kernel<T><<<dim3((unsigned)cdiv(B, 128), (unsigned)cdiv(G, 32), E), 128, 0, (cudaStream_t)stream>>>(  \
        W, w_softplus, V, idx, Fs, rate, G, F, B, E);                                                                  \
    GPZ_CHECK_LAUNCH();                                                                                                \
    return GPZ_OK;                                                                                                     \
  }

GPZ_POISSON_IMPL(f32, float)
GPZ_POISSON_IMPL(f64, double)

// counts stored as uint8 / int16 / int32 (y_kind 1 / 2 / 3; 0 = float): same outputs and workspace as gpz_poisson_fwdbwd_f32
extern "C" int gpz_poisson_fwdbwd_yt_f32(const void* y, int y_kind, int64_t y_ld, const int64_t* idx, const float* W, int w_softplus,
                                         const float* V, const float* mean, const float* spread, const float* eps, int G, int F, int B,
                                         int E, int n_var, float clamp_min, int with_lgamma, double* ll, float* gW, float* gV,
                                         float* gmean, float* gspread, void* ws, int64_t ws_bytes, void* stream) {
  PoissonArgs<float> a;
  a.y = reinterpret_cast<const float*>(y); a.y_ld = y_ld; a.idx = idx; a.W = W; a.w_softplus = w_softplus; a.V = V; a.mean = mean;
  a.spread = spread; a.eps = eps; a.G = G; a.F = F; a.B = B; a.E = E; a.n_var = n_var; a.clamp_min = clamp_min;
  a.with_lgamma = with_lgamma; a.gV = gV; a.gmean = gmean; a.gspread = gspread;
  return poisson_fwdbwd<float>(a, gW, ll, ws, (size_t)ws_bytes, (cudaStream_t)stream, y_kind);
}
