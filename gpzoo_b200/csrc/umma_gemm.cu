// tcgen05 / TMEM / TMA batched GEMM in split-TF32 ("3xTF32") arithmetic for the fp32 hot path.
//
// The big contractions of the SVGP step (the triangular products that replace cholesky_solve, gp.py:218, and
// svgp_forward, utilities.py:392-395, plus their backward) are genuine dense GEMMs.  The parity target (1e-4 of the
// fp64 reference on gradients) rules out single-pass TF32 (10-bit mantissa), so every operand x travels as the pair
// (x, lo) with lo = x - tf32_trunc(x) and each k-step issues three tensor-core MMAs into the same TMEM accumulator:
//     D += A*B_lo ;  D += A_lo*B ;  D += A*B          (the tensor core truncates A, B to tf32 itself)
// which recovers ~2^-21 relative accuracy per product at 1/3 of the TF32 rate.
//
//   D[b] (m x n) = alpha * A[b] (m x k, K-major: row-major, k contiguous) * op(B[b]) (+ Cin[b])
//   op(B): B stored k x n with n contiguous ("MN-major", the layout of Kzx / A / gC / gA: L x M x N), or
//          B stored n x k with k contiguous ("K-major", used for the reductions over the N spots).
//
// One CTA = one 128 x 256 output tile.  Warp 0: TMA producer (cp.async.bulk.tensor into a 4-stage ring of 48 KB
// stages, 64B-swizzled K-major tiles / 128B(32B-atom)-swizzled MN-major tiles).  Warp 1: allocates 256 TMEM columns and issues
// tcgen05.mma.kind::tf32 (M=128, N=256, K=8) from one thread, releasing stages with tcgen05.commit.  Warps 2-5:
// epilogue, tcgen05.ld of the fp32 accumulator -> alpha/Cin/lo-split -> global (or fp32 atomics for split-K).
// Triangular operands skip whole k-blocks; lower-triangular outputs skip whole tiles.
#include <cuda.h>

#include <atomic>
#include <cstdlib>

#include "common.cuh"
#include "gpzoo_b200.h"
#include "umma_gemm.h"

namespace gpz {
namespace umma {

constexpr int BM = 128, BN = 256, BK = 16, STAGES = 4;
constexpr int A_BYTES = BM * BK * 4;          // 8 KB   (128 rows x 64 B, SWIZZLE_64B)
constexpr int B_BYTES = BN * BK * 4;          // 16 KB  (MN-major: 8 chunks x 16 rows x 128 B; K-major: 256 rows x 64 B)
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // raw + lo of both operands = 48 KB
constexpr int EPI_LD = 36;                    // padded row length (floats) of the per-warp 32 x 32 epilogue staging tile
constexpr int EPI_BYTES = 4 * 32 * EPI_LD * 4;  // 4 epilogue warps
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_BYTES;
constexpr int NTHREADS = 192;
constexpr uint32_t TMEM_COLS = 512;      // two 128 x 256 fp32 accumulators

struct Params {
  float* D; float* Dlo; const float* Cin;
  int m, n, k;
  int64_t ldd, sD;
  int batch, splitk;
  int a_tri, b_tri, d_tri;
  int n_terms;            // 3: split-TF32, 1: plain TF32
  int b_map4d;            // MN-major B loaded with one 4-D box per stage (needs n % 32 == 0)
  // fused epilogues of the SVGP predictive op (csrc/predict.cu); vectors are per batch entry
  //   1: col1[n] += sum_i D[i,n]^2 ; col2[n] += sum_i rowv[i] D[i,n]                     (A = Linv Kzx: sum A^2 and mean)
  //   2: col1[n] += sum_i D[i,n]^2                                                        (C = T^T A: sum C^2)
  //   3: D = acc - 2 Aux[i,n] colv1[n] + rowv[i] colv2[n] ; rowacc[i] += sum_n Aux[i,n] colv2[n]   (gA and gq = A gm)
  int epi_mode;
  const float* Aux; const float* rowv; const float* colv1; const float* colv2;
  float* col1; float* col2; float* rowacc;
  float alpha;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
      "%25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout type [61,64)  (1 = SWIZZLE_128B_BASE32B, 2 = SWIZZLE_128B, 4 = SWIZZLE_64B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

struct TileInfo {
  int b, i0, j0, kb0, nkb;     // batch entry, tile origin, first k-block, number of k-blocks (0 = no MMA work)
  int zero;                    // nkb == 0 but the tile is part of the output: the epilogue writes alpha*0 (+Cin)
};

// tile index -> coordinates.  mt (row tile) is the fastest index so that consecutively scheduled tiles share the same
// B column block (L2 reuse); the dynamic scheduler balances the unequal k-ranges of triangular operands.
__device__ __forceinline__ TileInfo tile_info(const Params& p, int t, int mtiles, int ntiles) {
  TileInfo ti;
  const int mt = t % mtiles;
  int r = t / mtiles;
  const int nt = r % ntiles;
  r /= ntiles;
  const int split = r % p.splitk;
  ti.b = r / p.splitk;
  ti.i0 = mt * BM;
  ti.j0 = nt * BN;
  const int i1 = min(ti.i0 + BM, p.m), j1 = min(ti.j0 + BN, p.n);
  ti.kb0 = 0;
  ti.nkb = 0;
  ti.zero = 0;
  if ((p.d_tri == 1 && ti.j0 >= i1) || (p.d_tri == 2 && ti.i0 >= j1)) return ti;
  int k_lo = 0, k_hi = p.k;
  if (p.a_tri == 1) k_hi = min(k_hi, i1);
  if (p.a_tri == 2) k_lo = max(k_lo, ti.i0);
  if (p.b_tri == 1) k_lo = max(k_lo, ti.j0);
  if (p.b_tri == 2) k_hi = min(k_hi, j1);
  const int kb0 = k_lo / BK;
  const int nkb_all = k_hi > kb0 * BK ? (k_hi - kb0 * BK + BK - 1) / BK : 0;
  int kb_begin = 0, kb_end = nkb_all;
  if (p.splitk > 1) {
    const int chunk = (nkb_all + p.splitk - 1) / p.splitk;
    kb_begin = split * chunk;
    kb_end = min(nkb_all, kb_begin + chunk);
  }
  ti.kb0 = kb0 + kb_begin;
  ti.nkb = max(0, kb_end - kb_begin);
  ti.zero = (ti.nkb == 0 && p.splitk == 1) ? 1 : 0;      // empty k-range of triangular operands: the product is 0
  return ti;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Persistent kernel: one CTA per SM, tiles handed out by an atomic counter.  The 512 TMEM columns hold two 128 x 256 fp32
// accumulators, so the epilogue of tile i overlaps the MMAs of tile i+1.
template <bool B_KMAJOR>
__global__ void __launch_bounds__(NTHREADS, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo,
                 const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapBlo, const Params p,
                 unsigned int* __restrict__ tile_counter, int total_tiles, int mtiles, int ntiles) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ uint32_t tmem_base_slot;
  __shared__ int sched_tile[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_full = empty_bar + STAGES;     // [2]
  uint64_t* accum_empty = accum_full + 2;        // [2]
  uint64_t* sched_full = accum_empty + 2;        // [2]
  uint64_t* sched_empty = sched_full + 2;        // [2]

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(accum_full + s, 1);
      mbar_init(accum_empty + s, 4);             // one arrival per epilogue warp
      mbar_init(sched_full + s, 1);
      mbar_init(sched_empty + s, 5);             // MMA thread + 4 epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t tx_bytes = p.n_terms == 3 ? (uint32_t)STAGE_BYTES : (uint32_t)(A_BYTES + B_BYTES);

  if (warp == 0) {
    // ===== scheduler + TMA producer (one thread) =====
    if (lane == 0) {
      uint32_t it = 0;                                     // running k-block counter across tiles (stage ring position)
      for (uint32_t iter = 0;; ++iter) {
        const int slot = iter & 1;
        mbar_wait(sched_empty + slot, ((iter >> 1) & 1u) ^ 1u);
        const int t = (int)atomicAdd(tile_counter, 1u);
        sched_tile[slot] = t < total_tiles ? t : -1;
        mbar_arrive(sched_full + slot);                    // release: the tile index is visible to the waiters
        if (t >= total_tiles) break;
        const TileInfo ti = tile_info(p, t, mtiles, ntiles);
        for (int kb = 0; kb < ti.nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(empty_bar + s, ph ^ 1u);
          unsigned char* st = smem + s * STAGE_BYTES;
          mbar_expect_tx(full_bar + s, tx_bytes);
          const int kc = (ti.kb0 + kb) * BK;
          tma_load_3d(&mapA, full_bar + s, st, kc, ti.i0, ti.b);
          if (B_KMAJOR) {
            tma_load_3d(&mapB, full_bar + s, st + 2 * A_BYTES, kc, ti.j0, ti.b);
          } else if (p.b_map4d) {                  // one box {32 cols, 16 k-rows, 8 column chunks}: chunk-major in smem
            tma_load_4d(&mapB, full_bar + s, st + 2 * A_BYTES, 0, kc, ti.j0 / 32, ti.b);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 32; ++c)        // 32-column chunks: 16 k-rows x 128 B each, 2 KB apart
              tma_load_3d(&mapB, full_bar + s, st + 2 * A_BYTES + c * 2048, ti.j0 + c * 32, kc, ti.b);
          }
          if (p.n_terms == 3) {
            tma_load_3d(&mapAlo, full_bar + s, st + A_BYTES, kc, ti.i0, ti.b);
            if (B_KMAJOR) {
              tma_load_3d(&mapBlo, full_bar + s, st + 2 * A_BYTES + B_BYTES, kc, ti.j0, ti.b);
            } else if (p.b_map4d) {
              tma_load_4d(&mapBlo, full_bar + s, st + 2 * A_BYTES + B_BYTES, 0, kc, ti.j0 / 32, ti.b);
            } else {
#pragma unroll
              for (int c = 0; c < BN / 32; ++c)
                tma_load_3d(&mapBlo, full_bar + s, st + 2 * A_BYTES + B_BYTES + c * 2048, ti.j0 + c * 32, kc, ti.b);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32 [10,13)=2,
      // a_major [15]=0 (K), b_major [16], N>>3 [17,23), M>>4 [24,29)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((B_KMAJOR ? 0u : 1u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      uint32_t it = 0, acc_iter = 0;
      for (uint32_t iter = 0;; ++iter) {
        const int slot = iter & 1;
        mbar_wait(sched_full + slot, (iter >> 1) & 1u);
        const int t = sched_tile[slot];
        mbar_arrive(sched_empty + slot);
        if (t < 0) break;
        const TileInfo ti = tile_info(p, t, mtiles, ntiles);
        if (ti.nkb == 0) continue;
        const uint32_t as = acc_iter & 1u;
        mbar_wait(accum_empty + as, ((acc_iter >> 1) & 1u) ^ 1u);       // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + as * (uint32_t)BN;
        for (int kb = 0; kb < ti.nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(full_bar + s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t sb = sa + 2 * A_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            // K-major SW64: rows of 64 B, 8-row groups 512 B apart; one K=8 step = 32 B inside the row
            const uint64_t a_raw = make_desc(sa + ks * 32, 16, 512, 4);
            const uint64_t a_lo = make_desc(sa + A_BYTES + ks * 32, 16, 512, 4);
            uint64_t b_raw, b_lo;
            if (B_KMAJOR) {
              b_raw = make_desc(sb + ks * 32, 16, 512, 4);
              b_lo = make_desc(sb + B_BYTES + ks * 32, 16, 512, 4);
            } else {
              // MN-major tf32 must use SWIZZLE_128B_BASE32B (layout type 1; 32 B chunks swizzled inside a 4-row x 128 B
              // atom): chunks of 32 columns (128 B rows) 2 KB apart (= LBO), 4-row k groups 512 B apart (= SBO);
              // one K=8 step = two k groups = 1 KB
              b_raw = make_desc(sb + ks * 1024, 2048, 512, 1);
              b_lo = make_desc(sb + B_BYTES + ks * 1024, 2048, 512, 1);
            }
            const uint32_t acc0 = (kb > 0 || ks > 0) ? 1u : 0u;
            if (p.n_terms == 3) {
              umma_tf32(tmem_d, a_raw, b_lo, idesc, acc0);
              umma_tf32(tmem_d, a_lo, b_raw, idesc, 1u);
              umma_tf32(tmem_d, a_raw, b_raw, idesc, 1u);
            } else {
              umma_tf32(tmem_d, a_raw, b_raw, idesc, acc0);
            }
          }
          umma_commit(empty_bar + s);               // frees the smem stage once these MMAs have read it
        }
        umma_commit(accum_full + as);               // accumulator complete
        ++acc_iter;
      }
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====
    const int q = warp & 3;
    const bool vec_ok = (p.ldd % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0) &&
                        (!p.Dlo || (reinterpret_cast<uintptr_t>(p.Dlo) & 15) == 0) &&
                        (!p.Cin || (reinterpret_cast<uintptr_t>(p.Cin) & 15) == 0) && ((p.sD % 4) == 0);
    float* epi = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256) + (warp - 2) * 32 * EPI_LD;
    uint32_t acc_iter = 0;
    for (uint32_t iter = 0;; ++iter) {
      const int slot = iter & 1;
      mbar_wait(sched_full + slot, (iter >> 1) & 1u);
      const int t = sched_tile[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(sched_empty + slot);
      if (t < 0) break;
      const TileInfo ti = tile_info(p, t, mtiles, ntiles);
      const bool has_acc = ti.nkb > 0;
      if (!has_acc && !ti.zero) continue;
      const uint32_t as = acc_iter & 1u;
      if (has_acc) {
        mbar_wait(accum_full + as, (acc_iter >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const int gi = ti.i0 + q * 32 + lane;
      const int wrow0 = ti.i0 + q * 32;
      float racc[8];                          // mode 3: per-row partial sums of this lane's 8 rows (coalesced path)
      float qreg[8];                          // rowv of this lane's 8 rows (coalesced path)
      float racc_s = 0.f;                     // mode 3, scalar path: row gi
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        racc[u] = 0.f;
        const int rr = wrow0 + u * 4 + (lane >> 3);
        qreg[u] = (p.epi_mode != 0 && p.epi_mode != 2 && rr < p.m) ? p.rowv[(int64_t)ti.b * p.m + rr] : 0.f;
      }
      const float q_s = (p.epi_mode != 0 && p.epi_mode != 2 && gi < p.m) ? p.rowv[(int64_t)ti.b * p.m + gi] : 0.f;
      float* Drow = p.D + (int64_t)ti.b * p.sD + (int64_t)gi * p.ldd;
      float* Lrow = p.Dlo ? p.Dlo + (int64_t)ti.b * p.sD + (int64_t)gi * p.ldd : nullptr;
      const float* Crow = p.Cin ? p.Cin + (int64_t)ti.b * p.sD + (int64_t)gi * p.ldd : nullptr;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        if (has_acc) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + as * (uint32_t)BN + (uint32_t)(c * 32), r);
        } else {
#pragma unroll
          for (int u = 0; u < 32; ++u) r[u] = 0u;
        }
        const int gj0 = ti.j0 + c * 32;
        // warp-uniform: the whole 32 x 32 block of this warp lies inside the matrix and inside the stored triangle
        const int wi0 = ti.i0 + q * 32;
        const bool warp_full = (wi0 + 32 <= p.m) && (gj0 + 32 <= p.n) &&
                               !((p.d_tri == 1 && gj0 + 31 > wi0) || (p.d_tri == 2 && gj0 < wi0 + 31));
        if (!warp_full && (gi >= p.m || gj0 >= p.n)) continue;
        if (p.splitk > 1) {
          for (int u = 0; u < 32; ++u) {
            const int gj = gj0 + u;
            if (gj >= p.n) break;
            if ((p.d_tri == 1 && gj > gi) || (p.d_tri == 2 && gj < gi)) continue;
            atomicAdd(Drow + gj, p.alpha * __uint_as_float(r[u]));
          }
        } else if (warp_full && vec_ok) {
          // transpose the 32 x 32 block through shared memory so that every store instruction writes whole 128-byte
          // lines: lane -> (row = 4*it + lane/8, 4 columns at (lane%8)*4)
#pragma unroll
          for (int u = 0; u < 32; u += 4)
            *reinterpret_cast<float4*>(epi + lane * EPI_LD + u) =
                make_float4(__uint_as_float(r[u]), __uint_as_float(r[u + 1]), __uint_as_float(r[u + 2]), __uint_as_float(r[u + 3]));
          __syncwarp();
          const int cq = (lane & 7) * 4;
          float4 cs1 = make_float4(0.f, 0.f, 0.f, 0.f), cs2 = make_float4(0.f, 0.f, 0.f, 0.f);
          float4 cv1 = cs1, cv2 = cs1;
          if (p.epi_mode == 3) {
            cv1 = *reinterpret_cast<const float4*>(p.colv1 + (int64_t)ti.b * p.n + gj0 + cq);
            cv2 = *reinterpret_cast<const float4*>(p.colv2 + (int64_t)ti.b * p.n + gj0 + cq);
          }
#pragma unroll
          for (int itr = 0; itr < 8; ++itr) {
            const int rr = itr * 4 + (lane >> 3);
            float4 v = *reinterpret_cast<const float4*>(epi + rr * EPI_LD + cq);
            v.x *= p.alpha; v.y *= p.alpha; v.z *= p.alpha; v.w *= p.alpha;
            const int64_t o = (int64_t)ti.b * p.sD + (int64_t)(ti.i0 + q * 32 + rr) * p.ldd + gj0 + cq;
            if (p.Cin) {
              const float4 cc = *reinterpret_cast<const float4*>(p.Cin + o);
              v.x += cc.x; v.y += cc.y; v.z += cc.z; v.w += cc.w;
            }
            if (p.epi_mode == 3) {
              const float4 ax = *reinterpret_cast<const float4*>(p.Aux + o);
              const float qv = qreg[itr];
              v.x += fmaf(qv, cv2.x, -2.f * ax.x * cv1.x);
              v.y += fmaf(qv, cv2.y, -2.f * ax.y * cv1.y);
              v.z += fmaf(qv, cv2.z, -2.f * ax.z * cv1.z);
              v.w += fmaf(qv, cv2.w, -2.f * ax.w * cv1.w);
              float rp = ax.x * cv2.x + ax.y * cv2.y + ax.z * cv2.z + ax.w * cv2.w;
              rp += __shfl_xor_sync(0xffffffffu, rp, 1);
              rp += __shfl_xor_sync(0xffffffffu, rp, 2);
              rp += __shfl_xor_sync(0xffffffffu, rp, 4);
              racc[itr] += rp;
            } else if (p.epi_mode != 0) {
              cs1.x = fmaf(v.x, v.x, cs1.x); cs1.y = fmaf(v.y, v.y, cs1.y); cs1.z = fmaf(v.z, v.z, cs1.z); cs1.w = fmaf(v.w, v.w, cs1.w);
              if (p.epi_mode == 1) {
                const float qv = qreg[itr];
                cs2.x = fmaf(qv, v.x, cs2.x); cs2.y = fmaf(qv, v.y, cs2.y); cs2.z = fmaf(qv, v.z, cs2.z); cs2.w = fmaf(qv, v.w, cs2.w);
              }
            }
            *reinterpret_cast<float4*>(p.D + o) = v;
            if (p.Dlo) {
              float4 lo;
              lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
              lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
              lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
              lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
              *reinterpret_cast<float4*>(p.Dlo + o) = lo;
            }
          }
          if (p.epi_mode == 1 || p.epi_mode == 2) {
            // column sums over the warp's 32 rows: combine the four row groups (lane/8), then lanes 0..7 own 4 columns each
#pragma unroll
            for (int sh = 8; sh <= 16; sh <<= 1) {
              cs1.x += __shfl_xor_sync(0xffffffffu, cs1.x, sh); cs1.y += __shfl_xor_sync(0xffffffffu, cs1.y, sh);
              cs1.z += __shfl_xor_sync(0xffffffffu, cs1.z, sh); cs1.w += __shfl_xor_sync(0xffffffffu, cs1.w, sh);
              if (p.epi_mode == 1) {
                cs2.x += __shfl_xor_sync(0xffffffffu, cs2.x, sh); cs2.y += __shfl_xor_sync(0xffffffffu, cs2.y, sh);
                cs2.z += __shfl_xor_sync(0xffffffffu, cs2.z, sh); cs2.w += __shfl_xor_sync(0xffffffffu, cs2.w, sh);
              }
            }
            if (lane < 8) {
              float* c1 = p.col1 + (int64_t)ti.b * p.n + gj0 + cq;
              atomicAdd(c1, cs1.x); atomicAdd(c1 + 1, cs1.y); atomicAdd(c1 + 2, cs1.z); atomicAdd(c1 + 3, cs1.w);
              if (p.epi_mode == 1) {
                float* c2 = p.col2 + (int64_t)ti.b * p.n + gj0 + cq;
                atomicAdd(c2, cs2.x); atomicAdd(c2 + 1, cs2.y); atomicAdd(c2 + 2, cs2.z); atomicAdd(c2 + 3, cs2.w);
              }
            }
          }
          __syncwarp();
        } else {
          for (int u = 0; u < 32; ++u) {
            const int gj = gj0 + u;
            if (gj >= p.n) break;
            if ((p.d_tri == 1 && gj > gi) || (p.d_tri == 2 && gj < gi)) continue;
            float v = p.alpha * __uint_as_float(r[u]);
            if (Crow) v += Crow[gj];
            if (p.epi_mode == 3) {
              const float ax = p.Aux[(int64_t)ti.b * p.sD + (int64_t)gi * p.ldd + gj];
              const float g2 = p.colv2[(int64_t)ti.b * p.n + gj];
              v += fmaf(q_s, g2, -2.f * ax * p.colv1[(int64_t)ti.b * p.n + gj]);
              racc_s = fmaf(ax, g2, racc_s);
            } else if (p.epi_mode != 0) {
              atomicAdd(p.col1 + (int64_t)ti.b * p.n + gj, v * v);
              if (p.epi_mode == 1) atomicAdd(p.col2 + (int64_t)ti.b * p.n + gj, q_s * v);
            }
            Drow[gj] = v;
            if (Lrow) Lrow[gj] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
          }
        }
      }
      if (p.epi_mode == 3) {
        if ((lane & 7) == 0) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int rr = wrow0 + u * 4 + (lane >> 3);
            if (rr < p.m && racc[u] != 0.f) atomicAdd(p.rowacc + (int64_t)ti.b * p.m + rr, racc[u]);
          }
        }
        if (gi < p.m && racc_s != 0.f) atomicAdd(p.rowacc + (int64_t)ti.b * p.m + gi, racc_s);
      }
      if (has_acc) {
        // this warp is done reading the accumulator stage: hand it back to the MMA thread
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(accum_empty + as);
        ++acc_iter;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// K-major operand: rows x k (k contiguous, row stride ld), batch stride sB  ->  3-D map {k, rows, batch}, box {16, box_rows, 1}
static int make_map_kmajor(CUtensorMap* map, const float* base, int rows, int k, int64_t ld, int64_t sB, int batch, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return GPZ_ERR_UNSUPPORTED;
  cuuint64_t gdim[3] = {(cuuint64_t)k, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(batch > 1 ? sB : (int64_t)rows * ld) * 4};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? GPZ_OK : GPZ_ERR_BADARG;
}

// MN-major operand: k x n (n contiguous, row stride ld)  ->  3-D map {n, k, batch}, box {32, 16, 1} (128 B rows, SWIZZLE_128B_ATOM_32B)
static int make_map_mnmajor(CUtensorMap* map, const float* base, int k, int n, int64_t ld, int64_t sB, int batch) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return GPZ_ERR_UNSUPPORTED;
  cuuint64_t gdim[3] = {(cuuint64_t)n, (cuuint64_t)k, (cuuint64_t)batch};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(batch > 1 ? sB : (int64_t)k * ld) * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)BK, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? GPZ_OK : GPZ_ERR_BADARG;
}

// MN-major operand with n % 32 == 0: 4-D view {32, k, n/32, batch} (strides 4 B, ld, 128 B, sB) so that ONE box {32, 16, 8, 1}
// lands chunk-major in shared memory (8 chunks x 16 rows x 128 B) - 1 TMA instruction per stage instead of 8.
static int make_map_mnmajor4d(CUtensorMap* map, const float* base, int k, int n, int64_t ld, int64_t sB, int batch) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return GPZ_ERR_UNSUPPORTED;
  cuuint64_t gdim[4] = {32, (cuuint64_t)k, (cuuint64_t)(n / 32), (cuuint64_t)batch};
  cuuint64_t gstr[3] = {(cuuint64_t)ld * 4, 128, (cuuint64_t)(batch > 1 ? sB : (int64_t)k * ld) * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)BK, (cuuint32_t)(BN / 32), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? GPZ_OK : GPZ_ERR_BADARG;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

constexpr int NUM_COUNTERS = 256;
__device__ unsigned int g_tile_counters[NUM_COUNTERS];

}  // namespace umma
}  // namespace gpz

using namespace gpz;
using namespace gpz::umma;

// 1 if the tcgen05 path can take this problem (alignment / divisibility), 0 otherwise (caller uses gpz_gemm_f32)
extern "C" int gpz_umma_gemm_supported(int b_kmajor, int m, int n, int k, int64_t lda, int64_t ldb, int64_t ldd) {
  if (m < 1 || n < 1 || k < 1) return 0;
  if (lda % 4 || ldb % 4) return 0;
  (void)ldd;
  return 1;
}

extern "C" int gpz_umma_gemm_f32(int b_kmajor, int m, int n, int k, float alpha, const float* A, const float* Alo, int64_t lda,
                                 int64_t sA, const float* B, const float* Blo, int64_t ldb, int64_t sB, const float* Cin, float* D,
                                 float* Dlo, int64_t ldd, int64_t sD, int batch, int a_tri, int b_tri, int d_tri, int splitk,
                                 int n_terms, void* stream) {
  return umma_gemm_ex(b_kmajor, m, n, k, alpha, A, Alo, lda, sA, B, Blo, ldb, sB, Cin, D, Dlo, ldd, sD, batch, a_tri, b_tri, d_tri,
                      splitk, n_terms, nullptr, stream);
}

int umma_gemm_ex(int b_kmajor, int m, int n, int k, float alpha, const float* A, const float* Alo, int64_t lda, int64_t sA,
                 const float* B, const float* Blo, int64_t ldb, int64_t sB, const float* Cin, float* D, float* Dlo, int64_t ldd,
                 int64_t sD, int batch, int a_tri, int b_tri, int d_tri, int splitk, int n_terms, const UmmaEpilogue* epi,
                 void* stream) {
  if (!gpz_umma_gemm_supported(b_kmajor, m, n, k, lda, ldb, ldd)) return GPZ_ERR_UNSUPPORTED;
  if (n_terms != 1 && n_terms != 3) return GPZ_ERR_BADARG;
  if (n_terms == 3 && (!Alo || !Blo)) return GPZ_ERR_BADARG;
  if (!aligned16(A) || !aligned16(B) || (Alo && !aligned16(Alo)) || (Blo && !aligned16(Blo)) || (sA % 4) || (sB % 4))
    return GPZ_ERR_UNSUPPORTED;
  if (splitk < 1) splitk = 1;
  if (splitk > 1 && (Cin || Dlo)) return GPZ_ERR_BADARG;
  CUtensorMap mA, mAlo, mB, mBlo;
  int b_map4d = 0;
  int rc = make_map_kmajor(&mA, A, m, k, lda, sA, batch, BM);
  if (rc) return rc;
  rc = make_map_kmajor(&mAlo, Alo ? Alo : A, m, k, lda, sA, batch, BM);
  if (rc) return rc;
  if (b_kmajor) {
    rc = make_map_kmajor(&mB, B, n, k, ldb, sB, batch, BN);
    if (rc) return rc;
    rc = make_map_kmajor(&mBlo, Blo ? Blo : B, n, k, ldb, sB, batch, BN);
  } else {
    static int use4d = -1;
    if (use4d < 0) { const char* e = getenv("GPZ_UMMA_MAP4D"); use4d = e ? atoi(e) : 1; }
    b_map4d = (use4d && n % 32 == 0) ? 1 : 0;
    if (b_map4d) {
      rc = make_map_mnmajor4d(&mB, B, k, n, ldb, sB, batch);
      if (rc == GPZ_OK) rc = make_map_mnmajor4d(&mBlo, Blo ? Blo : B, k, n, ldb, sB, batch);
      if (rc != GPZ_OK) b_map4d = 0;          // driver refused the 4-D view: fall back to 8 plain boxes
    }
    if (!b_map4d) {
      rc = make_map_mnmajor(&mB, B, k, n, ldb, sB, batch);
      if (rc) return rc;
      rc = make_map_mnmajor(&mBlo, Blo ? Blo : B, k, n, ldb, sB, batch);
    }
  }
  if (rc) return rc;
  Params p;
  p.b_map4d = b_map4d;
  p.epi_mode = epi ? epi->mode : 0;
  p.Aux = epi ? epi->Aux : nullptr; p.rowv = epi ? epi->rowv : nullptr; p.colv1 = epi ? epi->colv1 : nullptr;
  p.colv2 = epi ? epi->colv2 : nullptr; p.col1 = epi ? epi->col1 : nullptr; p.col2 = epi ? epi->col2 : nullptr;
  p.rowacc = epi ? epi->rowacc : nullptr;
  if (p.epi_mode != 0 && (splitk > 1 || d_tri != 0)) return GPZ_ERR_BADARG;
  p.D = D; p.Dlo = Dlo; p.Cin = Cin; p.m = m; p.n = n; p.k = k; p.ldd = ldd; p.sD = sD; p.batch = batch; p.splitk = splitk;
  p.a_tri = a_tri; p.b_tri = b_tri; p.d_tri = d_tri; p.n_terms = n_terms; p.alpha = alpha;
  const int mtiles = (int)cdiv(m, BM), ntiles = (int)cdiv(n, BN);
  const int64_t total64 = (int64_t)mtiles * ntiles * batch * splitk;
  if (total64 > 0x7fffffff) return GPZ_ERR_UNSUPPORTED;
  const int total = (int)total64;
  cudaStream_t st = (cudaStream_t)stream;
  // dynamic tile scheduler: one counter per in-flight launch, taken round-robin from a small device array
  static unsigned int* counters = nullptr;
  static std::atomic<unsigned int> next_slot{0};
  static int num_sms = 0;
  if (!counters) {
    unsigned int* ptr = nullptr;
    GPZ_CUDA(cudaGetSymbolAddress((void**)&ptr, g_tile_counters));
    counters = ptr;
    int dev = 0;
    GPZ_CUDA(cudaGetDevice(&dev));
    GPZ_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  unsigned int* counter = counters + (next_slot.fetch_add(1) % NUM_COUNTERS);
  GPZ_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
  const int grid = total < num_sms ? total : num_sms;
  if (b_kmajor) {
    GPZ_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    umma_gemm_kernel<true><<<grid, NTHREADS, SMEM_BYTES, st>>>(mA, mAlo, mB, mBlo, p, counter, total, mtiles, ntiles);
  } else {
    GPZ_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    umma_gemm_kernel<false><<<grid, NTHREADS, SMEM_BYTES, st>>>(mA, mAlo, mB, mBlo, p, counter, total, mtiles, ntiles);
  }
  GPZ_CHECK_LAUNCH();
  return GPZ_OK;
}

// lo = x - tf32_trunc(x) for a flat array, and (optionally) the batched transpose of an M x M matrix with its lo part
__global__ void tf32_lo_kernel(const float* __restrict__ x, float* __restrict__ lo, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    float4 o;
    o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    *reinterpret_cast<float4*>(lo + i) = o;
  } else {
    for (int64_t j = i; j < n; ++j) lo[j] = x[j] - __uint_as_float(__float_as_uint(x[j]) & 0xFFFFE000u);
  }
}
__global__ void transpose_lo_kernel(const float* __restrict__ x, float* __restrict__ xt, float* __restrict__ xt_lo, int M) {
  __shared__ float tile[32][33];
  const float* X = x + (int64_t)blockIdx.z * M * M;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = by + r, j = bx + threadIdx.x;
    tile[r][threadIdx.x] = (i < M && j < M) ? X[(int64_t)i * M + j] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = bx + r, j = by + threadIdx.x;          // output (i, j) = input (j, i)
    if (i < M && j < M) {
      const float v = tile[threadIdx.x][r];
      const int64_t o = (int64_t)blockIdx.z * M * M + (int64_t)i * M + j;
      xt[o] = v;
      if (xt_lo) xt_lo[o] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    }
  }
}

extern "C" int gpz_tf32_lo_f32(const float* x, float* lo, int64_t n, void* stream) {
  if (n <= 0) return GPZ_OK;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(lo) & 15)) return GPZ_ERR_UNSUPPORTED;
  tf32_lo_kernel<<<(unsigned)cdiv(cdiv(n, 4), 256), 256, 0, (cudaStream_t)stream>>>(x, lo, n);
  GPZ_CHECK_LAUNCH();
  return GPZ_OK;
}
extern "C" int gpz_transpose_lo_f32(const float* x, float* xt, float* xt_lo, int M, int L, void* stream) {
  dim3 grid((unsigned)cdiv(M, 32), (unsigned)cdiv(M, 32), L);
  transpose_lo_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, xt, xt_lo, M);
  GPZ_CHECK_LAUNCH();
  return GPZ_OK;
}
