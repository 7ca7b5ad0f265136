// Internal (non-ABI) entry of the tcgen05 GEMM with the fused epilogues used by csrc/predict.cu.
#pragma once
#include <stdint.h>

struct UmmaEpilogue {
  int mode;              // see Params::epi_mode in umma_gemm.cu
  const float* Aux; const float* rowv; const float* colv1; const float* colv2;
  float* col1; float* col2; float* rowacc;
};

int umma_gemm_ex(int b_kmajor, int m, int n, int k, float alpha, const float* A, const float* Alo, int64_t lda, int64_t sA,
                 const float* B, const float* Blo, int64_t ldb, int64_t sB, const float* Cin, float* D, float* Dlo, int64_t ldd,
                 int64_t sD, int batch, int a_tri, int b_tri, int d_tri, int splitk, int n_terms, const UmmaEpilogue* epi,
                 void* stream);
