// Training-step update: one multi-tensor Adam launch over all parameters of the model, with the
// reference's post-step clamp of the loadings (utilities.py:621-623  optimizer.step(); model.W.clamp_(min=0)) fused in.
//
//   m = b1 m + (1 - b1) g ;  v = b2 v + (1 - b2) g^2 ;  p -= (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
//
// (torch.optim.Adam without weight decay / amsgrad, same operation order).  HBM-bound: 16 bytes read + 12 written per element.
#include <math.h>

#include "common.cuh"
#include "gpzoo_b200.h"

namespace gpz {

constexpr int ADAM_MAX_TENSORS = 40;
constexpr int ADAM_CHUNK = 1024;         // elements per CTA; a chunk never straddles two tensors

template <typename T> struct AdamTable {
  T* p[ADAM_MAX_TENSORS]; const T* g[ADAM_MAX_TENSORS]; T* m[ADAM_MAX_TENSORS]; T* v[ADAM_MAX_TENSORS];
  int64_t numel[ADAM_MAX_TENSORS];
  int chunk0[ADAM_MAX_TENSORS + 1];      // first chunk of every tensor (prefix sums)
  unsigned char clamp0[ADAM_MAX_TENSORS];
  int n;
};

template <typename T>
__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamTable<T> tab, T step_size, T b1, T b2, T eps,
                                                   T inv_sqrt_bc2) {
  const int chunk = blockIdx.x;
  int t = 0;
  while (t + 1 < tab.n && chunk >= tab.chunk0[t + 1]) ++t;
  const int64_t base = (int64_t)(chunk - tab.chunk0[t]) * ADAM_CHUNK;
  T* __restrict__ p = tab.p[t]; const T* __restrict__ g = tab.g[t]; T* __restrict__ m = tab.m[t]; T* __restrict__ v = tab.v[t];
  const int64_t n = tab.numel[t];
  const bool clamp = tab.clamp0[t] != 0;
#pragma unroll
  for (int u = 0; u < ADAM_CHUNK / 256; ++u) {
    const int64_t i = base + u * 256 + threadIdx.x;
    if (i < n) {
      const T gi = g[i];
      const T mi = b1 * m[i] + (T(1) - b1) * gi;          // torch: exp_avg.lerp_(grad, 1 - beta1)
      const T vi = b2 * v[i] + (T(1) - b2) * gi * gi;     //        exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      m[i] = mi;
      v[i] = vi;
      T pi = p[i] - step_size * (mi / (Num<T>::sqrt(vi) * inv_sqrt_bc2 + eps));
      if (clamp) pi = pi < T(0) ? T(0) : pi;
      p[i] = pi;
    }
  }
}

template <typename T>
int adam_step(int n_tensors, void* const* params, void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
              const int64_t* numel, const int* clamp0, double lr, double beta1, double beta2, double eps, int step, cudaStream_t st) {
  if (n_tensors < 0 || step < 1) return GPZ_ERR_BADARG;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  for (int t0 = 0; t0 < n_tensors; t0 += ADAM_MAX_TENSORS) {
    AdamTable<T> tab;
    tab.n = n_tensors - t0 < ADAM_MAX_TENSORS ? n_tensors - t0 : ADAM_MAX_TENSORS;
    int chunks = 0;
    for (int t = 0; t < tab.n; ++t) {
      tab.p[t] = (T*)params[t0 + t]; tab.g[t] = (const T*)grads[t0 + t]; tab.m[t] = (T*)exp_avg[t0 + t];
      tab.v[t] = (T*)exp_avg_sq[t0 + t];
      tab.numel[t] = numel[t0 + t];
      tab.clamp0[t] = clamp0 && clamp0[t0 + t] ? 1 : 0;
      tab.chunk0[t] = chunks;
      const int64_t c = cdiv(numel[t0 + t], ADAM_CHUNK);
      if (c + chunks > 0x7fffffff) return GPZ_ERR_UNSUPPORTED;
      chunks += (int)c;
    }
    tab.chunk0[tab.n] = chunks;
    if (chunks == 0) continue;
    adam_kernel<T><<<chunks, 256, 0, st>>>(tab, (T)(lr / bc1), (T)beta1, (T)beta2, (T)eps, (T)(1.0 / sqrt(bc2)));
    GPZ_CHECK_LAUNCH();
  }
  return GPZ_OK;
}

}  // namespace gpz

using namespace gpz;

extern "C" int gpz_adam_step_f32(int n_tensors, void* const* params, void* const* grads, void* const* exp_avg,
                                 void* const* exp_avg_sq, const int64_t* numel, const int* clamp0, double lr, double beta1,
                                 double beta2, double eps, int step, void* stream) {
  return adam_step<float>(n_tensors, params, grads, exp_avg, exp_avg_sq, numel, clamp0, lr, beta1, beta2, eps, step,
                          (cudaStream_t)stream);
}
extern "C" int gpz_adam_step_f64(int n_tensors, void* const* params, void* const* grads, void* const* exp_avg,
                                 void* const* exp_avg_sq, const int64_t* numel, const int* clamp0, double lr, double beta1,
                                 double beta2, double eps, int step, void* stream) {
  return adam_step<double>(n_tensors, params, grads, exp_avg, exp_avg_sq, numel, clamp0, lr, beta1, beta2, eps, step,
                           (cudaStream_t)stream);
}
