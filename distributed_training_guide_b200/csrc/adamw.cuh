// Per-element AdamW update shared by adamw.cu and the fused reduce-scatter+AdamW kernels.
#pragma once
#include <cmath>
#include "common.cuh"

namespace dtg {

struct AdamWHyper {
  float lr_wd;       // 1 - lr * weight_decay
  float beta1, beta2;
  float step_size;   // lr / (1 - beta1^t)
  float inv_sqrt_bc2;  // 1 / sqrt(1 - beta2^t)
  float eps;
  float grad_scale;
};

inline AdamWHyper make_adamw_hyper(float lr, float beta1, float beta2, float eps, float wd, int step, float gs) {
  AdamWHyper h;
  h.lr_wd = 1.f - lr * wd;
  h.beta1 = beta1;
  h.beta2 = beta2;
  const double bc1 = 1.0 - std::pow((double)beta1, (double)step);
  const double bc2 = 1.0 - std::pow((double)beta2, (double)step);
  h.step_size = (float)((double)lr / bc1);
  h.inv_sqrt_bc2 = (float)(1.0 / std::sqrt(bc2));
  h.eps = eps;
  h.grad_scale = gs;
  return h;
}

#ifdef __CUDACC__
// torch.optim.AdamW semantics: decoupled decay, bias-corrected moments.
__device__ __forceinline__ void adamw_update(float& p, float g, float& m, float& v, const AdamWHyper& h) {
  g *= h.grad_scale;
  p *= h.lr_wd;
  m = h.beta1 * m + (1.f - h.beta1) * g;
  v = h.beta2 * v + (1.f - h.beta2) * g * g;
  const float denom = sqrtf(v) * h.inv_sqrt_bc2 + h.eps;
  p -= h.step_size * (m / denom);
}

__device__ __forceinline__ void load_state8(const __nv_bfloat16* s, float (&f)[8]) { unpack8(ld8(s), f); }
__device__ __forceinline__ void store_state8(__nv_bfloat16* s, const float (&f)[8]) { st8(s, pack8(f)); }
__device__ __forceinline__ void load_state8(const float* s, float (&f)[8]) {
  const float4 a = reinterpret_cast<const float4*>(s)[0], b = reinterpret_cast<const float4*>(s)[1];
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store_state8(float* s, const float (&f)[8]) {
  reinterpret_cast<float4*>(s)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(s)[1] = make_float4(f[4], f[5], f[6], f[7]);
}

#endif  // __CUDACC__

}  // namespace dtg
