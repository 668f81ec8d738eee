// AdamW over a flat parameter range: ONE launch per flat group (reference: ATen's
// multi-tensor `_fused_adamw_`, reached from `torch.optim.AdamW(..., fused=True)` at
// 01-single-gpu/train_llm.py:73).  bf16 parameters/gradients, fp32 math, states in bf16 (the
// reference's behaviour: states inherit the bf16 parameter dtype) or fp32.
// The per-element update `adamw_update` is shared with the fused NVLink kernels in comm.cu.
#include "adamw.cuh"
#include "api.h"
#include "common.cuh"

namespace dtg {

template <typename StateT>
__global__ void __launch_bounds__(256) adamw_flat_kernel(__nv_bfloat16* __restrict__ p,
                                                        const __nv_bfloat16* __restrict__ g, StateT* __restrict__ m,
                                                        StateT* __restrict__ v, long long nvec, AdamWHyper hp) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float fp[8], fg[8], fm[8], fv[8];
    unpack8(ld8(p + i * 8), fp);
    unpack8(ld8(g + i * 8), fg);
    load_state8(m + i * 8, fm);
    load_state8(v + i * 8, fv);
#pragma unroll
    for (int j = 0; j < 8; ++j) adamw_update(fp[j], fg[j], fm[j], fv[j], hp);
    st8(p + i * 8, pack8(fp));
    store_state8(m + i * 8, fm);
    store_state8(v + i * 8, fv);
  }
}

void adamw_flat(void* p, const void* g, void* m, void* v, long long n, float lr, float beta1, float beta2, float eps,
                float wd, int step, float grad_scale, bool state_fp32, cudaStream_t s) {
  if (n % 8 != 0) throw std::runtime_error("adamw_flat: range must be a multiple of 8 elements");
  AdamWHyper hp = make_adamw_hyper(lr, beta1, beta2, eps, wd, step, grad_scale);
  const long long nvec = n / 8;
  long long grid = (nvec + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  if (state_fp32)
    adamw_flat_kernel<float><<<(int)grid, 256, 0, s>>>((__nv_bfloat16*)p, (const __nv_bfloat16*)g, (float*)m,
                                                      (float*)v, nvec, hp);
  else
    adamw_flat_kernel<__nv_bfloat16><<<(int)grid, 256, 0, s>>>((__nv_bfloat16*)p, (const __nv_bfloat16*)g,
                                                              (__nv_bfloat16*)m, (__nv_bfloat16*)v, nvec, hp);
  note_launch();
  DTG_LAUNCH_CHECK();
}

}  // namespace dtg
