// Python bindings for the tcgen05 flash-attention kernels.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <cstdlib>

#include "api.h"
#include "comm_api.h"

namespace dtg {
namespace {
using torch::Tensor;

void check_qkv(const Tensor& qkv, int64_t nh, int64_t nkv) {
  TORCH_CHECK(qkv.is_cuda() && qkv.is_contiguous() && qkv.scalar_type() == at::kBFloat16,
              "qkv must be a contiguous bf16 CUDA tensor");
  TORCH_CHECK(qkv.dim() == 4 && qkv.size(2) == nh + 2 * nkv && qkv.size(3) == 128,
              "qkv must be [B, S, nh+2*nkv, 128]");
}

// version: 0 = default (DTG_ATTN_FWD env, else 2), 1 = one query tile per CTA (attention_fwd.cu),
// 2 = two tiles per CTA, P kept in tensor memory (attention_fwd2.cu)
int default_fwd_version() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DTG_ATTN_FWD");
    v = e ? atoi(e) : 2;
    if (v != 1 && v != 2) v = 2;
  }
  return v;
}

std::tuple<Tensor, Tensor> py_attn_fwd(const Tensor& qkv, int64_t nh, int64_t nkv, double scale, int64_t version) {
  check_qkv(qkv, nh, nkv);
  const c10::cuda::CUDAGuard guard(qkv.device());
  const int64_t B = qkv.size(0), S = qkv.size(1);
  Tensor o = torch::empty({B, S, nh, 128}, qkv.options());
  Tensor lse = torch::empty({B, nh, S}, qkv.options().dtype(at::kFloat));
  if (version == 0) version = default_fwd_version();
  auto fn = version == 1 ? dtg::attn_fwd : dtg::attn_fwd2;
  fn(qkv.data_ptr(), o.data_ptr(), lse.data_ptr<float>(), (int)B, (int)S, (int)nh, (int)nkv, (float)scale,
     at::cuda::getCurrentCUDAStream().stream());
  return {o, lse};
}

Tensor py_attn_bwd(const Tensor& d_o, const Tensor& qkv, const Tensor& o, const Tensor& lse, int64_t nh, int64_t nkv,
                   double scale, const c10::optional<Tensor>& trace, int64_t mode) {
  check_qkv(qkv, nh, nkv);
  TORCH_CHECK(d_o.is_contiguous() && o.is_contiguous() && d_o.scalar_type() == at::kBFloat16, "dO/O must be contiguous bf16");
  const c10::cuda::CUDAGuard guard(qkv.device());
  const int64_t B = qkv.size(0), S = qkv.size(1);
  Tensor dqkv = torch::empty_like(qkv);
  Tensor delta = torch::empty({B, nh, S}, qkv.options().dtype(at::kFloat));
  float* tr = nullptr;
  if (trace.has_value()) {
    TORCH_CHECK(trace->scalar_type() == at::kLong && trace->numel() >= 1024, "trace must be int64[1024]");
    tr = reinterpret_cast<float*>(trace->data_ptr<int64_t>());
  }
  dtg::attn_bwd(qkv.data_ptr(), o.data_ptr(), d_o.data_ptr(), lse.data_ptr<float>(), delta.data_ptr<float>(), tr,
                dqkv.data_ptr(), (int)B, (int)S, (int)nh, (int)nkv, (float)scale, (int)mode,
                at::cuda::getCurrentCUDAStream().stream());
  return dqkv;
}
}  // namespace

void bind_attention(pybind11::module_& m) {
  m.def("attn_fwd", &py_attn_fwd, pybind11::arg("qkv"), pybind11::arg("nh"), pybind11::arg("nkv"), pybind11::arg("scale"),
        pybind11::arg("version") = 0);
  m.def("attn_bwd", &py_attn_bwd, pybind11::arg("d_o"), pybind11::arg("qkv"), pybind11::arg("o"), pybind11::arg("lse"),
        pybind11::arg("nh"), pybind11::arg("nkv"), pybind11::arg("scale"), pybind11::arg("trace") = pybind11::none(),
        pybind11::arg("mode") = 0);   // 0 = default (DTG_ATTN_BWD), 1 = P/dS through shared memory, 2 = P/dS in TMEM
}
}  // namespace dtg
