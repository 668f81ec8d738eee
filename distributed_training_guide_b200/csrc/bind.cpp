// Python bindings: torch tensors -> raw-pointer launchers (api.h).  Compiled by g++ only.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "api.h"
#include "comm_api.h"

namespace {

using torch::Tensor;

inline cudaStream_t stream() { return at::cuda::getCurrentCUDAStream().stream(); }

void check_bf16_2d(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.scalar_type() == at::kBFloat16, name, " must be bfloat16");
  TORCH_CHECK(t.dim() == 2, name, " must be 2-D");
  TORCH_CHECK(t.stride(1) == 1, name, " must have a contiguous last dimension");
}
void check_contig(const Tensor& t, const char* name, at::ScalarType dt) {
  TORCH_CHECK(t.is_cuda() && t.is_contiguous(), name, " must be a contiguous CUDA tensor");
  TORCH_CHECK(t.scalar_type() == dt, name, " has the wrong dtype");
}

void gemm(const Tensor& a, const Tensor& b, Tensor& out, bool trans_a, bool trans_b, bool accumulate, int variant) {
  check_bf16_2d(a, "a");
  check_bf16_2d(b, "b");
  check_bf16_2d(out, "out");
  const c10::cuda::CUDAGuard guard(a.device());
  const int M = (int)(trans_a ? a.size(1) : a.size(0));
  const int K = (int)(trans_a ? a.size(0) : a.size(1));
  const int N = (int)(trans_b ? b.size(0) : b.size(1));
  const int Kb = (int)(trans_b ? b.size(1) : b.size(0));
  TORCH_CHECK(K == Kb, "gemm: inner dimensions differ (", K, " vs ", Kb, ")");
  TORCH_CHECK(out.size(0) == M && out.size(1) == N, "gemm: out has the wrong shape");
  dtg::gemm_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, a.stride(0), b.stride(0), out.stride(0),
                 /*a_kmajor=*/!trans_a, /*b_kmajor=*/trans_b, accumulate, variant, stream());
}

std::tuple<Tensor, Tensor, c10::optional<Tensor>> rmsnorm_fwd(const Tensor& x, const Tensor& w, double eps,
                                                              const c10::optional<Tensor>& res) {
  check_contig(x, "x", at::kBFloat16);
  check_contig(w, "w", at::kBFloat16);
  const c10::cuda::CUDAGuard guard(x.device());
  const int T = (int)x.size(0), H = (int)x.size(1);
  Tensor y = torch::empty_like(x);
  Tensor rstd = torch::empty({T}, x.options().dtype(at::kFloat));
  c10::optional<Tensor> h;
  const void* rp = nullptr;
  void* hp = nullptr;
  if (res.has_value()) {
    check_contig(*res, "residual", at::kBFloat16);
    h = torch::empty_like(x);
    rp = res->data_ptr();
    hp = h->data_ptr();
  }
  dtg::rmsnorm_fwd(x.data_ptr(), rp, w.data_ptr(), y.data_ptr(), hp, rstd.data_ptr<float>(), T, H, (float)eps,
                   stream());
  return {y, rstd, h};
}

std::tuple<Tensor, Tensor> rmsnorm_bwd(const Tensor& dy, const Tensor& h, const Tensor& w, const Tensor& rstd,
                                       const c10::optional<Tensor>& dres) {
  check_contig(dy, "dy", at::kBFloat16);
  check_contig(h, "h", at::kBFloat16);
  const c10::cuda::CUDAGuard guard(dy.device());
  const int T = (int)dy.size(0), H = (int)dy.size(1);
  Tensor dx = torch::empty_like(dy);
  Tensor dw = torch::empty({H}, dy.options().dtype(at::kFloat));
  Tensor partial = torch::empty({dtg::rmsnorm_bwd_grid(T), H}, dy.options().dtype(at::kFloat));
  const void* dr = nullptr;
  if (dres.has_value()) {
    check_contig(*dres, "dres", at::kBFloat16);
    dr = dres->data_ptr();
  }
  dtg::rmsnorm_bwd(dy.data_ptr(), h.data_ptr(), w.data_ptr(), rstd.data_ptr<float>(), dr, dx.data_ptr(),
                   partial.data_ptr<float>(), dw.data_ptr<float>(), T, H, stream());
  return {dx, dw};
}

void rope_inplace(Tensor& qkv, const Tensor& cos, const Tensor& sin, int64_t n_rot, bool inverse) {
  // qkv: [B, S, heads, d] contiguous; cos/sin fp32 [S, d/2] or [B, S, d/2]
  check_contig(qkv, "qkv", at::kBFloat16);
  check_contig(cos, "cos", at::kFloat);
  check_contig(sin, "sin", at::kFloat);
  TORCH_CHECK(qkv.dim() == 4, "qkv must be [B,S,heads,d]");
  const c10::cuda::CUDAGuard guard(qkv.device());
  const int64_t B = qkv.size(0), S = qkv.size(1), NH = qkv.size(2), d = qkv.size(3);
  const bool per_token = cos.dim() == 3;
  TORCH_CHECK(cos.size(-1) == d / 2 && cos.size(per_token ? 1 : 0) == S, "cos/sin table has the wrong shape");
  dtg::rope_inplace(qkv.data_ptr(), cos.data_ptr<float>(), sin.data_ptr<float>(), B * S, (int)S, (int)NH, (int)n_rot,
                    (int)d, per_token, inverse, stream());
}

Tensor swiglu_fwd(const Tensor& gu) {
  check_contig(gu, "gu", at::kBFloat16);
  const c10::cuda::CUDAGuard guard(gu.device());
  const int64_t T = gu.size(0), I = gu.size(1) / 2;
  Tensor h = torch::empty({T, I}, gu.options());
  dtg::swiglu_fwd(gu.data_ptr(), h.data_ptr(), T, (int)I, stream());
  return h;
}
Tensor swiglu_bwd(const Tensor& dh, const Tensor& gu) {
  check_contig(dh, "dh", at::kBFloat16);
  check_contig(gu, "gu", at::kBFloat16);
  const c10::cuda::CUDAGuard guard(gu.device());
  Tensor dgu = torch::empty_like(gu);
  dtg::swiglu_bwd(dh.data_ptr(), gu.data_ptr(), dgu.data_ptr(), gu.size(0), (int)(gu.size(1) / 2), stream());
  return dgu;
}

Tensor cross_entropy_fwd_bwd(Tensor& logits, const Tensor& targets) {
  check_contig(logits, "logits", at::kBFloat16);
  check_contig(targets, "targets", at::kLong);
  const c10::cuda::CUDAGuard guard(logits.device());
  const int T = (int)logits.size(0), V = (int)logits.size(1);
  TORCH_CHECK(targets.numel() == T, "targets must have one entry per logits row");
  Tensor scratch = torch::empty({T + 2}, logits.options().dtype(at::kFloat));
  float* sp = scratch.data_ptr<float>();
  dtg::cross_entropy_fwd_bwd(logits.data_ptr(), (const long long*)targets.data_ptr<int64_t>(),
                             sp + 2, sp, sp + 1, T, V, stream());
  return scratch.slice(0, 1, 2).reshape({});
}

void scale_inplace(Tensor& x, const Tensor& scale) {
  check_contig(x, "x", at::kBFloat16);
  check_contig(scale, "scale", at::kFloat);
  const c10::cuda::CUDAGuard guard(x.device());
  dtg::scale_inplace(x.data_ptr(), scale.data_ptr<float>(), x.numel(), stream());
}

Tensor embedding_fwd(const Tensor& ids, const Tensor& w) {
  check_contig(ids, "ids", at::kLong);
  check_contig(w, "w", at::kBFloat16);
  const c10::cuda::CUDAGuard guard(w.device());
  Tensor out = torch::empty({ids.numel(), w.size(1)}, w.options());
  dtg::embedding_fwd((const long long*)ids.data_ptr<int64_t>(), w.data_ptr(), out.data_ptr(), ids.numel(),
                     (int)w.size(1), stream());
  return out;
}
void embedding_bwd_sorted(const Tensor& dout, const Tensor& ids_sorted, const Tensor& perm, Tensor& dw, bool accumulate) {
  TORCH_CHECK(dout.is_contiguous() && dw.is_contiguous() && ids_sorted.is_contiguous() && perm.is_contiguous(), "contiguous");
  TORCH_CHECK(dout.scalar_type() == at::kBFloat16 && dw.scalar_type() == at::kBFloat16, "bf16 gradients");
  TORCH_CHECK(ids_sorted.scalar_type() == at::kLong && perm.scalar_type() == at::kLong && perm.numel() == ids_sorted.numel(),
              "int64 ids / permutation");
  const c10::cuda::CUDAGuard guard(dw.device());
  dtg::embedding_bwd_sorted(dout.data_ptr(), (const long long*)ids_sorted.data_ptr<int64_t>(),
                            (const long long*)perm.data_ptr<int64_t>(), dw.data_ptr(), ids_sorted.numel(),
                            (int)dw.size(1), accumulate, at::cuda::getCurrentCUDAStream().stream());
}

void embedding_bwd(const Tensor& dout, const Tensor& ids, Tensor& dw) {
  check_contig(dout, "dout", at::kBFloat16);
  check_contig(ids, "ids", at::kLong);
  check_contig(dw, "dw", at::kBFloat16);
  const c10::cuda::CUDAGuard guard(dw.device());
  dtg::embedding_bwd(dout.data_ptr(), (const long long*)ids.data_ptr<int64_t>(), dw.data_ptr(), ids.numel(),
                     (int)dw.size(1), stream());
}

void adamw_flat(Tensor& p, const Tensor& g, Tensor& m, Tensor& v, double lr, double b1, double b2, double eps,
                double wd, int64_t step, double grad_scale) {
  check_contig(p, "p", at::kBFloat16);
  check_contig(g, "g", at::kBFloat16);
  TORCH_CHECK(m.scalar_type() == v.scalar_type(), "exp_avg / exp_avg_sq dtypes differ");
  const bool fp32 = m.scalar_type() == at::kFloat;
  TORCH_CHECK(fp32 || m.scalar_type() == at::kBFloat16, "optimizer state must be bf16 or fp32");
  TORCH_CHECK(p.numel() == g.numel() && p.numel() == m.numel() && p.numel() == v.numel(), "size mismatch");
  const c10::cuda::CUDAGuard guard(p.device());
  dtg::adamw_flat(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), (float)lr, (float)b1, (float)b2,
                  (float)eps, (float)wd, (int)step, (float)grad_scale, fp32, stream());
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "distributed_training_guide_b200 sm_100a kernels";
  m.def("launch_count", []() { return (uint64_t)dtg::launch_count(); });
  m.def("gemm", &gemm, py::arg("a"), py::arg("b"), py::arg("out"), py::arg("trans_a") = false,
        py::arg("trans_b") = false, py::arg("accumulate") = false, py::arg("variant") = 0);
  m.def("gemm_max_active_clusters", [](int cg) { return dtg::gemm_max_active_clusters(cg); });
  m.def("rmsnorm_fwd", &rmsnorm_fwd);
  m.def("rmsnorm_bwd", &rmsnorm_bwd);
  m.def("rope_inplace", &rope_inplace);
  m.def("swiglu_fwd", &swiglu_fwd);
  m.def("swiglu_bwd", &swiglu_bwd);
  m.def("cross_entropy_fwd_bwd", &cross_entropy_fwd_bwd);
  m.def("scale_inplace", &scale_inplace);
  m.def("embedding_fwd", &embedding_fwd);
  m.def("embedding_bwd", &embedding_bwd);
  m.def("embedding_bwd_sorted", &embedding_bwd_sorted);
  m.def("adamw_flat", &adamw_flat);
  dtg::bind_comm(m);
  dtg::bind_attention(m);
  dtg::bind_tp(m);
  dtg::bind_dataloader(m);
  dtg::bind_symm_vmm(m);
}
