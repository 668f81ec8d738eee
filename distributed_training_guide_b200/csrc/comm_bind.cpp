// Symmetric-memory runtime (allocation outside the caching allocator, CUDA IPC export/import of
// peer mappings) and Python bindings of the NVLink collective kernels.
//
// Setup protocol (parallel/symm.py): every rank cudaMalloc's the same number of bytes, exports a
// 64-byte cudaIpcMemHandle, the handles are all-gathered through the torch.distributed store/NCCL
// bootstrap group, and each rank opens its peers' handles -> a table of N base pointers to the
// "same" buffer.  After that no library is involved: kernels address peers directly over NVLink.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <pybind11/stl.h>
#include <torch/extension.h>

#include <cstring>
#include <mutex>
#include <unordered_map>

#include "comm.cuh"
#include "comm_api.h"
#include "common.cuh"

namespace dtg {
namespace {
using torch::Tensor;

std::mutex g_mu;
std::unordered_map<uint64_t, size_t> g_local;  // ptr -> bytes (owned allocations)
size_t g_local_bytes = 0;

inline cudaStream_t stream() { return at::cuda::getCurrentCUDAStream().stream(); }

// Allocate `nbytes` of device memory for peer mapping; returns (uint8 tensor view, ipc handle bytes).
std::tuple<Tensor, py::bytes> symm_alloc(int64_t nbytes, int64_t device) {
  const c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
  void* p = nullptr;
  const size_t bytes = ((size_t)nbytes + 511) & ~(size_t)511;
  DTG_CUDA_CHECK(cudaMalloc(&p, bytes));
  DTG_CUDA_CHECK(cudaMemset(p, 0, bytes));
  DTG_CUDA_CHECK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  DTG_CUDA_CHECK(cudaIpcGetMemHandle(&h, p));
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_local[(uint64_t)p] = bytes;
    g_local_bytes += bytes;
  }
  auto deleter = [](void* q) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_local.find((uint64_t)q);
    if (it != g_local.end()) {
      g_local_bytes -= it->second;
      g_local.erase(it);
    }
    cudaFree(q);
  };
  Tensor t = torch::from_blob(p, {(int64_t)bytes}, deleter,
                              torch::TensorOptions().dtype(torch::kUInt8).device(torch::kCUDA, (c10::DeviceIndex)device));
  return {t, py::bytes(reinterpret_cast<const char*>(&h), sizeof(h))};
}

// An independent tensor (own version counter, own autograd identity) over bytes [off, off + nbytes) of a chunk
// allocation; it keeps the chunk alive.  Sub-allocations must NOT be slices of one base tensor: an in-place write
// to a gradient buffer would then bump the version of every parameter saved for backward.
Tensor symm_alias(const Tensor& chunk, int64_t off, int64_t nbytes) {
  TORCH_CHECK(chunk.scalar_type() == at::kByte && off >= 0 && nbytes >= 0 && off + nbytes <= chunk.numel(), "bad alias");
  Tensor keep = chunk;
  return torch::from_blob((char*)chunk.data_ptr() + off, {nbytes}, [keep](void*) {}, chunk.options());
}

uint64_t symm_open(const std::string& handle, int64_t device) {
  TORCH_CHECK(handle.size() == sizeof(cudaIpcMemHandle_t), "bad IPC handle size");
  const c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle.data(), sizeof(h));
  void* p = nullptr;
  DTG_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  return (uint64_t)p;
}

void symm_close(uint64_t ptr) { cudaIpcCloseMemHandle((void*)ptr); }

int64_t symm_allocated_bytes() { return (int64_t)g_local_bytes; }

SymmPtrs rotated(const std::vector<uint64_t>& ptrs, int rank) {
  SymmPtrs s{};
  const int n = (int)ptrs.size();
  TORCH_CHECK(n >= 1 && n <= kMaxRanks, "1..8 ranks supported");
  for (int k = 0; k < n; ++k) s.ptr[k] = (char*)ptrs[(rank + k) % n];
  return s;
}
SymmPads pads_of(const std::vector<uint64_t>& ptrs) {
  SymmPads s{};
  TORCH_CHECK(ptrs.size() <= (size_t)kMaxRanks, "1..8 ranks supported");
  for (size_t k = 0; k < ptrs.size(); ++k) s.ptr[k] = (uint32_t*)ptrs[k];
  return s;
}
int* err_ptr(const c10::optional<Tensor>& err) { return err.has_value() ? err->data_ptr<int>() : nullptr; }

void allreduce_scale(const std::vector<uint64_t>& buf, const std::vector<uint64_t>& pads, int64_t elem_off, int64_t n,
                     double scale, int64_t rank, int64_t epoch, const c10::optional<Tensor>& err, int64_t blocks) {
  comm_allreduce_scale(rotated(buf, (int)rank), pads_of(pads), (size_t)elem_off, (size_t)n, (float)scale, (int)rank,
                       (int)buf.size(), (uint32_t)epoch, err_ptr(err), (int)blocks, stream());
}

void rs_adamw(const std::vector<uint64_t>& grads, const std::vector<uint64_t>& params,
              const c10::optional<Tensor>& param_local, Tensor& m, Tensor& v, bool push_params,
              const std::vector<uint64_t>& pads, int64_t elem_off, int64_t n, double lr, double b1, double b2, double eps,
              double wd, int64_t step, double grad_scale, int64_t rank, int64_t epoch,
              const c10::optional<Tensor>& err, int64_t blocks) {
  const bool fp32 = m.scalar_type() == at::kFloat;
  TORCH_CHECK(m.scalar_type() == v.scalar_type() && (fp32 || m.scalar_type() == at::kBFloat16), "bad state dtype");
  const int nr = (int)grads.size();
  TORCH_CHECK(m.numel() * nr == n && v.numel() * nr == n, "optimizer shard must hold n / nranks elements");
  AdamWHyper hp = make_adamw_hyper((float)lr, (float)b1, (float)b2, (float)eps, (float)wd, (int)step, (float)grad_scale);
  void* pl = nullptr;
  if (!push_params) {
    TORCH_CHECK(param_local.has_value() && param_local->numel() * nr == n, "param shard must hold n / nranks elements");
    pl = param_local->data_ptr();
  }
  comm_rs_adamw(rotated(grads, (int)rank), push_params ? rotated(params, (int)rank) : SymmPtrs{}, pl, m.data_ptr(),
                v.data_ptr(), fp32, push_params, pads_of(pads), (size_t)elem_off, (size_t)n, hp, (int)rank, nr,
                (uint32_t)epoch, err_ptr(err), (int)blocks, stream());
}

// ---- NVLS (multicast) variants ----------------------------------------------------------------------------------
void nvls_allreduce_scale(uint64_t mc, const std::vector<uint64_t>& pads, int64_t elem_off, int64_t n, double scale,
                          int64_t rank, int64_t epoch, const c10::optional<Tensor>& err, int64_t blocks) {
  comm_nvls_allreduce_scale((void*)mc, pads_of(pads), (size_t)elem_off, (size_t)n, (float)scale, (int)rank,
                            (int)pads.size(), (uint32_t)epoch, err_ptr(err), (int)blocks, stream());
}

void nvls_rs_adamw(uint64_t grads_mc, uint64_t params_mc, uint64_t params_local, Tensor& m, Tensor& v, bool push_params,
                   const std::vector<uint64_t>& pads, int64_t elem_off, int64_t n, double lr, double b1, double b2,
                   double eps, double wd, int64_t step, double grad_scale, int64_t rank, int64_t epoch,
                   const c10::optional<Tensor>& err, int64_t blocks) {
  const bool fp32 = m.scalar_type() == at::kFloat;
  TORCH_CHECK(m.scalar_type() == v.scalar_type() && (fp32 || m.scalar_type() == at::kBFloat16), "bad state dtype");
  const int nr = (int)pads.size();
  TORCH_CHECK(m.numel() * nr == n && v.numel() * nr == n, "optimizer shard must hold n / nranks elements");
  AdamWHyper hp = make_adamw_hyper((float)lr, (float)b1, (float)b2, (float)eps, (float)wd, (int)step, (float)grad_scale);
  comm_nvls_rs_adamw((const void*)grads_mc, (void*)params_mc, (void*)params_local, m.data_ptr(), v.data_ptr(), fp32,
                     push_params, pads_of(pads), (size_t)elem_off, (size_t)n, hp, (int)rank, nr, (uint32_t)epoch,
                     err_ptr(err), (int)blocks, stream());
}

void allgather(const std::vector<uint64_t>& shards, Tensor& full, const std::vector<uint64_t>& pads, int64_t shard_off,
               int64_t per, int64_t rank, int64_t epoch, const c10::optional<Tensor>& err, bool barrier,
               int64_t blocks) {
  TORCH_CHECK(full.is_contiguous() && full.scalar_type() == at::kBFloat16, "full must be contiguous bf16");
  TORCH_CHECK(full.numel() >= per * (int64_t)shards.size(), "full buffer too small");
  if (blocks <= 0) {  // copy-engine variant
    comm_allgather_ce(rotated(shards, (int)rank), full.data_ptr(), pads_of(pads), (size_t)shard_off, (size_t)per,
                      (int)rank, (int)shards.size(), (uint32_t)epoch, err_ptr(err), barrier, stream());
    return;
  }
  comm_allgather(rotated(shards, (int)rank), full.data_ptr(), pads_of(pads), (size_t)shard_off, (size_t)per, (int)rank,
                 (int)shards.size(), (uint32_t)epoch, err_ptr(err), barrier, (int)blocks, stream());
}

void gather_range(const std::vector<uint64_t>& shards, Tensor& full, const std::vector<uint64_t>& pads, int64_t begin,
                  int64_t end, int64_t per, int64_t rank, int64_t epoch, const c10::optional<Tensor>& err, bool barrier) {
  TORCH_CHECK(full.is_contiguous() && full.scalar_type() == at::kBFloat16, "full must be contiguous bf16");
  TORCH_CHECK(0 <= begin && begin <= end && end <= full.numel(), "bad range");
  SymmPtrs sp{};
  TORCH_CHECK(shards.size() <= (size_t)kMaxRanks, "1..8 ranks supported");
  for (size_t k = 0; k < shards.size(); ++k) sp.ptr[k] = (char*)shards[k];
  comm_gather_range_ce(sp, full.data_ptr(), pads_of(pads), (size_t)begin, (size_t)end, (size_t)per, (int)rank,
                       (int)shards.size(), (uint32_t)epoch, err_ptr(err), barrier, stream());
}

void reduce_scatter(const std::vector<uint64_t>& grads, Tensor& out, const std::vector<uint64_t>& pads, int64_t elem_off,
                    int64_t n, double scale, int64_t rank, int64_t epoch, const c10::optional<Tensor>& err,
                    int64_t blocks) {
  TORCH_CHECK(out.is_contiguous() && out.scalar_type() == at::kBFloat16 && out.numel() * (int64_t)grads.size() == n,
              "out must be a contiguous bf16 shard of n / nranks elements");
  comm_reduce_scatter(rotated(grads, (int)rank), out.data_ptr(), pads_of(pads), (size_t)elem_off, (size_t)n,
                      (float)scale, (int)rank, (int)grads.size(), (uint32_t)epoch, err_ptr(err), (int)blocks, stream());
}

void barrier(const std::vector<uint64_t>& pads, int64_t rank, int64_t epoch, const c10::optional<Tensor>& err) {
  comm_barrier(pads_of(pads), (int)rank, (int)pads.size(), (uint32_t)epoch, err_ptr(err), stream());
}

}  // namespace

void bind_comm(pybind11::module_& m) {
  m.def("symm_alloc", &symm_alloc);
  m.def("symm_open", &symm_open);
  m.def("symm_alias", &symm_alias);
  m.def("symm_close", &symm_close);
  m.def("symm_allocated_bytes", &symm_allocated_bytes);
  m.attr("SYMM_PAD_BYTES") = (int64_t)kPadBytes;
  m.attr("SYMM_MAX_CHANNELS") = (int64_t)kMaxChannels;
  m.def("comm_allreduce_scale", &allreduce_scale);
  m.def("comm_rs_adamw", &rs_adamw);
  m.def("comm_allgather", &allgather);
  m.def("comm_nvls_allreduce_scale", &nvls_allreduce_scale);
  m.def("comm_nvls_rs_adamw", &nvls_rs_adamw);
  m.def("comm_reduce_scatter", &reduce_scatter);
  m.def("comm_gather_range", &gather_range);
  m.def("comm_barrier", &barrier);
}
}  // namespace dtg
