// Symmetric-memory arena on the CUDA virtual-memory-management API, with NVSwitch multicast binding.
//
// The reference never owns its communication buffers (NCCL registers whatever torch hands it).  Here one
// *chunk* is: a physical allocation per rank (cuMemCreate, exportable as a POSIX fd), every peer's allocation
// imported and mapped into ONE contiguous virtual range laid out [rank 0 | rank 1 | ... | rank N-1] (so a peer
// pointer is `base + r * size + offset`, no per-buffer tables to exchange), and — when the fabric supports it —
// an NVLS multicast object (cuMulticastCreate / AddDevice / BindMem) mapped at a second range: a store to that
// address lands in every rank's copy, a `multimem.ld_reduce` from it returns the in-switch sum.
// parallel/symm.py drives the collective part (fd exchange over a Unix socket, ordering through the c10d store —
// no NCCL anywhere in set-up) and sub-allocates buffers from chunks, so a 7B data-parallel engine needs a handful
// of exchanges instead of one per flat group.
//
// The driver entry points are fetched with cudaGetDriverEntryPoint: the extension must import on machines
// without libcuda (the CPU test/build container).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <pybind11/stl.h>
#include <torch/extension.h>
#include <unistd.h>

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "comm_api.h"

namespace dtg {
namespace {

template <typename Fn>
Fn driver_fn(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
    cudaGetLastError();
    throw std::runtime_error(std::string("CUDA driver entry point unavailable: ") + name);
  }
  return reinterpret_cast<Fn>(p);
}
#define DRV(name) driver_fn<decltype(&name)>(#name)

void cu_check(CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return;
  const char* s = nullptr;
  try {
    DRV(cuGetErrorString)(r, &s);
  } catch (...) {
  }
  throw std::runtime_error(std::string(what) + " failed: " + (s ? s : "unknown") + " (CUresult " +
                           std::to_string((int)r) + ")");
}

size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

CUmemAllocationProp alloc_prop(int device) {
  CUmemAllocationProp prop{};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

// (posix-fd export supported, multicast supported) on `device`
std::pair<bool, bool> vmm_support(int64_t device) {
  try {
    int fd_ok = 0, mc_ok = 0;
    CUdevice dev;
    const c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
    cudaFree(nullptr);  // make sure the primary context exists
    cu_check(DRV(cuDeviceGet)(&dev, (int)device), "cuDeviceGet");
    auto attr = DRV(cuDeviceGetAttribute);
    cu_check(attr(&fd_ok, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev), "cuDeviceGetAttribute");
    if (attr(&mc_ok, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS) mc_ok = 0;
    return {fd_ok != 0, mc_ok != 0};
  } catch (const std::exception&) {
    return {false, false};
  }
}

class VmmChunk : public std::enable_shared_from_this<VmmChunk> {
 public:
  // `nbytes` is rounded up to the allocation (and, if `want_mc`, multicast) granularity; every rank of the group
  // must pass the same value.
  VmmChunk(int64_t device, int64_t nbytes, int64_t world, int64_t rank, bool want_mc)
      : device_((int)device), world_((int)world), rank_((int)rank), want_mc_(want_mc) {
    TORCH_CHECK(world >= 1 && rank >= 0 && rank < world, "bad rank/world");
    const c10::cuda::CUDAGuard guard((c10::DeviceIndex)device_);
    cudaFree(nullptr);
    const CUmemAllocationProp prop = alloc_prop(device_);
    size_t gran = 0;
    cu_check(DRV(cuMemGetAllocationGranularity)(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED),
             "cuMemGetAllocationGranularity");
    if (want_mc_) {
      CUmulticastObjectProp mp = mc_prop(gran);
      size_t mg = 0;
      cu_check(DRV(cuMulticastGetGranularity)(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED),
               "cuMulticastGetGranularity");
      gran = std::max(gran, mg);
    }
    size_ = round_up((size_t)std::max<int64_t>(nbytes, 1), gran);
    cu_check(DRV(cuMemCreate)(&local_, size_, &prop, 0), "cuMemCreate");
    handles_.assign((size_t)world_, 0);
    handles_[(size_t)rank_] = local_;
  }

  ~VmmChunk() { release(); }

  int64_t size() const { return (int64_t)size_; }
  uint64_t base() const { return (uint64_t)va_; }
  uint64_t mc_base() const { return (uint64_t)mc_va_; }

  // ---- unicast: export mine, import the peers', map all of them side by side -------------------
  int64_t export_fd() {
    int fd = -1;
    cu_check(DRV(cuMemExportToShareableHandle)(&fd, local_, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
             "cuMemExportToShareableHandle");
    return fd;
  }
  void import_peer(int64_t peer, int64_t fd) {
    TORCH_CHECK(peer >= 0 && peer < world_ && peer != rank_, "bad peer");
    CUmemGenericAllocationHandle h;
    cu_check(DRV(cuMemImportFromShareableHandle)(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
             "cuMemImportFromShareableHandle");
    handles_[(size_t)peer] = h;
  }
  void map_all() {
    const c10::cuda::CUDAGuard guard((c10::DeviceIndex)device_);
    for (auto h : handles_) TORCH_CHECK(h != 0, "a peer allocation was not imported");
    cu_check(DRV(cuMemAddressReserve)(&va_, size_ * (size_t)world_, 0, 0, 0), "cuMemAddressReserve");
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device_;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    for (int r = 0; r < world_; ++r) {
      cu_check(DRV(cuMemMap)(va_ + (size_t)r * size_, size_, 0, handles_[(size_t)r], 0), "cuMemMap");
      mapped_ = r + 1;
    }
    cu_check(DRV(cuMemSetAccess)(va_, size_ * (size_t)world_, &acc, 1), "cuMemSetAccess");
    C10_CUDA_CHECK(cudaMemset((void*)(va_ + (size_t)rank_ * size_), 0, size_));
    C10_CUDA_CHECK(cudaDeviceSynchronize());
  }

  // ---- multicast: rank 0 creates + exports, everyone adds its device, binds its memory, maps ----------
  int64_t mc_create_export() {
    CUmulticastObjectProp mp = mc_prop(size_);
    cu_check(DRV(cuMulticastCreate)(&mc_, &mp), "cuMulticastCreate");
    have_mc_ = true;
    int fd = -1;
    cu_check(DRV(cuMemExportToShareableHandle)(&fd, mc_, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
             "cuMemExportToShareableHandle(multicast)");
    return fd;
  }
  void mc_import(int64_t fd) {
    cu_check(DRV(cuMemImportFromShareableHandle)(&mc_, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
             "cuMemImportFromShareableHandle(multicast)");
    have_mc_ = true;
  }
  void mc_add_device() {
    CUdevice dev;
    cu_check(DRV(cuDeviceGet)(&dev, device_), "cuDeviceGet");
    cu_check(DRV(cuMulticastAddDevice)(mc_, dev), "cuMulticastAddDevice");
  }
  // call after EVERY rank has added its device
  void mc_bind_and_map() {
    const c10::cuda::CUDAGuard guard((c10::DeviceIndex)device_);
    cu_check(DRV(cuMulticastBindMem)(mc_, 0, local_, 0, size_, 0), "cuMulticastBindMem");
    mc_bound_ = true;
    cu_check(DRV(cuMemAddressReserve)(&mc_va_, size_, 0, 0, 0), "cuMemAddressReserve(multicast)");
    cu_check(DRV(cuMemMap)(mc_va_, size_, 0, mc_, 0), "cuMemMap(multicast)");
    mc_mapped_ = true;
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device_;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    cu_check(DRV(cuMemSetAccess)(mc_va_, size_, &acc, 1), "cuMemSetAccess(multicast)");
  }

  // uint8 view of this rank's slot [offset, offset + nbytes); the tensor keeps the chunk alive
  torch::Tensor local_view(int64_t offset, int64_t nbytes) {
    TORCH_CHECK(va_ != 0, "chunk is not mapped");
    TORCH_CHECK(offset >= 0 && nbytes >= 0 && (size_t)(offset + nbytes) <= size_, "view outside the chunk");
    auto self = shared_from_this();
    void* p = (void*)(va_ + (size_t)rank_ * size_ + (size_t)offset);
    return torch::from_blob(p, {nbytes}, [self](void*) {},
                            torch::TensorOptions().dtype(torch::kUInt8).device(torch::kCUDA, (c10::DeviceIndex)device_));
  }

  void release() {
    if (released_) return;
    released_ = true;
    try {
      auto unmap = DRV(cuMemUnmap);
      auto rel = DRV(cuMemRelease);
      auto afree = DRV(cuMemAddressFree);
      cudaDeviceSynchronize();
      if (mc_mapped_) unmap(mc_va_, size_);
      if (mc_va_) afree(mc_va_, size_);
      if (mc_bound_) {
        CUdevice dev;
        if (DRV(cuDeviceGet)(&dev, device_) == CUDA_SUCCESS) DRV(cuMulticastUnbind)(mc_, dev, 0, size_);
      }
      if (have_mc_) rel(mc_);
      for (int r = 0; r < mapped_; ++r) unmap(va_ + (size_t)r * size_, size_);
      if (va_) afree(va_, size_ * (size_t)world_);
      for (auto h : handles_)
        if (h) rel(h);
    } catch (...) {  // interpreter/driver shutdown: nothing left to release against
    }
  }

 private:
  CUmulticastObjectProp mc_prop(size_t size) const {
    CUmulticastObjectProp mp{};
    mp.numDevices = (unsigned)world_;
    mp.size = size;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return mp;
  }

  int device_, world_, rank_;
  bool want_mc_;
  size_t size_ = 0;
  CUmemGenericAllocationHandle local_ = 0, mc_ = 0;
  std::vector<CUmemGenericAllocationHandle> handles_;
  CUdeviceptr va_ = 0, mc_va_ = 0;
  int mapped_ = 0;
  bool have_mc_ = false, mc_bound_ = false, mc_mapped_ = false, released_ = false;
};

}  // namespace

void bind_symm_vmm(pybind11::module_& m) {
  m.def("vmm_support", &vmm_support);
  pybind11::class_<VmmChunk, std::shared_ptr<VmmChunk>>(m, "VmmChunk")
      .def(pybind11::init<int64_t, int64_t, int64_t, int64_t, bool>(), pybind11::arg("device"), pybind11::arg("nbytes"),
           pybind11::arg("world"), pybind11::arg("rank"), pybind11::arg("want_multicast") = false)
      .def("size", &VmmChunk::size)
      .def("base", &VmmChunk::base)
      .def("mc_base", &VmmChunk::mc_base)
      .def("export_fd", &VmmChunk::export_fd)
      .def("import_peer", &VmmChunk::import_peer)
      .def("map_all", &VmmChunk::map_all)
      .def("mc_create_export", &VmmChunk::mc_create_export)
      .def("mc_import", &VmmChunk::mc_import)
      .def("mc_add_device", &VmmChunk::mc_add_device)
      .def("mc_bind_and_map", &VmmChunk::mc_bind_and_map)
      .def("local_view", &VmmChunk::local_view)
      .def("release", &VmmChunk::release);
}
}  // namespace dtg
