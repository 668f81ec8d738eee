// NVLS (NVSwitch multicast) variants of the data-parallel bucket kernels.  EXPERIMENTAL: selected with DTG_NVLS=1
// when the symmetric buffers were bound to a multicast address (parallel/symm.py allocates them through
// torch.distributed._symmetric_memory in that mode); the peer-pointer kernels in comm.cu stay the default.
//
//   multimem.ld_reduce  one load returns the sum over every GPU's copy of the address, added inside the switch
//                       (fp32 accumulation): the reduce-scatter receives 1/N of the bytes of the pull version;
//   multimem.st         one store lands in every GPU's copy: the parameter all-gather sends each byte once.
//
// Per GPU and bucket the NVLink traffic drops from (N-1)/N * bytes in each direction for each of reduce-scatter
// and all-gather to that amount once in total, and no SM time goes into adding the peers' contributions.
// The device-side barrier protocol (epoch flags over the peer-mapped signal pad) is the one of comm.cu.
#include "adamw.cuh"
#include "comm.cuh"
#include "comm_device.cuh"
#include "common.cuh"
#include "ptx.cuh"

namespace dtg {
using namespace ptx;

__device__ __forceinline__ void unpack_u4(const uint4& v, float (&f)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = 0.f;
  add8(f, v);
}

// all-reduce (sum) with a fused scale: every rank reduces its 1/N slice in the switch and multicasts the result
__global__ void __launch_bounds__(kCommThreads) nvls_allreduce_scale_kernel(char* mc, SymmPads pads, size_t elem_off,
                                                                            size_t n, float scale, int rank, int nranks,
                                                                            uint32_t epoch, int* err) {
  symm_barrier(pads.ptr, rank, nranks, blockIdx.x, epoch, err);
  const size_t per = n / nranks;
  const size_t base = (elem_off + (size_t)rank * per) * 2;
  const size_t nvec = per / 8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    float acc[8];
    unpack_u4(multimem_ld_reduce_bf16x8(mc + base + i * 16), acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= scale;
    multimem_st_v4(mc + base + i * 16, pack8_u4(acc));
  }
  symm_barrier(pads.ptr, rank, nranks, blockIdx.x, epoch + 1, err);
}

// ZeRO-1 bucket (PUSH): in-switch reduce-scatter -> AdamW on my shard -> multicast the new parameters to every replica.
// FSDP (!PUSH): in-switch reduce-scatter -> AdamW on my parameter shard, which stays sharded (`params_local` is then
// the shard buffer, indexed from 0, and `params_mc` is unused).
template <typename StateT, bool PUSH>
__global__ void __launch_bounds__(kCommThreads) nvls_rs_adamw_kernel(const char* grads_mc, char* params_mc,
                                                                     char* params_local, StateT* m, StateT* v,
                                                                     SymmPads pads, size_t elem_off, size_t n,
                                                                     AdamWHyper hp, int rank, int nranks, uint32_t epoch,
                                                                     int* err) {
  symm_barrier(pads.ptr, rank, nranks, blockIdx.x, epoch, err);
  const size_t per = n / nranks;
  const size_t base = (elem_off + (size_t)rank * per) * 2;
  const size_t nvec = per / 8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    float g[8], p[8], fm[8], fv[8];
    unpack_u4(multimem_ld_reduce_bf16x8(grads_mc + base + i * 16), g);
    __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(PUSH ? params_local + base : params_local) + i * 8;
    unpack8(ld8(pl), p);
    load_state8(m + i * 8, fm);
    load_state8(v + i * 8, fv);
#pragma unroll
    for (int j = 0; j < 8; ++j) adamw_update(p[j], g[j], fm[j], fv[j], hp);  // hp.grad_scale carries 1/N
    store_state8(m + i * 8, fm);
    store_state8(v + i * 8, fv);
    if (PUSH) multimem_st_v4(params_mc + base + i * 16, pack8_u4(p));
    else *reinterpret_cast<uint4*>(pl) = pack8_u4(p);
  }
  symm_barrier(pads.ptr, rank, nranks, blockIdx.x, epoch + 1, err);
}

static void check_nvls(size_t n, int nranks, int blocks, const void* mc) {
  if (mc == nullptr) throw std::runtime_error("NVLS collective called without a multicast address");
  if (n % ((size_t)nranks * 8) != 0) throw std::runtime_error("collective size must be a multiple of 8*nranks elements");
  if (blocks < 1 || blocks > kMaxChannels) throw std::runtime_error("comm grid exceeds the signal-pad channels");
}

void comm_nvls_allreduce_scale(void* mc, const SymmPads& pads, size_t elem_off, size_t n, float scale, int rank,
                               int nranks, uint32_t epoch, int* err, int blocks, cudaStream_t s) {
  check_nvls(n, nranks, blocks, mc);
  nvls_allreduce_scale_kernel<<<blocks, kCommThreads, 0, s>>>((char*)mc, pads, elem_off, n, scale, rank, nranks, epoch,
                                                              err);
  note_launch();
  DTG_LAUNCH_CHECK();
}

void comm_nvls_rs_adamw(const void* grads_mc, void* params_mc, void* params_local, void* m, void* v, bool state_fp32,
                        bool push_params, const SymmPads& pads, size_t elem_off, size_t n, const AdamWHyper& hp, int rank,
                        int nranks, uint32_t epoch, int* err, int blocks, cudaStream_t s) {
  check_nvls(n, nranks, blocks, grads_mc);
  if (push_params) check_nvls(n, nranks, blocks, params_mc);
#define LAUNCH_NVLS(ST, PUSH)                                                                                         \
  nvls_rs_adamw_kernel<ST, PUSH><<<blocks, kCommThreads, 0, s>>>((const char*)grads_mc, (char*)params_mc,             \
                                                                 (char*)params_local, (ST*)m, (ST*)v, pads, elem_off, \
                                                                 n, hp, rank, nranks, epoch, err)
  if (state_fp32) {
    if (push_params) LAUNCH_NVLS(float, true); else LAUNCH_NVLS(float, false);
  } else {
    if (push_params) LAUNCH_NVLS(__nv_bfloat16, true); else LAUNCH_NVLS(__nv_bfloat16, false);
  }
#undef LAUNCH_NVLS
  note_launch();
  DTG_LAUNCH_CHECK();
}

}  // namespace dtg
