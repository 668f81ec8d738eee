// Device-side pieces shared by the NVLink collective kernels (comm.cu, comm_nvls.cu).
#pragma once
#include "comm.cuh"
#include "common.cuh"
#include "ptx.cuh"

#ifdef __CUDACC__
namespace dtg {
using namespace ptx;

// ---- device-side barrier --------------------------------------------------------------------
// pads[p] -> rank p's pad array: uint32 [kMaxChannels][kMaxRanks].  Block b uses channel b.
__device__ __forceinline__ void symm_barrier(uint32_t* const* pads, int rank, int nranks, int channel, uint32_t epoch,
                                             int* error_flag) {
  __syncthreads();
  if ((int)threadIdx.x < nranks) {
    const int p = threadIdx.x;
    __threadfence_system();
    st_release_sys(pads[p] + channel * kMaxRanks + rank, epoch);
    const uint32_t* mine = pads[rank] + channel * kMaxRanks + p;
    const long long t0 = clock64();
    // fail fast: once any barrier of this group has timed out (a peer died or diverged) later barriers do not
    // spin for another ~10 s each; the host raises at its next check_health()
    const bool poisoned = error_flag != nullptr && *reinterpret_cast<volatile int*>(error_flag) != 0;
    while (!poisoned && (int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      if (clock64() - t0 > kSpinTimeoutCycles) {  // a dead peer must not hang the GPU forever
        if (error_flag) atomicExch(error_flag, 1 + p);
        break;
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void add8(float (&acc)[8], const uint4& v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    acc[2 * i] += f.x;
    acc[2 * i + 1] += f.y;
  }
}
__device__ __forceinline__ uint4 pack8_u4(const float (&f)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return r;
}

}  // namespace dtg
#endif
