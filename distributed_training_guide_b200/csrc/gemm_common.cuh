// Tile configuration shared by the tcgen05 GEMM and the fused GEMM+collective kernels.
#pragma once
#include <cuda.h>
#include <cstdint>

#ifndef DTG_DEFAULT_GEMM_VARIANT
#define DTG_DEFAULT_GEMM_VARIANT 1
#endif

namespace dtg {

template <int CG>
struct GemmCfg {
  static constexpr int BM = 128;          // C rows per CTA (UMMA_M = BM * CG)
  static constexpr int BN = 256;          // C columns per tile (UMMA_N)
  static constexpr int BK = 64;           // one 128-byte swizzle span of bf16 per stage
  static constexpr int B_ROWS = BN / CG;  // rows of B (N) this CTA stages; the pair shares B
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (CG == 1) ? 4 : 6;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment slack
};

// TMA descriptor builders (gemm_tcgen05.cu)
CUtensorMap make_tmap_bf16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box, bool swizzle128);
CUtensorMap make_tmap_2d(const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                         uint32_t box_inner, uint32_t box_outer);
void set_gemm_variant(int v);
int default_gemm_variant();

}  // namespace dtg
