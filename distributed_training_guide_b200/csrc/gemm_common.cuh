// Tile configuration shared by the tcgen05 GEMM and the fused GEMM+collective kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "common.cuh"

#ifndef DTG_DEFAULT_GEMM_VARIANT
#define DTG_DEFAULT_GEMM_VARIANT 3
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
  // B_MODE 3 (operand B gathered from the ranks' FSDP shards by warp 3 of every CTA): a 2-slot bounce ring
  static constexpr int GATHER_PIECE = 16384;
  static constexpr int GATHER_BYTES = 2 * GATHER_PIECE;
};

template <int N>
struct TmapSet {
  CUtensorMap m[N];
};

// Geometry of a tensor-parallel GEMM (all zero / unused for the plain GEMM)
struct GemmDist {
  int rows_per_peer;                 // rows (M- or K-direction) each rank contributes / owns
  int m_tile_shift;                  // rotate the M tile order so every rank starts on its own rows
  int k_shift;                       // rotate the K block order likewise
  __nv_bfloat16* c_ptr[kMaxRanks];   // C_MODE 1: destination base (already offset to my slot) per owner
  // A_MODE 3 (all-gather by communication CTAs inside the GEMM kernel)
  const char* ag_src[kMaxRanks];     // the symmetric [M, K] buffer on every rank, ROTATED: [0] = mine
  uint32_t* ag_flags;                // local: one word per 256-row tile, set to ag_epoch once the tile landed
  uint32_t ag_epoch;
  uint32_t* pads[kMaxRanks];         // signal pads (not rotated) for the start-of-kernel barrier
  uint32_t bar_epoch;
  int n_comm;                        // clusters (CTA pairs) that copy instead of multiplying
  int rank, nranks;
  long long tile_bytes;              // bytes of one 256-row tile of A (contiguous: lda == K)
  // B_MODE 3 (FSDP unshard inside the consuming GEMM): operand B is a weight whose bytes [bg_begin, bg_end) of the
  // group's flat layout are spread over the ranks' shards (rank p owns flat bytes [p*per, (p+1)*per)).  Warp 3 of
  // every CTA bounces 16 KB pieces shard -> smem -> local full buffer and counts them per `1 << bg_chunk_shift`
  // byte chunk; the TMA producer acquires the counters of the chunks under a B box before loading it.
  const char* bg_src[kMaxRanks];     // shard base per rank (NOT rotated; [rank] is my own shard)
  char* bg_dst;                      // local full (unsharded) flat buffer of the group
  long long bg_per_bytes;            // shard size in bytes
  long long bg_begin, bg_end;        // flat byte range of B (piece aligned; chunk aligned at both ends)
  uint32_t* bg_cnt;                  // one counter per chunk of the flat buffer (monotonic over generations)
  uint32_t bg_target;                // counter value at which a chunk of this generation is complete
  int bg_chunk_shift;                // log2(chunk bytes)
  int bg_row_bytes;                  // bytes of one row of B as stored (ldb * 2)
  int bg_rows;                       // rows of B as stored
  int n_tile_shift;                  // rotate the N tile order so every rank starts on the rows it owns
  // L2-aware rasterisation of the plain GEMM: tiles run M-fastest inside groups of `group_m` row tiles (0 = one
  // group = the whole M extent); `num_n_tiles` is set by the launcher.
  int group_m, num_n_tiles;
};

#ifdef __CUDACC__
__host__ __device__ __forceinline__ int tile_m(int t, int num_m_tiles, const GemmDist& d) {
  int m = t % num_m_tiles + d.m_tile_shift;
  return m >= num_m_tiles ? m - num_m_tiles : m;
}
// Tile order.  Default: M fastest (consecutive CTAs share the B tile) within a group of `group_m` row tiles
// whose slice of A stays L2-resident while B streams past once per group: the ~74 tiles in flight then touch
// group_m row panels of A and 74/group_m column panels of B instead of every row panel of A, which is what keeps
// a tall GEMM (wgrad with M = 22016 or 32000) from re-streaming all of A from HBM for every column of tiles.
// With an in-kernel all-gather
// (`local_m_tiles` > 0): first every tile of my own rows (pure local work while the communication CTAs
// fetch), then the remote row tiles in the order they are being fetched.
__host__ __device__ __forceinline__ void tile_mn(int t, int num_m_tiles, const GemmDist& d, int local_m_tiles, int& m, int& n) {
  if (local_m_tiles <= 0) {
    if (d.group_m > 0 && d.group_m < num_m_tiles) {
      const int per_group = d.group_m * d.num_n_tiles;
      const int g = t / per_group, r = t - g * per_group;
      const int m0 = g * d.group_m;
      const int rest_m = num_m_tiles - m0;
      const int gsz = d.group_m < rest_m ? d.group_m : rest_m;
      m = m0 + r % gsz;
      n = r / gsz;
      return;
    }
    m = tile_m(t, num_m_tiles, d);
    n = t / num_m_tiles + d.n_tile_shift;
    if (n >= d.num_n_tiles && d.n_tile_shift) n -= d.num_n_tiles;
    return;
  }
  const int num_n = d.k_shift;  // reused field: number of N tiles (K is never gathered in this mode)
  const int phase_a = local_m_tiles * num_n;
  int mm;
  if (t < phase_a) {
    mm = t % local_m_tiles;
    n = t / local_m_tiles;
  } else {
    const int u = t - phase_a, rest = num_m_tiles - local_m_tiles;
    mm = local_m_tiles + u % rest;
    n = u / rest;
  }
  mm += d.m_tile_shift;
  m = mm >= num_m_tiles ? mm - num_m_tiles : mm;
}
#endif

// TMA descriptor builders (gemm_tcgen05.cu)
CUtensorMap make_tmap_bf16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box, bool swizzle128);
CUtensorMap make_tmap_2d(const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                         uint32_t box_inner, uint32_t box_outer);
void set_gemm_variant(int v);
int default_gemm_variant();

}  // namespace dtg
