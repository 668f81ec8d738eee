// Causal flash-attention forward on tcgen05 (head_dim 128, GQA), reading q/k/v straight out of
// the fused qkv activation [B, S, nh + 2*nkv, 128] through one strided 4-D TMA descriptor (no
// transposes, no repeat_kv copies).  One CTA = one (batch, q-head, 128-query block):
//
//   warp 0    TMA producer     Q once; K_j / V_j into 2-stage rings (separate barriers so QK^T
//                              can start as soon as K lands)
//   warp 1    MMA issuer       S[j%2] = Q K_j^T  (128x128x128, fp32 in TMEM, double-buffered so
//                              the next block's scores are computed during this block's softmax)
//                              O     += P_j V_j  (P from shared memory, V as an MN-major operand)
//   warp 2    TMEM allocator   512 columns: S0 | S1 | O
//   warps 4-11 softmax         two threads per query row (64 score columns each; two warps per SM
//                              sub-partition hide each other's latency): tcgen05.ld of the half row,
//                              online max (halves combined through shared memory + a 64-thread named
//                              barrier) / sum in fp32 with ex2.approx, P -> bf16 -> 128B-swizzled shared
//                              memory, O rescaled in TMEM only when some row's running max moved,
//                              final O / l and the logsumexp written from registers.
//
// Replaces torch SDPA / flash-attn-2 (mma.sync) that the reference uses.
#include <cuda.h>

#include "api.h"
#include "attention_common.cuh"
#include "common.cuh"
#include "gemm_common.cuh"
#include "ptx.cuh"

namespace dtg {
using namespace ptx;

namespace fwd {
constexpr int BM = 128, BN = 128, D = 128;
constexpr int TILE_BYTES = 128 * 128 * 2;  // 32 KB: two 64-column halves of [128 rows x 128 B]
constexpr int HALF_BYTES = TILE_BYTES / 2;
constexpr int OFF_Q = 0, OFF_K = TILE_BYTES, OFF_V = 3 * TILE_BYTES, OFF_P = 5 * TILE_BYTES;
constexpr int OFF_BAR = 6 * TILE_BYTES;
constexpr int OFF_RED = OFF_BAR + 256;          // float [3][2][128]: row max exchange (double-buffered by block
                                                // parity: a fast warp may already publish block j+1) + row sums
constexpr int SMEM_BYTES = OFF_RED + 3072 + 1024;
constexpr int THREADS = 384;
constexpr uint32_t TM_S0 = 0, TM_S1 = 128, TM_O = 256;
}  // namespace fwd

__global__ void __launch_bounds__(fwd::THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ o, float* __restrict__ lse,
                int S, int nh, int nkv, float scale_log2, int num_m_blocks) {
  using namespace fwd;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* v_full = bars + 3;    // [2]
  uint64_t* k_empty = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* s_empty = bars + 11;  // [2]
  uint64_t* p_full = bars + 13;
  uint64_t* pv_done = bars + 14;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // longest rows first: CTAs are dispatched in blockIdx order, causal work grows with the q block
  // (grid = (B*nh, num_m_blocks): x varies fastest, so every head's longest block goes out first)
  const int m_block = num_m_blocks - 1 - (int)blockIdx.y;
  const int head = blockIdx.x % nh;
  const int batch = blockIdx.x / nh;
  const int kv_head = head / (nh / nkv);
  const int n_blocks = m_block + 1;  // causal, BM == BN
  const int q0 = m_block * BM;

  if (warp == 0 && elect_one()) prefetch_tensormap(&tm_qkv);
  if (warp == 1 && elect_one()) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 8);
    }
    mbar_init(p_full, 8);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(tmem_ptr_smem, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (elect_one()) {
      // Q tile: two 64-column halves
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_4d(&tm_qkv, q_full, smem + OFF_Q, 0, head, q0, batch);
      tma_load_4d(&tm_qkv, q_full, smem + OFF_Q + HALF_BYTES, 64, head, q0, batch);
      const int kh = nh + kv_head, vh = nh + nkv + kv_head;
      for (int j = 0; j < n_blocks; ++j) {
        const int st = j & 1;
        const uint32_t ph = (uint32_t)((j >> 1) & 1);
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
        tma_load_4d(&tm_qkv, &k_full[st], smem + OFF_K + st * TILE_BYTES, 0, kh, j * BN, batch);
        tma_load_4d(&tm_qkv, &k_full[st], smem + OFF_K + st * TILE_BYTES + HALF_BYTES, 64, kh, j * BN, batch);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
        tma_load_4d(&tm_qkv, &v_full[st], smem + OFF_V + st * TILE_BYTES, 0, vh, j * BN, batch);
        tma_load_4d(&tm_qkv, &v_full[st], smem + OFF_V + st * TILE_BYTES + HALF_BYTES, 64, vh, j * BN, batch);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, false, true);
      const uint32_t sq = smem_u32(smem + OFF_Q), sp = smem_u32(smem + OFF_P);
      auto issue_qk = [&](int j) {
        const int st = j & 1;
        const uint32_t sk = smem_u32(smem + OFF_K + st * TILE_BYTES);
        const uint32_t d_tm = tmem_base + (st ? TM_S1 : TM_S0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (uint32_t)((kk >> 2) * HALF_BYTES + (kk & 3) * 32);
          mma_f16_ss<1>(d_tm, desc_kmajor_sw128(sq + off), desc_kmajor_sw128(sk + off), idesc_qk, kk ? 1u : 0u);
        }
        mma_commit(&s_full[st]);
        mma_commit(&k_empty[st]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0);
      for (int j = 0; j < n_blocks; ++j) {
        if (j + 1 < n_blocks) {
          const int st = (j + 1) & 1;
          const uint32_t ph = (uint32_t)(((j + 1) >> 1) & 1);
          mbar_wait(&k_full[st], ph);
          mbar_wait(&s_empty[st], ph ^ 1);  // softmax of block j-1 has drained this score buffer
          tc_fence_after();
          issue_qk(j + 1);
        }
        const int st = j & 1;
        mbar_wait(p_full, (uint32_t)(j & 1));
        mbar_wait(&v_full[st], (uint32_t)((j >> 1) & 1));
        tc_fence_after();
        const uint32_t sv = smem_u32(smem + OFF_V + st * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t da = desc_kmajor_sw128(sp + (uint32_t)((kk >> 2) * HALF_BYTES + (kk & 3) * 32));
          const uint64_t db = desc_mnmajor_sw128(sv + (uint32_t)(kk * 2048), HALF_BYTES);
          mma_f16_ss<1>(tmem_base + TM_O, da, db, idesc_pv, (j | kk) ? 1u : 0u);
        }
        mma_commit(pv_done);
        mma_commit(&v_empty[st]);
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax / correction / epilogue =====================
    const int q = warp & 3;             // TMEM lane quarter this warp may touch
    const int half = (warp - 4) >> 2;   // which 64 score columns (and which 64 output columns) are mine
    const int row = q * 32 + lane;      // query row within the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* red = reinterpret_cast<float*>(smem + OFF_RED);
    float m_run = -INFINITY, l_run = 0.f;  // l_run: partial row sum over my columns
    uint8_t* sp = smem + OFF_P + half * HALF_BYTES + row * 128;
    for (int j = 0; j < n_blocks; ++j) {
      const int st = j & 1;
      mbar_wait(&s_full[st], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      float s[64];
      {
        uint32_t r0[32], r1[32];  // both loads in flight before the single wait
        tmem_ld_32x32b_x32(lane_addr + (st ? TM_S1 : TM_S0) + half * 64, r0);
        tmem_ld_32x32b_x32(lane_addr + (st ? TM_S1 : TM_S0) + half * 64 + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          s[i] = __uint_as_float(r0[i]);
          s[32 + i] = __uint_as_float(r1[i]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);  // scores are in registers now
      if (j == n_blocks - 1) {                   // diagonal block: mask keys after the query
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (half * 64 + i > row) s[i] = -INFINITY;
      }
      float mx4[4] = {s[0], s[1], s[2], s[3]};  // four independent chains instead of one 64-deep one
#pragma unroll
      for (int i = 4; i < 64; i += 4) {
        mx4[0] = fmaxf(mx4[0], s[i]);
        mx4[1] = fmaxf(mx4[1], s[i + 1]);
        mx4[2] = fmaxf(mx4[2], s[i + 2]);
        mx4[3] = fmaxf(mx4[3], s[i + 3]);
      }
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      float* redj = red + (j & 1) * 256;
      redj[half * 128 + row] = mx;
      named_bar_sync(1 + q, 64);  // the two warps that share these 32 rows
      mx = fmaxf(m_run, fmaxf(mx, redj[(half ^ 1) * 128 + row]));
      const float alpha = fast_exp2((m_run - mx) * scale_log2);  // 0 on the first block (m_run = -inf)
      const float mb = mx * scale_log2;
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        const float p0 = fast_exp2(fmaf(s[i], scale_log2, -mb));
        const float p1 = fast_exp2(fmaf(s[i + 1], scale_log2, -mb));
        sum4[(i >> 1) & 3] += p0 + p1;
        __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
      }
      const float sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      l_run = l_run * alpha + sum;
      m_run = mx;
      // P_j may only overwrite the shared buffer / O may only be touched once PV_{j-1} is done
      if (j > 0) mbar_wait(pv_done, (uint32_t)((j - 1) & 1));
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(sp + ((c ^ (row & 7)) << 4)) =
            make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
      if (j > 0 && __any_sync(0xffffffffu, alpha < 1.f)) {
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(lane_addr + TM_O + half * 64 + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st_32x32b_x32(lane_addr + TM_O + half * 64 + c * 32, r);
        }
        tmem_st_wait();
      }
      fence_proxy_async();  // generic-proxy writes of P -> visible to the tensor core's async proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // epilogue: O / l -> bf16 -> global ; logsumexp
    red[512 + half * 128 + row] = l_run;
    named_bar_sync(1 + q, 64);
    const float l_tot = l_run + red[512 + (half ^ 1) * 128 + row];
    mbar_wait(pv_done, (uint32_t)((n_blocks - 1) & 1));
    tc_fence_after();
    const float inv_l = 1.f / l_tot;
    const long long tok = (long long)batch * S + q0 + row;
    __nv_bfloat16* orow = o + (tok * nh + head) * (long long)D + half * 64;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(lane_addr + TM_O + half * 64 + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[v * 8 + i]) * inv_l;
        st8(orow + c * 32 + v * 8, pack8(f));
      }
    }
    // natural-log logsumexp of the scaled scores
    if (half == 0)
      lse[((long long)batch * nh + head) * S + q0 + row] = m_run * scale_log2 * 0.6931471805599453f + __logf(l_tot);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

CUtensorMap make_tmap_heads(const void* base, int B, int S, int heads, int box_rows) {
  // [B, S, heads, 128] bf16 viewed as dims {128, heads, S, B}; box {64, 1, box_rows, 1}, 128B swizzle
  uint64_t dims[4] = {128, (uint64_t)heads, (uint64_t)S, (uint64_t)B};
  uint64_t strides[3] = {128ull * 2, (uint64_t)heads * 128 * 2, (uint64_t)S * heads * 128 * 2};
  uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
  return make_tmap_bf16(base, 4, dims, strides, box, true);
}

void attn_fwd(const void* qkv, void* o, float* lse, int B, int S, int nh, int nkv, float scale, cudaStream_t s) {
  if (S % 128 != 0) throw std::runtime_error("attn_fwd: sequence length must be a multiple of 128");
  if (nh % nkv != 0) throw std::runtime_error("attn_fwd: nh must be a multiple of nkv");
  const CUtensorMap tm = make_tmap_heads(qkv, B, S, nh + 2 * nkv, 128);
  static bool attr = false;
  if (!attr) {
    DTG_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd::SMEM_BYTES));
    attr = true;
  }
  const int num_m = S / 128;
  attn_fwd_kernel<<<dim3(B * nh, num_m, 1), fwd::THREADS, fwd::SMEM_BYTES, s>>>(tm, (__nv_bfloat16*)o, lse, S, nh, nkv,
                                                                      scale * 1.4426950408889634f, num_m);
  note_launch();
  DTG_LAUNCH_CHECK();
}

}  // namespace dtg
