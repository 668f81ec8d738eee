// Causal flash-attention forward, second generation (head_dim 128, GQA): TWO 128-query tiles per CTA share every
// K/V block, each tile has its own softmax warpgroup, and P never leaves tensor memory.
//
//   warp 0      TMA producer    Q_A, Q_B once; K_j / V_j into 2-stage rings (one load serves both tiles)
//   warp 1      MMA issuer      S_t = Q_t K_j^T (SS form) and O_t += P_t V_j (TS form: A = P_t read from TMEM), the two
//                               tiles interleaved so the tensor core runs tile B's GEMMs while tile A is in softmax
//   warp 2      TMEM allocator  512 columns: S_A | S_B | O_A | O_B;  P_t overwrites the first 64 columns of S_t
//   warps 4-7   softmax tile A  one thread per query row: pass 1 row max (tcgen05.ld in 32-column chunks), pass 2
//   warps 8-11  softmax tile B  exp2 -> bf16 pairs -> tcgen05.st over the scores just consumed; the running max only
//                               moves (and O is only rescaled, by the same thread, in TMEM) when it grew by more than
//                               2^8 — the stale max is exact after the final 1/l normalisation
//
// Why: the first-generation kernel (attention_fwd.cu) had one tile per CTA, so QK -> softmax -> PV was a serial chain
// with the tensor pipe 31 % busy (profiles/prof_attn.md); here the chain of one tile hides under the other's, K/V
// traffic per query is halved, and the P round trip through 32 KB of shared memory is gone.
// Replaces torch SDPA / flash-attn-2 that the reference uses.
#include <cuda.h>

#include <type_traits>

#include "api.h"
#include "attention_common.cuh"
#include "common.cuh"
#include "gemm_common.cuh"
#include "ptx.cuh"

namespace dtg {
using namespace ptx;

CUtensorMap make_tmap_heads(const void* base, int B, int S, int heads, int box_rows);  // attention_fwd.cu

namespace fwd2 {
constexpr int BM = 128, BN = 128, D = 128;
constexpr int TILE_BYTES = 128 * 128 * 2;  // 32 KB: two 64-column halves of [128 rows x 128 B]
constexpr int HALF_BYTES = TILE_BYTES / 2;
constexpr int OFF_Q = 0;                   // [2] tiles
constexpr int OFF_K = 2 * TILE_BYTES;      // [2] stages
constexpr int OFF_V = 4 * TILE_BYTES;      // [2] stages
constexpr int OFF_BAR = 6 * TILE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr int THREADS = 384;
constexpr uint32_t TM_S = 0, TM_O = 256;   // + tile * 128
constexpr float RESCALE_LOG2 = 8.f;        // move the running max only when it grew by more than 2^8
}  // namespace fwd2

__global__ void __launch_bounds__(fwd2::THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ o, float* __restrict__ lse,
                 int S, int nh, int nkv, float scale_log2, int num_pairs) {
  using namespace fwd2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;    // [2]
  uint64_t* k_full = bars + 2;    // [2]
  uint64_t* v_full = bars + 4;    // [2]
  uint64_t* k_empty = bars + 6;   // [2]
  uint64_t* v_empty = bars + 8;   // [2]
  uint64_t* s_full = bars + 10;   // [2] per tile: S_t(j) is in TMEM (and every earlier MMA of the CTA has completed)
  uint64_t* p_full = bars + 12;   // [2] per tile: P_t(j) is in TMEM, O_t rescaled if needed
  uint64_t* o_done = bars + 14;   // [2] per tile: the last PV has completed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // longest rows first (grid = (B*nh, num_pairs), x fastest: every head's longest pair goes out first)
  const int pair = num_pairs - 1 - (int)blockIdx.y;
  const int head = blockIdx.x % nh;
  const int batch = blockIdx.x / nh;
  const int kv_head = head / (nh / nkv);
  const int q0 = pair * 2 * BM;
  const bool has_b = q0 + BM < S;                 // S may be an odd number of 128-row tiles
  const int nA = pair * 2 + 1;                    // causal, BM == BN: key blocks tile A attends to
  const int nB = has_b ? nA + 1 : 0;
  const int n_blocks = has_b ? nB : nA;

  if (warp == 0 && elect_one()) prefetch_tensormap(&tm_qkv);
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_done[i], 1);
    }
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
      for (int t = 0; t < (has_b ? 2 : 1); ++t) {
        mbar_arrive_expect_tx(&q_full[t], TILE_BYTES);
        tma_load_4d(&tm_qkv, &q_full[t], smem + OFF_Q + t * TILE_BYTES, 0, head, q0 + t * BM, batch);
        tma_load_4d(&tm_qkv, &q_full[t], smem + OFF_Q + t * TILE_BYTES + HALF_BYTES, 64, head, q0 + t * BM, batch);
      }
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
      auto n_of = [&](int t) { return t ? nB : nA; };
      auto issue_qk = [&](int t, int j) {            // S_t = Q_t K_j^T
        const int st = j & 1;
        const uint32_t sq = smem_u32(smem + OFF_Q + t * TILE_BYTES);
        const uint32_t sk = smem_u32(smem + OFF_K + st * TILE_BYTES);
        const uint32_t d_tm = tmem_base + TM_S + (uint32_t)t * 128;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (uint32_t)((kk >> 2) * HALF_BYTES + (kk & 3) * 32);
          mma_f16_ss<1>(d_tm, desc_kmajor_sw128(sq + off), desc_kmajor_sw128(sk + off), idesc_qk, kk ? 1u : 0u);
        }
        mma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int j) {            // O_t (+)= P_t V_j, P_t from tensor memory
        const int st = j & 1;
        const uint32_t sv = smem_u32(smem + OFF_V + st * TILE_BYTES);
        const uint32_t p_tm = tmem_base + TM_S + (uint32_t)t * 128;
        const uint32_t o_tm = tmem_base + TM_O + (uint32_t)t * 128;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          mma_f16_ts(o_tm, p_tm + (uint32_t)(kk * 8), desc_mnmajor_sw128(sv + (uint32_t)(kk * 2048), HALF_BYTES),
                     idesc_pv, (j | kk) ? 1u : 0u);
      };
      // The last tile that reads K_j / V_j releases the ring slot.  Tile B always has the longer key range.
      const int last_t = has_b ? 1 : 0;
      mbar_wait(&k_full[0], 0);
      for (int t = 0; t <= last_t; ++t) {
        mbar_wait(&q_full[t], 0);
        tc_fence_after();
        issue_qk(t, 0);
      }
      mma_commit(&k_empty[0]);
      for (int j = 0; j < n_blocks; ++j) {
        const int st = j & 1;
        const uint32_t jph = (uint32_t)(j & 1);
        bool v_waited = false, k_waited = false;
        for (int t = 0; t <= last_t; ++t) {
          if (j >= n_of(t)) continue;
          mbar_wait(&p_full[t], jph);
          if (!v_waited) {
            mbar_wait(&v_full[st], (uint32_t)((j >> 1) & 1));
            v_waited = true;
          }
          tc_fence_after();
          issue_pv(t, j);
          if (j + 1 == n_of(t)) mma_commit(&o_done[t]);
          if (t == last_t) mma_commit(&v_empty[st]);
          if (j + 1 < n_of(t)) {
            if (!k_waited) {
              mbar_wait(&k_full[(j + 1) & 1], (uint32_t)(((j + 1) >> 1) & 1));
              k_waited = true;
              tc_fence_after();
            }
            issue_qk(t, j + 1);
            if (t == last_t) mma_commit(&k_empty[(j + 1) & 1]);
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax + epilogue: one warpgroup per tile, one thread per query row =====================
    const int t = (warp - 4) >> 2;      // tile
    if (t == 0 || has_b) {
      const int q = warp & 3;           // TMEM lane quarter this warp may touch
      const int row = q * 32 + lane;    // query row within the tile == TMEM lane
      const int n_t = t ? nB : nA;
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      const uint32_t s_addr = lane_addr + TM_S + (uint32_t)t * 128;
      const uint32_t o_addr = lane_addr + TM_O + (uint32_t)t * 128;
      float m_use = -INFINITY, l_run = 0.f;   // m_use: the (possibly stale) row max every exponent is taken against
      // one block of scores; DIAG (the last block of the tile) masks keys after the query
      auto block = [&](auto diag_tag, int j) {
        constexpr bool DIAG = decltype(diag_tag)::value;
        // ---- pass 1: row maximum (two 64-column halves, both loads of a half in flight before the wait) ----
        float mx = -INFINITY;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32(s_addr + h * 64, r0);
          tmem_ld_32x32b_x32(s_addr + h * 64 + 32, r1);
          tmem_ld_wait();
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float a = __uint_as_float(r0[i]), b = __uint_as_float(r1[i]);
            if (DIAG) {
              if (h * 64 + i > row) a = -INFINITY;
              if (h * 64 + 32 + i > row) b = -INFINITY;
            }
            m4[i & 3] = fmaxf(m4[i & 3], fmaxf(a, b));
          }
          mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
        }
        // ---- lazy running max: O (and l) are rescaled only when some row of the warp saw its max grow by more
        //      than 2^RESCALE_LOG2.  The decision is WARP-UNIFORM: tcgen05.ld/st are .sync.aligned. ----
        const bool moved = (mx - m_use) * scale_log2 > RESCALE_LOG2;      // always true on the first block
        if (__any_sync(0xffffffffu, moved)) {
          const float m_new = fmaxf(mx, m_use);
          const float alpha = fast_exp2((m_use - m_new) * scale_log2);   // 0 on the first block (m_use = -inf)
          l_run *= alpha;
          m_use = m_new;
          if (j > 0) {   // s_full(j) was committed after PV_t(j-1): O_t is quiescent
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t r[32];
              tmem_ld_32x32b_x32(o_addr + c * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
              tmem_st_32x32b_x32(o_addr + c * 32, r);
            }
          }
        }
        // ---- pass 2: P = exp2(S * scale - m) -> bf16 pairs, written over the scores already consumed ----
        const float mb = m_use * scale_log2;
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32(s_addr + h * 64, r0);
          tmem_ld_32x32b_x32(s_addr + h * 64 + 32, r1);
          tmem_ld_wait();
          uint32_t pk[32];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0 = fast_exp2(fmaf(__uint_as_float(r0[i]), scale_log2, -mb));
            float p1 = fast_exp2(fmaf(__uint_as_float(r0[i + 1]), scale_log2, -mb));
            float p2 = fast_exp2(fmaf(__uint_as_float(r1[i]), scale_log2, -mb));
            float p3 = poly_exp2(fmaf(__uint_as_float(r1[i + 1]), scale_log2, -mb));
            if (DIAG) {
              if (h * 64 + i > row) p0 = 0.f;
              if (h * 64 + i + 1 > row) p1 = 0.f;
              if (h * 64 + 32 + i > row) p2 = 0.f;
              if (h * 64 + 32 + i + 1 > row) p3 = 0.f;
            }
            sum4[(i >> 1) & 3] += (p0 + p1) + (p2 + p3);
            __nv_bfloat162 lo = __floats2bfloat162_rn(p0, p1), hi = __floats2bfloat162_rn(p2, p3);
            pk[i >> 1] = *reinterpret_cast<uint32_t*>(&lo);
            pk[16 + (i >> 1)] = *reinterpret_cast<uint32_t*>(&hi);
          }
          tmem_st_32x32b_x32(s_addr + h * 32, pk);   // P columns [32h, 32h+32) lie inside the score columns already read
        }
        l_run += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      };
      // Interior blocks, common case in ONE pass over the scores: exponentiate against the current (stale) max while
      // tracking the true row max; only if some row of the warp outgrew the threshold is the block redone exactly
      // (the scores are still intact: P is kept in registers until the check).  Returns false if it must be redone.
      auto fast_block = [&]() -> bool {
        const float mb = m_use * scale_log2;
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t pk[64];
        // the whole 128-column row in flight behind ONE wait: the softmax phase is bound by tensor-memory round
        // trips (few warps, each serialising its tcgen05.ld -> wait), not by instruction issue or the SFU
        uint32_t ra[32], rb[32], rc[32], rd[32];
        tmem_ld_32x32b_x32(s_addr, ra);
        tmem_ld_32x32b_x32(s_addr + 32, rb);
        tmem_ld_32x32b_x32(s_addr + 64, rc);
        tmem_ld_32x32b_x32(s_addr + 96, rd);
        tmem_ld_wait();
        auto half = [&](const uint32_t(&r0)[32], const uint32_t(&r1)[32], int h) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float a0 = __uint_as_float(r0[i]), a1 = __uint_as_float(r0[i + 1]);
            const float b0 = __uint_as_float(r1[i]), b1 = __uint_as_float(r1[i + 1]);
            mx4[(i >> 1) & 3] = fmaxf(mx4[(i >> 1) & 3], fmaxf(fmaxf(a0, a1), fmaxf(b0, b1)));
            const float p0 = fast_exp2(fmaf(a0, scale_log2, -mb)), p1 = fast_exp2(fmaf(a1, scale_log2, -mb));
            // one exponential in four on the FMA pipes (poly_exp2): the SFU alone would be as busy as the tensor core
            const float p2 = fast_exp2(fmaf(b0, scale_log2, -mb)), p3 = poly_exp2(fmaf(b1, scale_log2, -mb));
            sum4[(i >> 1) & 3] += (p0 + p1) + (p2 + p3);
            __nv_bfloat162 lo = __floats2bfloat162_rn(p0, p1), hi = __floats2bfloat162_rn(p2, p3);
            pk[h * 32 + (i >> 1)] = *reinterpret_cast<uint32_t*>(&lo);
            pk[h * 32 + 16 + (i >> 1)] = *reinterpret_cast<uint32_t*>(&hi);
          }
        };
        half(ra, rb, 0);
        half(rc, rd, 1);
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        if (__any_sync(0xffffffffu, (mx - m_use) * scale_log2 > RESCALE_LOG2)) return false;
        uint32_t(&lo32)[32] = *reinterpret_cast<uint32_t(*)[32]>(&pk[0]);
        uint32_t(&hi32)[32] = *reinterpret_cast<uint32_t(*)[32]>(&pk[32]);
        tmem_st_32x32b_x32(s_addr, lo32);
        tmem_st_32x32b_x32(s_addr + 32, hi32);
        l_run += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
        return true;
      };
      for (int j = 0; j < n_t; ++j) {
        mbar_wait(&s_full[t], (uint32_t)(j & 1));
        tc_fence_after();
        if (j == n_t - 1) block(std::true_type{}, j);
        else if (j == 0 || !fast_block()) block(std::false_type{}, j);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }
      // ---- epilogue: O / l -> bf16 -> global ; logsumexp ----
      mbar_wait(&o_done[t], 0);
      tc_fence_after();
      const float inv_l = 1.f / l_run;
      const long long tok = (long long)batch * S + q0 + t * BM + row;
      __nv_bfloat16* orow = o + (tok * nh + head) * (long long)D;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(o_addr + c * 32, r);
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
      lse[((long long)batch * nh + head) * S + q0 + t * BM + row] =
          m_use * scale_log2 * 0.6931471805599453f + __logf(l_run);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

void attn_fwd2(const void* qkv, void* o, float* lse, int B, int S, int nh, int nkv, float scale, cudaStream_t s) {
  if (S % 128 != 0) throw std::runtime_error("attn_fwd: sequence length must be a multiple of 128");
  if (nh % nkv != 0) throw std::runtime_error("attn_fwd: nh must be a multiple of nkv");
  const CUtensorMap tm = make_tmap_heads(qkv, B, S, nh + 2 * nkv, 128);
  static bool attr = false;
  if (!attr) {
    DTG_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd2::SMEM_BYTES));
    attr = true;
  }
  const int num_pairs = (S / 128 + 1) / 2;
  attn_fwd2_kernel<<<dim3(B * nh, num_pairs, 1), fwd2::THREADS, fwd2::SMEM_BYTES, s>>>(
      tm, (__nv_bfloat16*)o, lse, S, nh, nkv, scale * 1.4426950408889634f, num_pairs);
  note_launch();
  DTG_LAUNCH_CHECK();
}

}  // namespace dtg
