// Host-side launcher API of the sm_100a kernels (raw pointers + stream; no torch types so the
// .cu files compile in seconds).  bind.cpp adapts torch tensors onto these.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace dtg {

unsigned long long launch_count();

// ---- elementwise.cu ------------------------------------------------------------------------
void rmsnorm_fwd(const void* x, const void* res, const void* w, void* y, void* h_out, float* rstd, int T, int H,
                 float eps, cudaStream_t s);
int rmsnorm_bwd_grid(int T);
void rmsnorm_bwd(const void* dy, const void* h, const void* w, const float* rstd, const void* dres, void* dx,
                 float* dw_partial, float* dw, int T, int H, cudaStream_t s);
void rope_inplace(void* qkv, const float* cos, const float* sin, long long T, int S, int n_heads, int n_rot, int d,
                  bool per_token, bool inverse, cudaStream_t s);
void swiglu_fwd(const void* gu, void* h, long long T, int I, cudaStream_t s);
void swiglu_bwd(const void* dh, const void* gu, void* dgu, long long T, int I, cudaStream_t s);
void embedding_fwd(const long long* ids, const void* w, void* out, long long T, int H, cudaStream_t s);
void embedding_bwd(const void* dout, const long long* ids, void* dw, long long T, int H, cudaStream_t s);
void embedding_bwd_sorted(const void* dout, const long long* ids_sorted, const long long* perm, void* dw, long long T, int H,
                          bool accumulate, cudaStream_t s);
void scale_inplace(void* x, const float* scale, long long n, cudaStream_t s);

// ---- cross_entropy.cu ----------------------------------------------------------------------
// logits [T,V] bf16 are overwritten with dlogits = (softmax - onehot) / n_valid; loss = mean CE.
void cross_entropy_fwd_bwd(void* logits, const long long* targets, float* row_loss, float* n_valid, float* loss,
                           int T, int V, cudaStream_t s);

// ---- adamw.cu ------------------------------------------------------------------------------
// state_fp32: exp_avg / exp_avg_sq stored as fp32 instead of bf16
void adamw_flat(void* p, const void* g, void* m, void* v, long long n, float lr, float beta1, float beta2, float eps,
                float wd, int step, float grad_scale, bool state_fp32, cudaStream_t s);

// ---- gemm_tcgen05.cu -----------------------------------------------------------------------
// out[M,N] (+)= op(A) @ op(B), bf16 in / fp32 TMEM accumulate / bf16 out, row-major storage:
//   a_kmajor: A stored [M,K] (else [K,M]);  b_kmajor: B stored [N,K] (else [K,N]).
// lda/ldb/ldc are row strides in elements.  variant: 0 = auto, 1 = 1-CTA 128x256, 2 = 2-CTA 256x256
void gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, long long lda, long long ldb, long long ldc,
               bool a_kmajor, bool b_kmajor, bool accumulate, int variant, cudaStream_t s);

int gemm_max_active_clusters(int cg);
// tensor-parallel variants over symmetric buffers (see gemm_tcgen05.cu)
void gemm_bf16_dist(int mode, const void* const* a_srcs, const void* const* b_srcs, void* const* c_dsts, int M, int N,
                    int K, long long lda, long long ldb, long long ldc, bool b_kmajor, bool accumulate, int nranks,
                    int rank, int rows_per_peer, cudaStream_t s);

void gemm_bf16_bgather(const void* A, void* full_base, void* C, int M, int N, int K, long long lda, long long ldb,
                       long long ldc, bool b_kmajor, const void* const* shards, long long per_bytes, long long w_off,
                       long long w_bytes, uint32_t* counters, uint32_t target, int chunk_shift, uint32_t* const* pads,
                       int nranks, int rank, uint32_t bar_epoch, cudaStream_t s);
void gemm_bf16_ag(const void* const* a_bufs, const void* B, void* C, int M, int N, int K, long long ldb, long long ldc,
                  bool b_kmajor, int nranks, int rank, int rows_per_peer, uint32_t* flags, uint32_t ag_epoch,
                  uint32_t* const* pads, uint32_t bar_epoch, int n_comm, cudaStream_t s);

// ---- attention.cu --------------------------------------------------------------------------
// qkv: [B,S,nh+2*nkv,128] bf16 (q heads | k heads | v heads); o: [B,S,nh,128]; lse: [B,nh,S] fp32
void attn_fwd(const void* qkv, void* o, float* lse, int B, int S, int nh, int nkv, float scale, cudaStream_t s);
void attn_fwd2(const void* qkv, void* o, float* lse, int B, int S, int nh, int nkv, float scale, cudaStream_t s);
void attn_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, float* delta, float* dq_acc,
              void* dqkv, int B, int S, int nh, int nkv, float scale, int mode, cudaStream_t s);

}  // namespace dtg
