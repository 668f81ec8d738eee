"""Op layer: every hot op of the Llama training step as an ``autograd.Function``.

On CUDA tensors each op calls a hand-written sm_100a kernel from the in-tree
extension (``csrc/``): tcgen05/TMEM/TMA GEMM (fwd / dgrad / wgrad), tcgen05
flash-attention, fused residual-add+RMSNorm, in-place RoPE on the fused qkv buffer,
SwiGLU, in-place softmax-cross-entropy, embedding gather / scatter-add and flat
AdamW.  On CPU tensors the same Functions run the reference math in
``ops/reference.py`` (chapter 01's CPU config and the gloo tests).

Replaces what the reference obtains implicitly from cuBLAS / SDPA / flash-attn /
ATen / Inductor.

Weight-gradient protocol: a parameter may carry ``_dtg_grad`` (a view into a flat,
possibly NVLink-symmetric, gradient buffer).  wgrad kernels then write (first use
after ``zero_grad``) or accumulate (later uses / micro-batches) straight into that
view and the Function returns ``None`` for the weight, so autograd never allocates
or copies a gradient.  ``parallel/flat.py`` owns those buffers.
"""
from __future__ import annotations

import math

import torch

from .. import _ext
from . import reference as ref

__all__ = [
    "linear", "fused_linear", "rms_norm", "add_rms_norm", "rope_qkv_", "attention_qkv", "swiglu", "cross_entropy",
    "embedding", "gemm", "ref",
]


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
# EXPERIMENTAL (DTG_WGRAD_STREAM=1, off by default): issue the weight-gradient GEMM of a linear layer on a side
# stream so that it runs next to the data-gradient GEMM of the same layer; both are persistent kernels, so the
# CTAs of one fill the SMs the other leaves idle in its last partial wave.  Engines call ``join_wgrad_stream()``
# before they publish a bucket's gradients.
_WGRAD_SIDE = {"enabled": bool(__import__("os").environ.get("DTG_WGRAD_STREAM")), "stream": None, "dirty": False}


def _wgrad_stream(device):
    st = _WGRAD_SIDE
    if st["stream"] is None:
        st["stream"] = torch.cuda.Stream(device=device)
    return st["stream"]


def join_wgrad_stream():
    """Make the current stream wait for weight gradients issued on the side stream (no-op when the feature is off)."""
    st = _WGRAD_SIDE
    if st["dirty"]:
        torch.cuda.current_stream().wait_stream(st["stream"])
        st["dirty"] = False


def _emit_weight_grad(param, compute_into, shape_like):
    """Route a weight gradient.

    ``compute_into(out, accumulate)`` must write the gradient into ``out`` (a tensor of
    the parameter's shape), adding to it when ``accumulate``.  Returns the tensor to hand
    back to autograd (``None`` when it went into the flat buffer).
    """
    buf = getattr(param, "_dtg_grad", None)
    if buf is not None:
        n = getattr(param, "_dtg_writes", 0)
        compute_into(buf, n > 0)
        param._dtg_writes = n + 1
        hook = getattr(param, "_dtg_ready_hook", None)
        if hook is not None:
            hook(param)
        return None
    out = torch.empty_like(shape_like)
    compute_into(out, False)
    return out


def gemm(a, b, out=None, trans_a=False, trans_b=False, accumulate=False):
    """out[M,N] (+)= op(a) @ op(b) on the tcgen05 GEMM (bf16 in, fp32 accumulate in TMEM).

    ``op(a)`` is ``a`` ([M,K]) or ``a.T`` when ``trans_a`` (a given as [K,M]);
    ``op(b)`` is ``b`` ([K,N]) or ``b.T`` when ``trans_b`` (b given as [N,K]).
    """
    M = a.shape[1] if trans_a else a.shape[0]
    N = b.shape[0] if trans_b else b.shape[1]
    if out is None:
        out = torch.empty(M, N, dtype=a.dtype, device=a.device)
        accumulate = False
    if _ext.use_cuda_kernel("gemm", a, b, out):
        _ext.load().gemm(a, b, out, trans_a, trans_b, accumulate)
    else:
        A = a.t() if trans_a else a
        Bm = b.t() if trans_b else b
        r = (A.float() @ Bm.float())
        if accumulate:
            out.add_(r.to(out.dtype))
        else:
            out.copy_(r.to(out.dtype))
    return out


# --------------------------------------------------------------------------------------
# linear
# --------------------------------------------------------------------------------------
class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, w_param):
        # x: [..., K], w: [N, K]; w_param is the nn.Parameter that owns the grad-buffer hooks
        x2 = x.reshape(-1, x.shape[-1])
        ctx.save_for_backward(x2, w)
        ctx.w_param = w_param if w_param is not None else w
        ctx.x_shape = x.shape
        gs = getattr(ctx.w_param, "_dtg_gather", None)   # FSDP: the weight may still have to be gathered
        if gs is not None and x2.is_cuda and gs.pending():
            y = gs.gemm(x2 if x2.is_contiguous() else x2.contiguous(), True)   # unshard fused into this GEMM
        else:
            y = gemm(x2, w, trans_b=True)
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            gs = getattr(ctx.w_param, "_dtg_gather", None)
            if gs is not None and dy2.is_cuda and gs.pending():
                dx = gs.gemm(dy2, False).view(ctx.x_shape)   # FSDP: re-gather the weight inside the dgrad GEMM
            else:
                dx = gemm(dy2, w).view(ctx.x_shape)  # [T,N] @ [N,K]
        dw = None
        if ctx.needs_input_grad[1] or getattr(ctx.w_param, "_dtg_grad", None) is not None:
            side = _WGRAD_SIDE["enabled"] and dy2.is_cuda and getattr(ctx.w_param, "_dtg_grad", None) is not None
            if side:
                cur, ss = torch.cuda.current_stream(), _wgrad_stream(dy2.device)
                ss.wait_stream(cur)               # dy2 / x2 were produced on the compute stream
                dy2.record_stream(ss)             # keep the caching allocator from recycling them too early
                x2.record_stream(ss)
                with torch.cuda.stream(ss):
                    dw = _emit_weight_grad(ctx.w_param,
                                           lambda out, acc: gemm(dy2, x2, out=out, trans_a=True, accumulate=acc), w)
                _WGRAD_SIDE["dirty"] = True
            else:
                dw = _emit_weight_grad(
                    ctx.w_param,
                    lambda out, acc: gemm(dy2, x2, out=out, trans_a=True, accumulate=acc),  # dy^T @ x
                    w,
                )
        return dx, dw, None


def linear(x, w, bias=None):
    """y = x @ w.T (+ bias).  bf16 CUDA tensors run the tcgen05 GEMM."""
    if bias is None and _ext.use_cuda_kernel("gemm", x, w) and x.dtype == torch.bfloat16:
        return _Linear.apply(x, w, w)
    return ref.linear(x, w, bias)


def fused_linear(x, w, owner=None):
    """``linear`` over a fused weight (q|k|v or gate|up).  ``owner`` is the
    ``models.llama.FusedWeight`` carrying the flat-gradient view, or None when ``w`` is an
    ordinary autograd tensor (e.g. a ``torch.cat`` of the individual parameters)."""
    if _ext.use_cuda_kernel("gemm", x, w) and x.dtype == torch.bfloat16:
        return _Linear.apply(x, w, owner if owner is not None else w)
    return ref.linear(x, w)


# --------------------------------------------------------------------------------------
# RMSNorm (+ fused residual add)
# --------------------------------------------------------------------------------------
def _norm_dw_into(dw32):
    def into(out, acc):
        if acc:
            out.add_(dw32.to(out.dtype))
        else:
            out.copy_(dw32)
    return into


class _RMSNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, eps):
        C = _ext.load()
        x2 = x.reshape(-1, x.shape[-1])
        y, rstd, _ = C.rmsnorm_fwd(x2, w, float(eps), None)
        ctx.save_for_backward(x2, w, rstd)
        ctx.shape = x.shape
        ctx.w_param = w
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        C = _ext.load()
        x2, w, rstd = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        dx, dw32 = C.rmsnorm_bwd(dy2, x2, w, rstd, None)
        dw = _emit_weight_grad(ctx.w_param, _norm_dw_into(dw32), w)
        return dx.view(ctx.shape), dw, None


class _AddRMSNorm(torch.autograd.Function):
    """(y, h) = (rmsnorm(x + r) * w, x + r) in one pass over the activations."""

    @staticmethod
    def forward(ctx, x, r, w, eps):
        C = _ext.load()
        x2 = x.reshape(-1, x.shape[-1])
        r2 = r.reshape(-1, r.shape[-1])
        y, rstd, h = C.rmsnorm_fwd(x2, w, float(eps), r2)
        ctx.save_for_backward(h, w, rstd)
        ctx.shape = x.shape
        ctx.w_param = w
        return y.view(x.shape), h.view(x.shape)

    @staticmethod
    def backward(ctx, dy, dh):
        C = _ext.load()
        h, w, rstd = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        dh2 = dh.reshape(-1, dh.shape[-1]).contiguous() if dh is not None else None
        # dx = rmsnorm_bwd(dy) + dh : the gradient of both x and r (h = x + r)
        dx, dw32 = C.rmsnorm_bwd(dy2, h, w, rstd, dh2)
        dw = _emit_weight_grad(ctx.w_param, _norm_dw_into(dw32), w)
        dx = dx.view(ctx.shape)
        return dx, dx, dw, None


def rms_norm(x, w, eps):
    if _ext.use_cuda_kernel("rmsnorm", x, w) and x.dtype == torch.bfloat16:
        return _RMSNorm.apply(x, w, eps)
    return ref.rms_norm(x, w, eps)


def add_rms_norm(x, residual, w, eps):
    """Fused ``h = x + residual; y = rmsnorm(h) * w`` -> (y, h)."""
    if _ext.use_cuda_kernel("rmsnorm", x, residual, w) and x.dtype == torch.bfloat16:
        return _AddRMSNorm.apply(x, residual, w, eps)
    return ref.add_rms_norm(x, residual, w, eps)


# --------------------------------------------------------------------------------------
# RoPE, applied in place on the q and k heads of the fused qkv activation
# --------------------------------------------------------------------------------------
class _RopeQKV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, cos, sin, n_rot_heads):
        # qkv: [B, S, n_total_heads, d]; the first n_rot_heads heads (q then k) are rotated
        C = _ext.load()
        ctx.save_for_backward(cos, sin)
        ctx.n_rot = n_rot_heads
        C.rope_inplace(qkv, cos, sin, n_rot_heads, False)
        # physically in place, but handed to autograd as a fresh tensor aliasing the same
        # storage: the producer GEMM never re-reads its output, so nothing observes the write
        return qkv.detach()

    @staticmethod
    def backward(ctx, dqkv):
        C = _ext.load()
        cos, sin = ctx.saved_tensors
        if not dqkv.is_contiguous():
            dqkv = dqkv.contiguous()
        C.rope_inplace(dqkv, cos, sin, ctx.n_rot, True)  # inverse rotation, in place on the grad
        return dqkv, None, None, None


def rope_qkv_(qkv, cos, sin, n_rot_heads):
    """Rotate heads [0, n_rot_heads) of ``qkv`` [B,S,heads,d] with cos/sin [S,d/2] or [B,S,d/2] (fp32)."""
    if _ext.use_cuda_kernel("rope", qkv) and qkv.dtype == torch.bfloat16:
        # views produced by a GEMM are fresh tensors, in-place is safe for autograd via mark_dirty
        return _RopeQKV.apply(qkv, cos.contiguous(), sin.contiguous(), n_rot_heads)
    rot = ref.rope_apply(qkv[:, :, :n_rot_heads], cos, sin)
    return torch.cat([rot, qkv[:, :, n_rot_heads:]], dim=2)


# --------------------------------------------------------------------------------------
# causal flash attention on the fused qkv buffer
# --------------------------------------------------------------------------------------
class _AttentionQKV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, nh, nkv, scale):
        C = _ext.load()
        o, lse = C.attn_fwd(qkv, nh, nkv, float(scale))
        ctx.save_for_backward(qkv, o, lse)
        ctx.meta = (nh, nkv, scale)
        return o

    @staticmethod
    def backward(ctx, do):
        C = _ext.load()
        qkv, o, lse = ctx.saved_tensors
        nh, nkv, scale = ctx.meta
        dqkv = C.attn_bwd(do.contiguous(), qkv, o, lse, nh, nkv, float(scale))
        return dqkv, None, None, None


def attention_qkv(qkv, nh, nkv, scale=None):
    """Causal self-attention. qkv: [B,S,nh+2*nkv,d] (q heads | k heads | v heads) -> [B,S,nh,d]."""
    d = qkv.shape[-1]
    scale = scale if scale is not None else 1.0 / math.sqrt(d)
    if _ext.use_cuda_kernel("attention", qkv) and qkv.dtype == torch.bfloat16 and d == 128 and qkv.shape[1] % 128 == 0:
        return _AttentionQKV.apply(qkv, nh, nkv, scale)
    q, k, v = qkv[:, :, :nh], qkv[:, :, nh:nh + nkv], qkv[:, :, nh + nkv:]
    if qkv.is_cuda:
        # head dims other than 128 (GPT-2 style / toy configs) and sequence lengths that are not a multiple of
        # the 128-row tile are outside the sm_100a kernel's scope; use the library kernel rather than the
        # O(S^2)-memory reference
        o = torch.nn.functional.scaled_dot_product_attention(
            q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True, scale=scale,
            enable_gqa=(nh != nkv))
        return o.transpose(1, 2)
    return ref.attention(q, k, v, causal=True, scale=scale)


# --------------------------------------------------------------------------------------
# SwiGLU on the fused [gate | up] activation
# --------------------------------------------------------------------------------------
class _SwiGLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gu):
        C = _ext.load()
        gu2 = gu.reshape(-1, gu.shape[-1])
        ctx.save_for_backward(gu2)
        ctx.shape = gu.shape
        return C.swiglu_fwd(gu2).view(*gu.shape[:-1], gu.shape[-1] // 2)

    @staticmethod
    def backward(ctx, dh):
        C = _ext.load()
        (gu2,) = ctx.saved_tensors
        dh2 = dh.reshape(-1, dh.shape[-1]).contiguous()
        return C.swiglu_bwd(dh2, gu2).view(ctx.shape)


def swiglu(gu):
    if _ext.use_cuda_kernel("swiglu", gu) and gu.dtype == torch.bfloat16:
        return _SwiGLU.apply(gu)
    return ref.swiglu(gu)


# --------------------------------------------------------------------------------------
# cross entropy (forward computes the loss and leaves dlogits in place of the logits)
# --------------------------------------------------------------------------------------
class _CrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets):
        C = _ext.load()
        # one pass: per-row logsumexp -> loss; second pass overwrites logits with
        # (softmax - onehot) / n_valid so backward is free of [T,V] temporaries.
        loss = C.cross_entropy_fwd_bwd(logits, targets)  # logits storage now holds dlogits
        ctx.save_for_backward(logits.detach())
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (dlogits,) = ctx.saved_tensors
        # dloss is a scalar; the common case (== 1) costs nothing
        C = _ext.load()
        C.scale_inplace(dlogits, dloss.reshape(1).float())
        return dlogits, None


def cross_entropy(logits, targets):
    """Mean CE over targets != -100. logits [T,V] (consumed in place on CUDA), targets [T] int64."""
    if _ext.use_cuda_kernel("cross_entropy", logits, targets) and logits.dtype == torch.bfloat16:
        return _CrossEntropy.apply(logits, targets.contiguous())
    return ref.cross_entropy(logits, targets)


# --------------------------------------------------------------------------------------
# embedding
# --------------------------------------------------------------------------------------
class _Embedding(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, w):
        C = _ext.load()
        ctx.save_for_backward(ids)
        ctx.w_param = w
        return C.embedding_fwd(ids.reshape(-1), w).view(*ids.shape, w.shape[1])

    @staticmethod
    def backward(ctx, dout):
        C = _ext.load()
        (ids,) = ctx.saved_tensors
        w = ctx.w_param
        d2 = dout.reshape(-1, dout.shape[-1]).contiguous()

        def into(out, acc):
            flat = ids.reshape(-1)
            if torch.are_deterministic_algorithms_enabled():
                # --deterministic: sorted runs summed in token order, no atomics (bit-identical run to run)
                if not acc:
                    out.zero_()
                srt, perm = torch.sort(flat, stable=True)
                C.embedding_bwd_sorted(d2, srt.contiguous(), perm.contiguous(), out, True)
                return
            if not acc:
                out.zero_()
            C.embedding_bwd(d2, flat, out)

        return None, _emit_weight_grad(w, into, w)


def embedding(ids, w):
    if _ext.use_cuda_kernel("embedding", ids, w) and w.dtype == torch.bfloat16:
        return _Embedding.apply(ids, w)
    return ref.embedding(ids, w)
