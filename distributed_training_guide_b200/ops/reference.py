"""Plain-PyTorch fp32 reference implementations of every op.

These serve two roles and are never a second GPU backend:
  * the CPU execution path (chapter 01's GPT-2 plumbing config, all ``gloo`` tests);
  * the numerics oracle the CUDA kernels are tested against (``tests/test_kernels_gpu.py``).

Semantics follow what the reference guide gets from ``transformers`` (SURVEY.md §3.2 /
K1-K9): RMSNorm with fp32 statistics, half-rotation RoPE, causal softmax attention
with GQA, SwiGLU, shifted-label mean cross-entropy with ``ignore_index=-100``.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def linear(x, w, bias=None):
    return F.linear(x, w, bias)


def rms_norm(x, w, eps):
    xf = x.float()
    rstd = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return ((xf * rstd).to(x.dtype) * w).to(x.dtype)


def add_rms_norm(x, residual, w, eps):
    """returns (normed, new_residual) with new_residual = x + residual."""
    h = x + residual
    return rms_norm(h, w, eps), h


def rope_tables(positions, head_dim, theta, scaling=None, dtype=torch.float32):
    """cos/sin tables [*positions.shape, head_dim//2] in fp32 (llama3 scaling supported)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32, device=positions.device) / head_dim))
    if scaling and scaling.get("rope_type", scaling.get("type")) == "llama3":
        factor = scaling["factor"]
        lo, hi = scaling["low_freq_factor"], scaling["high_freq_factor"]
        old = scaling["original_max_position_embeddings"]
        wavelen = 2 * math.pi / inv_freq
        smooth = (old / wavelen - lo) / (hi - lo)
        scaled = torch.where(wavelen > old / lo, inv_freq / factor, inv_freq)
        mid = (1 - smooth) * inv_freq / factor + smooth * inv_freq
        is_mid = (wavelen <= old / lo) & (wavelen >= old / hi)
        inv_freq = torch.where(is_mid, mid, scaled)
    ang = positions.to(torch.float32)[..., None] * inv_freq
    return ang.cos().to(dtype), ang.sin().to(dtype)


def rope_apply(x, cos, sin, inverse=False):
    """x: [B, S, nheads, d]; cos/sin: [S, d/2] or [B, S, d/2]. Half-rotation (HF Llama) layout."""
    d2 = x.shape[-1] // 2
    xf = x.float()
    x1, x2 = xf[..., :d2], xf[..., d2:]
    if cos.dim() == 2:
        c, s = cos[None, :, None, :], sin[None, :, None, :]
    else:
        c, s = cos[:, :, None, :], sin[:, :, None, :]
    if inverse:
        s = -s
    out = torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1)
    return out.to(x.dtype)


def attention(q, k, v, causal=True, scale=None):
    """q: [B, S, nh, d]; k, v: [B, S, nkv, d] -> [B, S, nh, d]. fp32 softmax."""
    B, S, nh, d = q.shape
    nkv = k.shape[2]
    scale = scale if scale is not None else 1.0 / math.sqrt(d)
    qf = q.float().permute(0, 2, 1, 3)
    kf = k.float().permute(0, 2, 1, 3)
    vf = v.float().permute(0, 2, 1, 3)
    if nkv != nh:
        rep = nh // nkv
        kf = kf.repeat_interleave(rep, dim=1)
        vf = vf.repeat_interleave(rep, dim=1)
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    if causal:
        mask = torch.ones(S, k.shape[1], dtype=torch.bool, device=q.device).tril()
        s = s.masked_fill(~mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, vf)
    return o.permute(0, 2, 1, 3).to(q.dtype)


def swiglu(gu):
    """gu: [..., 2*I] laid out as [gate | up] -> silu(gate) * up."""
    g, u = gu.chunk(2, dim=-1)
    return (F.silu(g.float()) * u.float()).to(gu.dtype)


def gelu_new(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x.pow(3))))


def shift_labels(labels):
    """HF causal-LM convention: token t predicts label t+1; last position ignored."""
    pad = torch.full_like(labels[..., :1], -100)
    return torch.cat([labels[..., 1:], pad], dim=-1)


def cross_entropy(logits, targets, ignore_index=-100):
    """logits [T, V] (any float dtype), targets [T] already shifted -> mean loss (fp32)."""
    return F.cross_entropy(logits.float(), targets, ignore_index=ignore_index, reduction="mean")


def embedding(ids, w):
    return F.embedding(ids, w)


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    """Single-tensor AdamW with fp32 math on (possibly bf16) storage, matching
    ``torch.optim.AdamW`` (decoupled decay, bias correction)."""
    pf, gf, mf, vf = p.float(), g.float() * grad_scale, m.float(), v.float()
    pf = pf * (1 - lr * weight_decay)
    mf = beta1 * mf + (1 - beta1) * gf
    vf = beta2 * vf + (1 - beta2) * gf * gf
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (vf.sqrt() / math.sqrt(bc2)) + eps
    pf = pf - (lr / bc1) * mf / denom
    p.copy_(pf)
    m.copy_(mf)
    v.copy_(vf)
