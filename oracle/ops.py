"""CPU oracle for the Long-VITA long-context hot path - single-operator restatements.

TEST INFRASTRUCTURE ONLY.  Nothing under `long-vita_b200/` imports this package; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may.

Each function restates, in plain PyTorch on the CPU, the arithmetic of one reference call site
(path:line relative to the Long-VITA repository).  Where the reference's arithmetic lives in an
un-vendored third-party package the docstring names it:
  * flash-attn 2 (`flash_attn_func`, v2 API, unpinned in requirements.txt) - attention;
  * transformers Qwen2 (`>=4.48.3`, requirements.txt:13) - RMSNorm / RoPE / MLP of the decoder;
  * Megatron-LM core_r0.7.0 @ 5f4c9ac9 and TransformerEngine - the mcore twins of the same ops.

PARITY PINNING: the reference repository ships no tests, golden vectors or fixtures (SURVEY.md
section 4), so these restatements are pinned against OUTPUTS OF THE REFERENCE ITSELF RUN HERE: its
whole `LongVITAForCausalLM.forward`, and InternViT + ResamplerProjector on their own, executed from
/root/reference (oracle/ref_loader.py, tests/golden/make_golden.py -> tests/golden/*.pt) and, for the
un-vendored pieces, against the installed third-party implementations the reference calls
(transformers' Qwen2 modules) - see tests/test_oracle_pinning.py.  Not pinned by reference outputs: the
flash-attn / TransformerEngine kernels (CUDA-only, un-vendored) - attention is pinned by its
definition and a live flash-attn 2.8 comparator on the GPU box - and the Megatron composition (Megatron is
an empty submodule), which shares every operator with the HF path.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
def attention(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    *,
    causal: bool,
    scale: Optional[float] = None,
    q_pos: Optional[torch.Tensor] = None,
    kv_pos: Optional[torch.Tensor] = None,
    head_chunk: int = 8,
    q_chunk: Optional[int] = None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """softmax(scale q k^T + mask) v in fp32 with its log-sum-exp.

    Follows `flash_attention_forward` (long_vita_megatron/core/transformer/dot_product_attention.py
    :294-394: scale = 1/sqrt(hn) :331-332, causal=True :378, GQA without pre-expansion :176-184)
    and the ViT branch (:312-329, causal=False), whose kernel is flash-attn 2 (un-vendored).
    q [b, sq, hq, d], k/v [b, sk, hkv, d] (any float dtype; math in fp32).  `q_pos` / `kv_pos` are
    global positions (int64 [sq] / [sk]); causal masks keys with kv_pos > q_pos.  Defaults give the
    bottom-right aligned mask of flash-attn >= 2.1.  Returns (out [b, sq, hq, d] fp32,
    lse [b, hq, sq] fp32).  `head_chunk` / `q_chunk` only bound the size of the score matrix held at once
    (rows of a softmax are independent; the arithmetic per row is unchanged)."""
    b, sq, hq, d = q.shape
    if q_chunk is not None and q_chunk < sq:
        if q_pos is None:
            q_pos = torch.arange(sq, dtype=torch.int64) + (k.shape[1] - sq)
        outs, lses = [], []
        for r0 in range(0, sq, q_chunk):
            o, l = attention(q[:, r0 : r0 + q_chunk], k, v, causal=causal, scale=scale, q_pos=q_pos[r0 : r0 + q_chunk],
                             kv_pos=kv_pos, head_chunk=head_chunk)
            outs.append(o)
            lses.append(l)
        return torch.cat(outs, dim=1), torch.cat(lses, dim=2)
    sk, hkv = k.shape[1], k.shape[2]
    g = hq // hkv
    scale = 1.0 / math.sqrt(d) if scale is None else scale
    if q_pos is None:
        q_pos = torch.arange(sq, dtype=torch.int64) + (sk - sq)
    if kv_pos is None:
        kv_pos = torch.arange(sk, dtype=torch.int64)
    qf = q.float().permute(0, 2, 1, 3)  # b h s d
    kf = k.float().permute(0, 2, 1, 3)
    vf = v.float().permute(0, 2, 1, 3)
    out = torch.empty((b, hq, sq, d), dtype=torch.float32)
    lse = torch.empty((b, hq, sq), dtype=torch.float32)
    mask = None
    if causal:
        mask = kv_pos[None, :] > q_pos[:, None]  # [sq, sk] True = hidden
    for h0 in range(0, hq, head_chunk):
        h1 = min(hq, h0 + head_chunk)
        kvh = torch.arange(h0, h1) // g
        s = torch.matmul(qf[:, h0:h1], kf[:, kvh].transpose(-1, -2)) * scale  # b h sq sk
        if mask is not None:
            s = s.masked_fill(mask, float("-inf"))
        l = torch.logsumexp(s, dim=-1)
        p = torch.exp(s - l.unsqueeze(-1))
        p = torch.nan_to_num(p, nan=0.0)  # fully masked rows
        out[:, h0:h1] = torch.matmul(p, vf[:, kvh])
        lse[:, h0:h1] = l
    return out.permute(0, 2, 1, 3).contiguous(), lse


def attention_grads(q, k, v, d_out, *, causal: bool, scale: Optional[float] = None, q_pos=None, kv_pos=None):
    """dq, dk, dv of `attention` by fp32 autograd over the same definition (the reference gets them
    from flash-attn 2's backward through autograd; no separate formula exists in the reference)."""
    qf, kf, vf = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    b, sq, hq, d = qf.shape
    sk, hkv = kf.shape[1], kf.shape[2]
    g = hq // hkv
    scale = 1.0 / math.sqrt(d) if scale is None else scale
    if q_pos is None:
        q_pos = torch.arange(sq, dtype=torch.int64) + (sk - sq)
    if kv_pos is None:
        kv_pos = torch.arange(sk, dtype=torch.int64)
    kk = kf.repeat_interleave(g, dim=2)
    vv = vf.repeat_interleave(g, dim=2)
    s = torch.einsum("bqhd,bkhd->bhqk", qf, kk) * scale
    if causal:
        s = s.masked_fill(kv_pos[None, :] > q_pos[:, None], float("-inf"))
    p = torch.nan_to_num(torch.softmax(s, dim=-1), nan=0.0)
    out = torch.einsum("bhqk,bkhd->bqhd", p, vv)
    out.backward(d_out.float())
    return qf.grad, kf.grad, vf.grad


def zigzag_positions(seq_len: int, cp: int, rank: int) -> torch.Tensor:
    """Global positions owned by `rank`: chunks {r, 2cp-1-r} of 2cp equal chunks
    (long_vita_megatron/training/utils.py:329-341, generation.py:517-539)."""
    c = seq_len // (2 * cp)
    return torch.cat([torch.arange(rank * c, (rank + 1) * c), torch.arange((2 * cp - 1 - rank) * c, (2 * cp - rank) * c)])


def zigzag_split(x: torch.Tensor, cp: int, rank: int, seq_dim: int = 1) -> torch.Tensor:
    """x.view(.., 2cp, S/2cp, ..)[[r, 2cp-1-r]] flattened back (training/utils.py:329-341)."""
    idx = zigzag_positions(x.shape[seq_dim], cp, rank)
    return x.index_select(seq_dim, idx)


def zigzag_unsplit(parts: Sequence[torch.Tensor], seq_dim: int = 1) -> torch.Tensor:
    """Inverse of zigzag_split over all ranks (the all-gather + re-order of
    long_vita_megatron/inference/text_generation/generation.py:542-566)."""
    cp = len(parts)
    full = torch.cat(list(parts), dim=seq_dim)
    seq_len = full.shape[seq_dim]
    order = torch.cat([zigzag_positions(seq_len, cp, r) for r in range(cp)])
    inv = torch.empty_like(order)
    inv[order] = torch.arange(seq_len)
    return full.index_select(seq_dim, inv)


def index_of_a_in_b(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Position in `b` of every element of `a` (each occurs exactly once in b);
    training/utils.py:347-350 computes the same through isin + double argsort."""
    sorted_b, perm = torch.sort(b)
    return perm[torch.searchsorted(sorted_b, a)]


def ring_attention_zigzag(q, k, v, cp: int, scale: Optional[float] = None):
    """Simulate the zig-zag ring schedule of TransformerEngine's AttnFuncWithCP (un-vendored; the
    schedule is described in SURVEY.md section 8e) in one process: every rank computes partial
    attention per ring step and merges by log-sum-exp.  q,k,v [b, S, h, d] full tensors.  Returns
    the list of per-rank outputs (fp32 [b, 2c, hq, d]) and LSEs."""
    S = q.shape[1]
    c = S // (2 * cp)
    outs, lses = [], []
    for r in range(cp):
        qpos = zigzag_positions(S, cp, r)
        ql = q.index_select(1, qpos)
        acc = None
        lse = None
        for step in range(cp):
            src = (r - step) % cp
            kpos = zigzag_positions(S, cp, src)
            kl, vl = k.index_select(1, kpos), v.index_select(1, kpos)
            if step == 0:
                o_blk, l_blk = attention(ql, kl, vl, causal=True, scale=scale, q_pos=qpos, kv_pos=kpos)
                rows = slice(0, 2 * c)
            elif step <= r:
                o_blk, l_blk = attention(ql, kl[:, :c], vl[:, :c], causal=False, scale=scale)
                rows = slice(0, 2 * c)
            else:
                o_blk, l_blk = attention(ql[:, c:], kl, vl, causal=False, scale=scale)
                rows = slice(c, 2 * c)
            if acc is None:
                acc, lse = o_blk.clone(), l_blk.clone()
            else:
                # out = out - sigmoid(lse_blk - lse) * (out - out_blk); lse = lse - logsigmoid(lse - lse_blk)
                lo, lb = lse[:, :, rows], l_blk
                w = torch.sigmoid(lb - lo).permute(0, 2, 1).unsqueeze(-1)
                acc[:, rows] = acc[:, rows] - w * (acc[:, rows] - o_blk)
                lse[:, :, rows] = lo - F.logsigmoid(lo - lb)
        outs.append(acc)
        lses.append(lse)
    return outs, lses


# ------------------------------------------------------------------------------------------------
# norms, rope, activations
# ------------------------------------------------------------------------------------------------
def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """x.float() * rsqrt(mean(x^2) + eps) -> cast to x.dtype -> * w.
    long_vita_megatron/core/transformer/custom_layers/transformer_engine.py:74-79; identical to
    transformers' Qwen2RMSNorm and to InternRMSNorm (modeling_intern_vit.py:33-44)."""
    xf = x.float()
    n = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return n.to(x.dtype) * w


def layernorm(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], eps: float = 1e-6) -> torch.Tensor:
    """torch.nn.LayerNorm as used at modeling_intern_vit.py:205-206 (eps = layer_norm_eps 1e-6)
    and resampler_projector.py:17 (eps default 1e-5 there - the caller passes it)."""
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), None if b is None else b.float(), eps).to(x.dtype)


def rope_inv_freq(dim: int, theta: float) -> torch.Tensor:
    """1 / theta^(2i/dim), fp32 (rotary_pos_embedding.py:70-78; HF Qwen2RotaryEmbedding default)."""
    return 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.int64).float() / dim))


def rope_tables(pos: torch.Tensor, inv_freq: torch.Tensor, dtype=torch.bfloat16):
    """cos/sin of cat(freqs, freqs), freqs = outer(pos, inv_freq) in fp32, cast to `dtype`
    (rotary_pos_embedding.py:95-106, 200-201)."""
    freqs = torch.outer(pos.float(), inv_freq.float())
    emb = torch.cat((freqs, freqs), dim=-1)
    return torch.cos(emb).to(dtype), torch.sin(emb).to(dtype)


def rope_apply(t: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """t * cos + rotate_half(t) * sin in t's dtype; t [n_tok, heads, dim], cos/sin [n_tok, dim].
    apply_rotary_pos_emb_bshd (rotary_pos_embedding.py:181-204) / HF apply_rotary_pos_emb."""
    half = t.shape[-1] // 2
    rot = torch.cat((-t[..., half:], t[..., :half]), dim=-1)
    return (t * cos[:, None, :]) + (rot * sin[:, None, :])


def swiglu(gate_up: torch.Tensor) -> torch.Tensor:
    """silu(gate) * up with fc1 = cat(gate, up) (tools/hf2mcore_long_vita.py:502-504; HF Qwen2MLP)."""
    inter = gate_up.shape[-1] // 2
    return F.silu(gate_up[..., :inter]) * gate_up[..., inter:]


def bias_gelu(x: torch.Tensor, bias: Optional[torch.Tensor], approximate: str = "none") -> torch.Tensor:
    """gelu(x + bias): exact erf for InternViT (pretrain_long_vita.py:206), tanh for SigLIP (:291)."""
    if bias is not None:
        x = x + bias
    return F.gelu(x, approximate=approximate)


def ls_residual(x, y, ls=None, bias=None):
    """x + (y + bias) * ls (modeling_intern_vit.py:224-226, intern_vit_model.py:63,77)."""
    if bias is not None:
        y = y + bias
    if ls is not None:
        y = y * ls
    return x + y


# ------------------------------------------------------------------------------------------------
# permutation / index operators (bit-exact)
# ------------------------------------------------------------------------------------------------
def pixel_shuffle_half(x: torch.Tensor) -> torch.Tensor:
    """x [n, w, h, c] -> [n, w/2, h/2, 4c]: out[n, w2, h2, (wi, hi, c)] = x[n, 2*w2+wi, 2*h2+hi, c].
    Same permutation as pixel_shuffle(scale_factor=0.5) at resampler_projector.py:36-46 /
    pretrain_long_vita.py:572-582 (verified against the imported reference function in
    tests/test_oracle_pinning.py)."""
    n, w, h, c = x.shape
    x = x.reshape(n, w // 2, 2, h // 2, 2, c)       # n w2 wi h2 hi c
    x = x.permute(0, 1, 3, 2, 4, 5)                 # n w2 h2 wi hi c
    return x.reshape(n, w // 2, h // 2, 4 * c).contiguous()


def embed_scatter(ids, table, feat=None, dst_idx=None, src_idx=None):
    """word_embeddings(ids) then rows overwritten by image features
    (language_model_embedding.py:102-131; modeling_long_vita.py:138-147). ids [n_tok] flat."""
    out = table[ids.view(-1)].clone()
    if feat is not None:
        f = feat.reshape(-1, table.shape[1])
        if src_idx is not None:
            f = f[src_idx.view(-1)]
        out[dst_idx.view(-1)] = f
    return out


def masked_linear_fwd(h: torch.Tensor, weight: torch.Tensor, logit_mask: torch.Tensor) -> torch.Tensor:
    """masked_select(h, mask) -> [M, b, c] @ W^T (layers.py:402-409). h [s, b, c], mask [b, s]."""
    s, b, c = h.shape
    sel = torch.masked_select(h, logit_mask.transpose(0, 1).unsqueeze(2)).reshape(-1, b, c)
    return torch.matmul(sel, weight.t())


def masked_linear_bwd(grad_out, h, weight, logit_mask):
    """dX = masked_scatter(zeros[s,b,c], dY W); dW = dY^T sel (layers.py:443-456, 512-520)."""
    s, b, c = h.shape
    m = logit_mask.transpose(0, 1).unsqueeze(2)
    gi = torch.matmul(grad_out, weight)
    gx = torch.zeros((s, b, c), dtype=gi.dtype).masked_scatter(m, gi)
    sel = torch.masked_select(h, m).reshape(-1, b, c)
    gw = torch.matmul(grad_out.reshape(-1, grad_out.shape[-1]).t(), sel.reshape(-1, c))
    return gx, gw


def patch_embed(images, conv_w, conv_b, cls, pos):
    """Conv2d(3, C, k=ps, s=ps) -> flatten -> cat cls -> + position embedding
    (modeling_intern_vit.py:96-108; the bicubic branch is the identity at the native 448)."""
    ps = conv_w.shape[-1]
    pe = F.conv2d(images, conv_w, conv_b, stride=ps)
    pe = pe.flatten(2).transpose(1, 2)
    cls_tok = cls.reshape(1, 1, -1).expand(pe.shape[0], 1, -1).to(pe.dtype)
    return torch.cat([cls_tok, pe], dim=1) + pos.reshape(1, -1, pe.shape[-1]).to(pe.dtype)
