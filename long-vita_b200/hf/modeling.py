"""B200-native Long-VITA forward with the reference's HF `forward()` surface.

`LongVITAModel.forward` / `LongVITAForCausalLM.forward` keep the argument list of
long_vita/models/long_vita_qwen2_intern/modeling_long_vita.py:74-89 and :238-255, and the
composition of :90-221 (vision tower -> drop cls -> projector -> embedding gather + index_put ->
decoder layers -> final norm) and :309-311 (lm_head over the last `num_logits_to_keep` rows).
All arithmetic runs in liblvb200.so (long_vita_b200.ops); PyTorch holds the buffers.

`use_cache=True` returns a `kv_cache.KVCache` in `past_key_values` (pre-allocated, post-RoPE rows); a
later call with that cache and the new token(s) runs incremental decoding: one token -> the
flash-decoding composition `ops.attention_decode`, several tokens -> the fused kernel with the
bottom-right-aligned causal mask over cache + new rows (chunked prefill).  Images are encoded only when
the cache is empty (modeling_long_vita.py:90).  The Megatron serving path of the reference re-prefills
every token (long_vita_megatron/inference/text_generation/generation.py:127-135); SURVEY.md 8f-2.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from .. import ops
from ..config import LongVITAConfig
from ..kv_cache import KVCache


@dataclass
class CausalLMOutput:
    loss: Optional[torch.Tensor]
    logits: torch.Tensor
    past_key_values: Optional[object] = None
    hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    attentions: Optional[Tuple[torch.Tensor, ...]] = None

    def to_tuple(self):
        return tuple(v for v in (self.loss, self.logits, self.past_key_values, self.hidden_states) if v is not None)


@dataclass
class BaseOutput:
    last_hidden_state: torch.Tensor
    past_key_values: Optional[object] = None
    hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    attentions: Optional[Tuple[torch.Tensor, ...]] = None

    def to_tuple(self):
        return tuple(v for v in (self.last_hidden_state, self.past_key_values, self.hidden_states) if v is not None)

    def __getitem__(self, i):
        return self.to_tuple()[i]


class InternVisionModel:
    """InternViT-300M tower (modeling_intern_vit.py:298-363): patch-embed GEMM + cls + pos,
    24 x [LN -> MHA -> *ls1 + res -> LN -> fc1 -> GELU -> fc2 -> *ls2 + res]."""

    def __init__(self, cfg: LongVITAConfig, w: Dict[str, torch.Tensor], prefix: str = "model.vision_model."):
        self.cfg = cfg.visual
        v = self.cfg
        e = prefix + "embeddings."
        self.cls = w[e + "class_embedding"]
        self.pos = w[e + "position_embedding"]
        self.patch_w = ops.pad_patch_weight(w[e + "patch_embedding.weight"])
        self.patch_b = w[e + "patch_embedding.bias"]
        self.layers: List[Dict[str, torch.Tensor]] = []
        for i in range(v.num_hidden_layers):
            p = f"{prefix}encoder.layers.{i}."
            self.layers.append({k[len(p):]: t for k, t in w.items() if k.startswith(p)})

    def forward(self, pixel_values: torch.Tensor) -> torch.Tensor:
        v = self.cfg
        n = pixel_values.shape[0]
        C, H, D = v.hidden_size, v.num_attention_heads, v.head_dim
        x = ops.patch_embed(pixel_values, self.patch_w, self.patch_b, self.cls, self.pos, v.patch_size)
        S = x.shape[1]
        x = x.view(n * S, C)
        for L in self.layers:
            h = ops.layernorm(x, L["norm1.weight"], L["norm1.bias"], v.layer_norm_eps)
            qkv = ops.linear(h, L["attn.qkv.weight"], L.get("attn.qkv.bias")).view(n, S, 3, H, D)
            # 'b s (three h d)' (modeling_intern_vit.py:165): q/k/v are strided views, consumed in place
            att = ops.attention_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], causal=False, scale=D ** -0.5)
            o = ops.linear(att.view(n * S, C), L["attn.proj.weight"], L["attn.proj.bias"])
            x = ops.ls_residual(x, o, L["ls1"])
            h = ops.layernorm(x, L["norm2.weight"], L["norm2.bias"], v.layer_norm_eps)
            f = ops.linear(h, L["mlp.fc1.weight"], L["mlp.fc1.bias"], act="gelu")
            f = ops.linear(f, L["mlp.fc2.weight"], L["mlp.fc2.bias"])
            x = ops.ls_residual(x, f, L["ls2"])
        return x.view(n, S, C)

    __call__ = forward


class ResamplerProjector:
    """drop cls -> pixel-shuffle x0.5 -> LayerNorm(4C) -> Linear(4C, C) -> GELU -> Linear(C, hidden)
    (resampler_projector.py:26-34; the cls drop is modeling_long_vita.py:97, fused here)."""

    def __init__(self, cfg: LongVITAConfig, w: Dict[str, torch.Tensor], prefix: str = "model.vision_projection."):
        self.cfg = cfg
        self.ln_w = w[prefix + "pre_proj_layernorm.weight"]
        self.ln_b = w[prefix + "pre_proj_layernorm.bias"]
        self.w0 = w[prefix + "mlp.0.weight"]
        self.w2 = w[prefix + "mlp.2.weight"]

    def forward(self, vit_out: torch.Tensor, has_cls: bool = True) -> torch.Tensor:
        v = self.cfg.visual
        x = ops.pixel_shuffle(vit_out, v.grid, has_cls=has_cls)
        n, t, c4 = x.shape
        x = ops.layernorm(x.view(n * t, c4), self.ln_w, self.ln_b, v.pre_proj_ln_eps)
        x = ops.linear(x, self.w0, None, act="gelu")
        x = ops.linear(x, self.w2)
        return x.view(n, t, -1)

    __call__ = forward


class DecoderLayer:
    """Qwen2 decoder layer (transformers Qwen2DecoderLayer, instantiated by the reference at
    modeling_long_vita.py:57-72; mcore twin transformer_layer.py:173-257) with fused QKV and fused
    gate|up weights (layout of tools/hf2mcore_long_vita.py:486-504)."""

    def __init__(self, cfg: LongVITAConfig, w: Dict[str, torch.Tensor], i: int):
        p = f"model.layers.{i}."
        self.cfg = cfg
        self.wqkv = torch.cat([w[p + "self_attn.q_proj.weight"], w[p + "self_attn.k_proj.weight"],
                               w[p + "self_attn.v_proj.weight"]], dim=0).contiguous()
        self.bqkv = torch.cat([w[p + "self_attn.q_proj.bias"], w[p + "self_attn.k_proj.bias"],
                               w[p + "self_attn.v_proj.bias"]], dim=0).contiguous()
        self.wo = w[p + "self_attn.o_proj.weight"]
        # rows interleaved (gate_i, up_i): SwiGLU runs in the GEMM epilogue, [T, 2I] never touches HBM
        self.w_gate_up = ops.interleave_gate_up(w[p + "mlp.gate_proj.weight"], w[p + "mlp.up_proj.weight"])
        self.w_down = w[p + "mlp.down_proj.weight"]
        self.ln1 = w[p + "input_layernorm.weight"]
        self.ln2 = w[p + "post_attention_layernorm.weight"]

    def forward(self, x: torch.Tensor, delta: Optional[torch.Tensor], cos, sin, attn_kwargs, cache=None,
                layer_idx: int = 0, shard_merge=None, append: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        """x [T, H] residual stream, `delta` the previous layer's MLP output not yet added
        (the add is fused into this layer's first RMSNorm).  Returns (x, delta).  With `cache` the new
        K/V rows are appended to it and attention runs over cache + new rows."""
        cfg = self.cfg
        T = x.shape[0]
        hq, hkv, d = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
        if delta is None:
            h = ops.rmsnorm(x, self.ln1, cfg.rms_norm_eps)
        else:
            h, x = ops.rmsnorm(delta, self.ln1, cfg.rms_norm_eps, residual=x)
        qkv = ops.linear(h, self.wqkv, self.bqkv)                      # [T, (hq + 2 hkv) d]
        q = qkv[:, : hq * d].view(T, hq, d)
        k = qkv[:, hq * d : (hq + hkv) * d].view(T, hkv, d)
        v = qkv[:, (hq + hkv) * d :].view(T, hkv, d)
        ops.rope(q, cos, sin, out=q)
        ops.rope(k, cos, sin, out=k)
        if cache is None:
            att = ops.attention_fwd(q.unsqueeze(0), k.unsqueeze(0), v.unsqueeze(0), causal=True, **attn_kwargs)
        elif T == 1 and shard_merge is not None:
            # context-parallel decode: the cache is sharded over the ranks (any split of the keys works for a
            # query that sees them all); this rank attends over its shard, `shard_merge` combines the ranks'
            # (out, lse) pairs.  Only the rank that owns the new position appends its K/V row.
            if append:
                kc, vc, length = cache.append(layer_idx, k, v)
            else:
                kc, vc, length = cache.k[layer_idx], cache.v[layer_idx], len(cache)
            if length > 0:
                o_loc, lse_loc = ops.attention_decode(q[0], kc, vc, length, return_lse=True)
            else:
                o_loc = torch.zeros((hq, d), dtype=torch.bfloat16, device=x.device)
                lse_loc = torch.full((hq,), float("-inf"), dtype=torch.float32, device=x.device)
            att = shard_merge(o_loc, lse_loc)
        else:
            kc, vc, length = cache.append(layer_idx, k, v)
            if T == 1:
                att = ops.attention_decode(q[0], kc, vc, length)
            else:      # first prefill (empty cache) or a chunk of new tokens: causal, aligned to the last key
                att = ops.attention_fwd(q.unsqueeze(0), kc[:length].unsqueeze(0), vc[:length].unsqueeze(0), causal=True)
        o = ops.linear(att.reshape(T, hq * d), self.wo)
        h, x = ops.rmsnorm(o, self.ln2, cfg.rms_norm_eps, residual=x)
        a = ops.linear(h, self.w_gate_up, act="swiglu")
        return x, ops.linear(a, self.w_down)

    def forward_cp(self, x: torch.Tensor, delta: Optional[torch.Tensor], cos, sin, ctx, cache=None,
                   layer_idx: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """The same layer on this rank's zig-zag shard (`ctx` is the rank's cp.CPContext): the QKV GEMM
        writes straight into the peer-mapped buffer, RoPE runs in place there, and the fused kernel
        pulls the other ranks' K/V itself (lv_attn_cp_fwd)."""
        cfg = self.cfg
        T = x.shape[0]
        hq, hkv, d = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
        if delta is None:
            h = ops.rmsnorm(x, self.ln1, cfg.rms_norm_eps)
        else:
            h, x = ops.rmsnorm(delta, self.ln1, cfg.rms_norm_eps, residual=x)
        qkv = ops.linear(h, self.wqkv, self.bqkv, out=ctx.qkv_buffer()).view(T, -1)
        q = qkv[:, : hq * d].view(T, hq, d)
        k = qkv[:, hq * d : (hq + hkv) * d].view(T, hkv, d)
        ops.rope(q, cos, sin, out=q)
        ops.rope(k, cos, sin, out=k)
        if cache is not None:      # this rank's zig-zag rows become its shard of the K/V cache
            cache.append(layer_idx, k, qkv[:, (hq + hkv) * d :].view(T, hkv, d))
        o = ops.linear(ctx.attention(), self.wo)
        h, x = ops.rmsnorm(o, self.ln2, cfg.rms_norm_eps, residual=x)
        a = ops.linear(h, self.w_gate_up, act="swiglu")
        return x, ops.linear(a, self.w_down)


def _map_tensors(obj, fn, seen=None):
    """Apply `fn` to every tensor reachable from the attributes of our plain weight holders (in place)."""
    seen = set() if seen is None else seen
    if id(obj) in seen:
        return
    seen.add(id(obj))
    items = obj.items() if isinstance(obj, dict) else vars(obj).items()
    for name, val in list(items):
        if isinstance(val, torch.Tensor):
            new = fn(val)
            if isinstance(obj, dict):
                obj[name] = new
            else:
                object.__setattr__(obj, name, new)
        elif isinstance(val, (list, tuple)):
            for i, e in enumerate(val):
                if isinstance(e, torch.Tensor):
                    val[i] = fn(e)
                elif isinstance(e, (dict, InternVisionModel, ResamplerProjector, DecoderLayer)):
                    _map_tensors(e, fn, seen)
        elif isinstance(val, (dict, InternVisionModel, ResamplerProjector, DecoderLayer)):
            _map_tensors(val, fn, seen)


class _WeightHolderModule(torch.nn.Module):
    """nn.Module whose weights live in plain tensors (fused / re-laid-out once at load for the kernels) rather than in
    nn.Parameters: `.to()` / `.cuda()` / `.eval()` / hooks work as for any module, `state_dict()` is rebuilt in the
    reference's HF names by the subclasses."""

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        _map_tensors(self, lambda t: fn(t) if t.is_floating_point() or t.dtype in (torch.int64, torch.int32) else t)
        return self


class LongVITAModel(_WeightHolderModule):
    def __init__(self, cfg: LongVITAConfig, weights: Dict[str, torch.Tensor]):
        super().__init__()
        self.config = cfg
        self.embed_tokens = weights["model.embed_tokens.weight"]
        self.norm_w = weights["model.norm.weight"]
        self.layers = [DecoderLayer(cfg, weights, i) for i in range(cfg.num_hidden_layers)]
        has_vit = "model.vision_model.embeddings.class_embedding" in weights
        self.vision_model = InternVisionModel(cfg, weights) if has_vit else None
        self.vision_projection = ResamplerProjector(cfg, weights) if has_vit else None
        self.inv_freq = (1.0 / (cfg.rope_theta ** (torch.arange(0, cfg.head_dim, 2, dtype=torch.int64).float()
                                                    / cfg.head_dim))).to(self.embed_tokens.device)
        self.vision_chunk = 256  # frames per ViT pass (pretrain_long_vita.py:522-533)
        self.default_new_tokens = 1024   # head-room of a cache created by use_cache=True (max_cache_len= overrides)

    def encode_images(self, images: torch.Tensor) -> torch.Tensor:
        feats = []
        for i in range(0, images.shape[0], self.vision_chunk):
            vit = self.vision_model(images[i : i + self.vision_chunk])
            feats.append(self.vision_projection(vit, has_cls=True))
        return feats[0] if len(feats) == 1 else torch.cat(feats, dim=0)

    def forward(
        self,
        input_ids: Optional[torch.Tensor] = None,
        attention_mask: Optional[torch.Tensor] = None,
        images: Optional[torch.Tensor] = None,
        image_indices: Optional[torch.Tensor] = None,
        position_ids: Optional[torch.Tensor] = None,
        past_key_values=None,
        inputs_embeds: Optional[torch.Tensor] = None,
        use_cache: Optional[bool] = None,
        output_attentions: Optional[bool] = None,
        output_hidden_states: Optional[bool] = None,
        return_dict: Optional[bool] = None,
        cache_position: Optional[torch.Tensor] = None,
        **flash_attn_kwargs,
    ):
        cfg = self.config
        if (input_ids is None) ^ (inputs_embeds is not None):
            raise ValueError("You must specify exactly one of input_ids or inputs_embeds")
        if past_key_values is not None and not isinstance(past_key_values, KVCache):
            if len(past_key_values) != 0:
                raise NotImplementedError("past_key_values must be the KVCache returned by a use_cache=True call")
            past_key_values = None
        if output_attentions:
            raise NotImplementedError("attention probabilities are never materialised by the fused kernel")
        if attention_mask is not None and not bool(attention_mask.to(torch.bool).all()):
            raise NotImplementedError("padding masks are not supported; pass unpadded sequences (batch 1)")

        past_len = 0 if past_key_values is None else past_key_values.get_seq_length()
        image_embeds = None
        if images is not None and past_len == 0:                  # modeling_long_vita.py:90
            image_embeds = self.encode_images(images)
            assert image_embeds.shape[0] == len(images)

        if inputs_embeds is None:
            b, s = input_ids.shape
            if b != 1:
                raise NotImplementedError("batch size 1 (the reference's long-context configuration)")
            if image_embeds is not None:
                idx_b, idx_s = image_indices.to(input_ids.device).unbind(dim=0)
                dst = (idx_b.reshape(-1) * s + idx_s.reshape(-1)).to(torch.int64)
                x = ops.embed_scatter(input_ids, self.embed_tokens, image_embeds, dst)
            else:
                x = ops.embed_scatter(input_ids, self.embed_tokens)
        else:
            b, s, _ = inputs_embeds.shape
            if b != 1:
                raise NotImplementedError("batch size 1")
            x = inputs_embeds.reshape(s, -1).contiguous()

        cache = past_key_values
        if use_cache and cache is None:
            capacity = int(flash_attn_kwargs.pop("max_cache_len", 0)) or s + self.default_new_tokens
            cache = KVCache(len(self.layers), capacity, cfg.num_key_value_heads, cfg.head_dim, x.device)
        flash_attn_kwargs.pop("max_cache_len", None)
        if cache_position is None:
            cache_position = torch.arange(past_len, past_len + s, device=x.device)     # modeling_long_vita.py:153-158
        if position_ids is None:
            position_ids = cache_position.unsqueeze(0)
        cos, sin = ops.rope_table(position_ids.reshape(-1).to(torch.int64), self.inv_freq)

        all_hidden = () if output_hidden_states else None
        delta = None
        for li, layer in enumerate(self.layers):
            if output_hidden_states:
                all_hidden += ((x if delta is None else x + delta).view(1, s, -1),)
            x, delta = layer.forward(x, delta, cos, sin, {}, cache, li)
        if cache is not None:
            cache.commit()
        if delta is None:
            h = ops.rmsnorm(x, self.norm_w, cfg.rms_norm_eps)
        else:
            h, _ = ops.rmsnorm(delta, self.norm_w, cfg.rms_norm_eps, residual=x)
        h = h.view(1, s, -1)
        if output_hidden_states:
            all_hidden += (h,)
        out = BaseOutput(last_hidden_state=h, past_key_values=cache, hidden_states=all_hidden)
        return out if (return_dict is None or return_dict) else out.to_tuple()

    def hf_state_dict(self, prefix: str = "model.") -> Dict[str, torch.Tensor]:
        """The weights under the reference's HF parameter names (views of the fused buffers; the padded patch-embed
        operand is cut back to the [C, 3, ps, ps] conv weight)."""
        sd = {prefix + "embed_tokens.weight": self.embed_tokens, prefix + "norm.weight": self.norm_w}
        cfg = self.config
        for i, L in enumerate(self.layers):
            p = f"{prefix}layers.{i}."
            q, kv = cfg.q_size, cfg.kv_size
            sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"] = L.wqkv[:q], L.bqkv[:q]
            sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.k_proj.bias"] = L.wqkv[q : q + kv], L.bqkv[q : q + kv]
            sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"] = L.wqkv[q + kv :], L.bqkv[q + kv :]
            sd[p + "self_attn.o_proj.weight"] = L.wo
            sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"] = L.w_gate_up[0::2], L.w_gate_up[1::2]
            sd[p + "mlp.down_proj.weight"] = L.w_down
            sd[p + "input_layernorm.weight"], sd[p + "post_attention_layernorm.weight"] = L.ln1, L.ln2
        if self.vision_model is not None:
            v, vm = cfg.visual, self.vision_model
            e = prefix + "vision_model.embeddings."
            sd[e + "class_embedding"], sd[e + "position_embedding"] = vm.cls, vm.pos
            k = 3 * v.patch_size * v.patch_size
            sd[e + "patch_embedding.weight"] = vm.patch_w[:, :k].reshape(-1, 3, v.patch_size, v.patch_size)
            sd[e + "patch_embedding.bias"] = vm.patch_b
            for i, L in enumerate(vm.layers):
                for name, t in L.items():
                    sd[f"{prefix}vision_model.encoder.layers.{i}.{name}"] = t
            pj = self.vision_projection
            q = prefix + "vision_projection."
            sd[q + "pre_proj_layernorm.weight"], sd[q + "pre_proj_layernorm.bias"] = pj.ln_w, pj.ln_b
            sd[q + "mlp.0.weight"], sd[q + "mlp.2.weight"] = pj.w0, pj.w2
        return sd


class GenerationDefaults:
    """The fields of transformers' GenerationConfig this build reads; `model.generation_config` may be replaced by a
    real GenerationConfig (tools/inference_long_vita.py:819-826 assigns one and then sets these attributes)."""

    def __init__(self):
        self.max_new_tokens = 1024
        self.do_sample = False
        self.use_cache = True
        self.eos_token_id = None
        self.pad_token_id = None


class LongVITAForCausalLM(_WeightHolderModule):
    def __init__(self, cfg: LongVITAConfig, weights: Dict[str, torch.Tensor]):
        super().__init__()
        self.config = cfg
        self.model = LongVITAModel(cfg, weights)
        self.lm_head = weights["lm_head.weight"]
        self.generation_config = GenerationDefaults()

    @property
    def device(self) -> torch.device:
        return self.lm_head.device

    @property
    def dtype(self) -> torch.dtype:
        return self.lm_head.dtype

    def state_dict(self, *args, **kwargs) -> Dict[str, torch.Tensor]:
        sd = self.model.hf_state_dict("model.")
        sd["lm_head.weight"] = self.lm_head
        return sd

    @classmethod
    def from_pretrained(cls, model_path: str, torch_dtype=torch.bfloat16, device_map=None, attn_implementation=None,
                        trust_remote_code: bool = True, **kwargs):
        """Load a Long-VITA HF checkpoint directory (config.json + *.safetensors shards), the call
        tools/inference_long_vita.py:811-817 makes through AutoModelForCausalLM.  Only bfloat16 runs on the fused
        path; `attn_implementation` is accepted and ignored (the fused kernel IS the attention implementation);
        `device_map` "auto" / None -> the current CUDA device."""
        import glob
        import json
        import os

        from safetensors import safe_open

        if torch_dtype not in (torch.bfloat16, "bfloat16", None, "auto"):
            raise NotImplementedError("the fused path computes in bfloat16 (the reference's torch_dtype)")
        c = json.load(open(os.path.join(model_path, "config.json")))
        vc = c.get("visual", c.get("vision_config", {})) or {}
        from dataclasses import fields, replace

        from ..config import VisionConfig

        def pick(dc, src):
            names = {f.name for f in fields(dc)}
            return {k: v for k, v in src.items() if k in names and not isinstance(v, dict)}

        cfg = replace(LongVITAConfig(**pick(LongVITAConfig, c)), visual=VisionConfig(**pick(VisionConfig, vc)))
        dev = torch.device("cuda", torch.cuda.current_device()) if device_map in (None, "auto") else torch.device(device_map)
        w: Dict[str, torch.Tensor] = {}
        shards = sorted(glob.glob(os.path.join(model_path, "*.safetensors")))
        if not shards:
            raise FileNotFoundError(f"no *.safetensors under {model_path}")
        for sh in shards:
            with safe_open(sh, framework="pt", device=str(dev)) as f:
                for name in f.keys():
                    w[name] = f.get_tensor(name).to(torch.bfloat16)
        return cls(cfg, w)

    @classmethod
    def from_synthetic(cls, cfg: LongVITAConfig, seed: int = 1234, device="cuda", perturb: bool = False,
                       num_layers: Optional[int] = None):
        """Random-init model materialised layer by layer on `device` (14B bf16 = 29.5 GB never
        exists twice): every decoder layer's q/k/v and gate/up matrices are fused as they are drawn."""
        from ..weights import global_weights, llm_layer_weights, vit_layer_weights

        w = global_weights(cfg, seed, device, torch.bfloat16, perturb)
        for i in range(cfg.visual.num_hidden_layers):
            w.update(vit_layer_weights(cfg, i, seed, device, torch.bfloat16, perturb))
        n_layers = cfg.num_hidden_layers if num_layers is None else num_layers
        self = cls.__new__(cls)
        torch.nn.Module.__init__(self)
        self.config = cfg
        self.lm_head = w["lm_head.weight"]
        self.generation_config = GenerationDefaults()
        model = LongVITAModel.__new__(LongVITAModel)
        torch.nn.Module.__init__(model)
        model.config = cfg
        model.embed_tokens = w["model.embed_tokens.weight"]
        model.norm_w = w["model.norm.weight"]
        model.vision_model = InternVisionModel(cfg, w)
        model.vision_projection = ResamplerProjector(cfg, w)
        model.inv_freq = (1.0 / (cfg.rope_theta ** (torch.arange(0, cfg.head_dim, 2, dtype=torch.int64).float()
                                                     / cfg.head_dim))).to(device)
        model.vision_chunk = 256
        model.default_new_tokens = 1024
        model.layers = []
        for i in range(n_layers):
            lw = llm_layer_weights(cfg, i, seed, device, torch.bfloat16, perturb)
            model.layers.append(DecoderLayer(cfg, lw, i))
            del lw
        self.model = model
        return self

    def forward(
        self,
        input_ids: Optional[torch.Tensor] = None,
        attention_mask: Optional[torch.Tensor] = None,
        images: Optional[torch.Tensor] = None,
        image_indices: Optional[torch.Tensor] = None,
        position_ids: Optional[torch.Tensor] = None,
        past_key_values=None,
        inputs_embeds: Optional[torch.Tensor] = None,
        labels: Optional[torch.Tensor] = None,
        use_cache: Optional[bool] = None,
        output_attentions: Optional[bool] = None,
        output_hidden_states: Optional[bool] = None,
        return_dict: Optional[bool] = None,
        cache_position: Optional[torch.Tensor] = None,
        num_logits_to_keep: int = 0,
        **kwargs,
    ):
        outputs = self.model(
            input_ids=input_ids, attention_mask=attention_mask, images=images, image_indices=image_indices,
            position_ids=position_ids, past_key_values=past_key_values, inputs_embeds=inputs_embeds,
            use_cache=use_cache, output_attentions=output_attentions, output_hidden_states=output_hidden_states,
            return_dict=True, cache_position=cache_position, **kwargs,
        )
        hidden = outputs.last_hidden_state  # [1, s, H]
        # hidden_states[:, -num_logits_to_keep:, :] (modeling_long_vita.py:311); 0 keeps every row
        sel = hidden[:, -num_logits_to_keep:, :] if num_logits_to_keep else hidden
        rows = sel.reshape(-1, sel.shape[-1])
        logits = ops.linear(rows, self.lm_head).view(1, sel.shape[1], -1)
        loss = None
        if labels is not None:
            # transformers' loss_function (ForCausalLMLoss): labels shifted by one, mean of the per-token CE in fp32 over
            # the labels != -100.  Per-token CE from the chunked LM-head + CE kernels (ops.lm_head_ce_fwd): same bf16
            # logits, fp32 loss, and no fp32 copy of the [s, vocab] logits.
            tgt = labels[:, -logits.shape[1]:][:, 1:].reshape(-1).to(rows.device)
            per_tok, _ = ops.lm_head_ce_fwd(rows[:-1], self.lm_head, torch.where(tgt == -100, torch.full_like(tgt, -1), tgt))
            n_valid = (tgt != -100).sum().clamp_min(1)
            loss = per_tok.sum() / n_valid
        out = CausalLMOutput(loss=loss, logits=logits, past_key_values=outputs.past_key_values,
                             hidden_states=outputs.hidden_states)
        return out if (return_dict is None or return_dict) else out.to_tuple()

    @torch.no_grad()
    def generate(self, inputs: Optional[torch.Tensor] = None, images: Optional[torch.Tensor] = None,
                 image_indices: Optional[torch.Tensor] = None, generation_config=None, input_ids: Optional[torch.Tensor] = None,
                 max_new_tokens: Optional[int] = None, eos_token_id=None, do_sample: Optional[bool] = None, **kwargs):
        """`model.generate(inputs=, images=, image_indices=)` as tools/inference_long_vita.py:868 calls it: greedy
        decoding (the script sets do_sample = False) with the K/V cache; stops at `eos_token_id` (int or list) or after
        `max_new_tokens`; returns prompt + generated ids [1, s + n] like transformers' GenerationMixin.  Inputs on the
        host are moved to the model's device (the script leaves them on the CPU under device_map="auto")."""
        gc = generation_config if generation_config is not None else self.generation_config
        ids = inputs if inputs is not None else input_ids
        if ids is None:
            raise ValueError("generate() needs `inputs` (token ids [1, s])")
        sample = getattr(gc, "do_sample", False) if do_sample is None else do_sample
        if sample or kwargs.get("num_beams", 1) not in (None, 1):
            raise NotImplementedError("greedy decoding only (the reference's inference script: do_sample = False)")
        n_new = max_new_tokens if max_new_tokens is not None else (getattr(gc, "max_new_tokens", None) or 1024)
        eos = eos_token_id if eos_token_id is not None else getattr(gc, "eos_token_id", None)
        eos = set() if eos is None else ({int(eos)} if isinstance(eos, int) else {int(e) for e in eos})
        dev = self.device
        ids = ids.to(dev)
        if images is not None:
            images = images.to(dev, dtype=torch.bfloat16)
        if image_indices is not None:
            image_indices = image_indices.to(dev)
        new = self.generate_greedy(ids, images, image_indices, max_new_tokens=int(n_new), eos_token_ids=eos)
        return torch.cat([ids, new], dim=1)

    @torch.no_grad()
    def generate_greedy(self, input_ids: torch.Tensor, images: Optional[torch.Tensor] = None,
                        image_indices: Optional[torch.Tensor] = None, max_new_tokens: int = 16,
                        eos_token_id: Optional[int] = None, eos_token_ids=None) -> torch.Tensor:
        """Greedy decoding with the K/V cache: one prefill, then one forward per token over a single new
        row (the reference's Megatron loop feeds the whole sequence again for every token,
        generation.py:127-135).  Returns the generated ids [1, n]."""
        s = input_ids.shape[1]
        stop = set(eos_token_ids or ())
        if eos_token_id is not None:
            stop.add(int(eos_token_id))
        out = self.forward(input_ids=input_ids, images=images, image_indices=image_indices, use_cache=True,
                           num_logits_to_keep=1, max_cache_len=s + max_new_tokens)
        cache = out.past_key_values
        new = []
        tok = out.logits[0, -1].float().argmax().view(1, 1)
        for _ in range(max_new_tokens):
            new.append(tok)
            if stop and int(tok) in stop:
                break
            if len(new) == max_new_tokens:
                break
            out = self.forward(input_ids=tok, past_key_values=cache, use_cache=True, num_logits_to_keep=1)
            tok = out.logits[0, -1].float().argmax().view(1, 1)
        return torch.cat(new, dim=1)
