"""Model geometry of the Long-VITA hot path.

Numbers follow long_vita/models/long_vita_qwen2_intern/config_14B.json (LLM :34-56, ViT :2-32)
and the constants of tools/inference_long_vita.py:730-748 (256 visual tokens per frame).
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace


@dataclass(frozen=True)
class VisionConfig:
    hidden_size: int = 1024
    intermediate_size: int = 4096
    num_hidden_layers: int = 24
    num_attention_heads: int = 16
    image_size: int = 448
    patch_size: int = 14
    layer_norm_eps: float = 1e-6
    qkv_bias: bool = True
    hidden_act: str = "gelu"          # exact (erf) GELU
    initializer_factor: float = 1.0   # ls1 / ls2 init
    downsample_ratio: float = 0.5     # pixel shuffle
    pre_proj_ln_eps: float = 1e-5     # torch.nn.LayerNorm default (resampler_projector.py:17)

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def num_patches(self) -> int:
        return self.grid * self.grid

    @property
    def tokens_per_image(self) -> int:
        return int(self.num_patches * self.downsample_ratio * self.downsample_ratio)


@dataclass(frozen=True)
class LongVITAConfig:
    vocab_size: int = 152064
    hidden_size: int = 5120
    intermediate_size: int = 13824
    num_hidden_layers: int = 48
    num_attention_heads: int = 40
    num_key_value_heads: int = 8
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1_000_000.0
    initializer_range: float = 0.02
    visual: VisionConfig = field(default_factory=VisionConfig)

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @property
    def q_size(self) -> int:
        return self.num_attention_heads * self.head_dim

    @property
    def kv_size(self) -> int:
        return self.num_key_value_heads * self.head_dim

    @staticmethod
    def long_vita_14b() -> "LongVITAConfig":
        return LongVITAConfig()

    @staticmethod
    def tiny(layers: int = 2, vit_layers: int = 2) -> "LongVITAConfig":
        """Small geometry that keeps every structural property of the 14B model (GQA 5:1,
        head_dim 128, ViT head_dim 64, 448/14 patches) so tests exercise the same kernels."""
        return LongVITAConfig(
            vocab_size=2048,
            hidden_size=640,          # 5 q heads x 128
            intermediate_size=1024,
            num_hidden_layers=layers,
            num_attention_heads=5,
            num_key_value_heads=1,
            visual=replace(VisionConfig(), hidden_size=128, intermediate_size=256, num_hidden_layers=vit_layers,
                           num_attention_heads=2),
        )
