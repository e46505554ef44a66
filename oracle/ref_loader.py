"""Import the reference's own HF vision modules (InternVisionModel, ResamplerProjector,
pixel_shuffle) from /root/reference WITHOUT copying them - only possible in the build container
where the reference is mounted; used by tests/golden/make_golden.py to generate golden vectors
and by tests/test_oracle_pinning.py when the mount exists.

Two shims are needed: `timm.models.layers.DropPath` is not installed (identity at
drop rate 0, modeling_intern_vit.py:12,214-215) and `long_vita/__init__.py` imports the data
package (natsort / decord missing), so the model sub-package is loaded under a synthetic parent.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = "/root/reference"
PKG_DIR = os.path.join(REF_ROOT, "long_vita", "models", "long_vita_qwen2_intern")


def available() -> bool:
    return os.path.isdir(PKG_DIR)


def _stub_timm():
    if "timm.models.layers" in sys.modules:
        return
    import torch

    class DropPath(torch.nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            assert self.drop_prob == 0.0 or not self.training
            return x

    import importlib.machinery

    def mk(name):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
        m.__path__ = []
        return m

    timm, models, layers = mk("timm"), mk("timm.models"), mk("timm.models.layers")
    timm.__version__ = "0.0.0-stub"
    layers.DropPath = DropPath
    timm.models = models
    models.layers = layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


def load():
    """Returns a namespace with InternVisionModel, InternVisionConfig, ResamplerProjector,
    pixel_shuffle taken from the reference sources where they lie."""
    if not available():
        raise RuntimeError("/root/reference is not mounted on this machine")
    _stub_timm()
    parent = "lv_ref_pkg"
    if parent not in sys.modules:
        pkg = types.ModuleType(parent)
        pkg.__path__ = [PKG_DIR]
        sys.modules[parent] = pkg

    def imp(name):
        full = f"{parent}.{name}"
        if full in sys.modules:
            return sys.modules[full]
        spec = importlib.util.spec_from_file_location(full, os.path.join(PKG_DIR, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
        return mod

    cfg = imp("configuration_intern_vit")
    vit = imp("modeling_intern_vit")
    proj = imp("resampler_projector")
    ns = types.SimpleNamespace(
        InternVisionConfig=cfg.InternVisionConfig,
        InternVisionModel=vit.InternVisionModel,
        ResamplerProjector=proj.ResamplerProjector,
        pixel_shuffle=proj.pixel_shuffle,
    )
    return ns
