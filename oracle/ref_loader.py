"""Import the reference's own HF modules (InternVisionModel, ResamplerProjector, pixel_shuffle, and - through
`load_long_vita()` - the whole LongVITAForCausalLM) from /root/reference WITHOUT copying them - only possible in the build container
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
        _imp=imp,
    )
    return ns


def load_long_vita():
    """The reference's whole-model classes `LongVITAConfig`, `LongVITAModel`, `LongVITAForCausalLM`
    (modeling_long_vita.py, executed from /root/reference where it lies - nothing is copied).

    The file targets transformers >= 4.48.3 (requirements.txt:13) and no longer runs as-is under the
    installed 5.5.0.  Three API shims, none of which touches the reference's arithmetic:
      1. `transformers.utils.LossKwargs` (imported at :24, used only as a typing base at :224) was
         removed -> an empty TypedDict is put back;
      2. `Qwen2Model._update_causal_mask` (called at :162) was removed -> replaced by the explicit
         additive causal mask [1, 1, s, s] it used to return for eager / sdpa attention;
      3. `Qwen2DecoderLayer.forward` now returns a tensor, the reference indexes `layer_outputs[0]`
         (:204) as under 4.48, and the cache keyword was renamed `past_key_value` -> `past_key_values`
         (:196) -> each layer's forward is wrapped to return a 1-tuple and to accept the old keyword
         (`wrap_decoder_layers(model)`, call it after constructing the model).
    With these the unmodified reference forward runs on CPU (`attn_implementation="eager"`,
    `use_flash_attn=False` in the visual config)."""
    import torch
    import transformers.utils as TU
    from typing import TypedDict

    ns = load()
    if not hasattr(TU, "LossKwargs"):
        TU.LossKwargs = TypedDict("LossKwargs", {}, total=False)
    cfg = ns._imp("configuration_long_vita")
    mod = ns._imp("modeling_long_vita")

    def _update_causal_mask(self, attention_mask, input_tensor, cache_position, past_key_values, output_attentions=False):
        s = input_tensor.shape[1]
        past = past_key_values.get_seq_length() if past_key_values is not None else 0
        m = torch.full((s, past + s), torch.finfo(input_tensor.dtype).min, dtype=input_tensor.dtype, device=input_tensor.device)
        return torch.triu(m, diagonal=past + 1)[None, None]          # key j hidden from new row i iff j > past + i

    if not hasattr(mod.LongVITAModel, "_update_causal_mask"):
        mod.LongVITAModel._update_causal_mask = _update_causal_mask

    def wrap_decoder_layers(model):
        for layer in model.model.layers:
            orig = layer.forward

            def fwd(*a, _orig=orig, **k):
                if "past_key_value" in k:                 # 4.48 keyword; 5.x renamed it (and would swallow the old one)
                    k["past_key_values"] = k.pop("past_key_value")
                out = _orig(*a, **k)
                return out if isinstance(out, tuple) else (out,)

            layer.forward = fwd
        return model

    ns.LongVITAConfig = cfg.LongVITAConfig
    ns.LongVITAModel = mod.LongVITAModel
    ns.LongVITAForCausalLM = mod.LongVITAForCausalLM
    ns.wrap_decoder_layers = wrap_decoder_layers
    return ns


def build_reference_long_vita(cfg, state_dict):
    """Instantiate the reference `LongVITAForCausalLM` for our `LongVITAConfig` geometry `cfg` and load
    `state_dict` (HF names - the names `weights.synthetic_state_dict` produces are the reference's)."""
    ns = load_long_vita()
    v = cfg.visual
    visual = dict(attention_dropout=0.0, drop_path_rate=0.0, dropout=0.0, hidden_act="gelu", hidden_size=v.hidden_size,
                  image_size=v.image_size, initializer_factor=1.0, initializer_range=0.02,
                  intermediate_size=v.intermediate_size, layer_norm_eps=v.layer_norm_eps, norm_type="layer_norm",
                  num_attention_heads=v.num_attention_heads, num_channels=3, num_hidden_layers=v.num_hidden_layers,
                  patch_size=v.patch_size, qk_normalization=False, qkv_bias=True, use_flash_attn=False)
    rc = ns.LongVITAConfig(visual=visual, attention_dropout=0.0, hidden_act="silu", hidden_size=cfg.hidden_size,
                           initializer_range=cfg.initializer_range, intermediate_size=cfg.intermediate_size,
                           max_position_embeddings=1 << 20, num_attention_heads=cfg.num_attention_heads,
                           num_hidden_layers=cfg.num_hidden_layers, num_key_value_heads=cfg.num_key_value_heads,
                           rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta, tie_word_embeddings=False,
                           use_cache=False, use_sliding_window=False, vocab_size=cfg.vocab_size,
                           attn_implementation="eager")
    model = ns.LongVITAForCausalLM(rc).eval()
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return ns.wrap_decoder_layers(model)


# ------------------------------------------------------------------------------------------------
# Megatron-side helpers of the reference that are plain torch once their imports are satisfied
# ------------------------------------------------------------------------------------------------
def load_megatron_training_utils(cp_size: int, cp_rank: int, seq_length: int):
    """The reference's `long_vita_megatron/training/utils.py` (get_batch_on_this_cp_rank :252-343,
    index_of_a_in_b :347-350), executed from /root/reference.  Megatron-LM is an empty submodule there, so the
    module's imports are satisfied by empty stand-ins (`megatron.training.get_args` returns a namespace with
    the three fields the function reads, `mpu.get_context_parallel_rank` returns `cp_rank`).  Returns
    (module, cpu_placement) where `cpu_placement()` is a context manager under which the function's
    `device='cuda'` / `.cuda()` / `pin_memory=True` placements resolve to the CPU - placement only, no
    arithmetic is touched."""
    import contextlib
    import importlib.machinery

    import torch

    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("/root/reference is not mounted on this machine")
    args = types.SimpleNamespace(reset_position_ids=False, context_parallel_size=cp_size, seq_length=seq_length)

    def mk(name, **attrs):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    mpu = mk("megatron.core.mpu", get_context_parallel_rank=lambda: cp_rank,
             get_context_parallel_world_size=lambda: cp_size)
    mk("megatron")
    mk("megatron.training", get_args=lambda: args, get_adlr_autoresume=lambda: None)
    mk("megatron.core", DistributedDataParallel=type("DDP", (), {}), mpu=mpu)
    mk("megatron.core.tensor_parallel", param_is_not_tensor_parallel_duplicate=lambda p: True)
    mk("megatron.legacy")
    mk("megatron.legacy.model", Float16Module=type("Float16Module", (), {}))
    mk("megatron.legacy.model.module", param_is_not_shared=lambda p: True)
    path = os.path.join(REF_ROOT, "long_vita_megatron", "training", "utils.py")
    spec = importlib.util.spec_from_file_location("lv_ref_megatron_training_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    @contextlib.contextmanager
    def cpu_placement():
        o_arange, o_tensor, o_cuda = torch.arange, torch.tensor, torch.Tensor.cuda

        def strip(kw):
            if kw.get("device") == "cuda":
                kw.pop("device")
            kw.pop("pin_memory", None)
            return kw

        def fix(kw):
            kw = strip(kw)
            if isinstance(kw.get("device"), int):        # torch.cuda.current_device() used as a device
                kw.pop("device")
            return kw

        o_cur = torch.cuda.current_device
        torch.arange = lambda *a, **k: o_arange(*a, **fix(k))
        torch.tensor = lambda *a, **k: o_tensor(*a, **fix(k))
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.cuda.current_device = lambda: 0
        try:
            yield
        finally:
            torch.arange, torch.tensor, torch.Tensor.cuda = o_arange, o_tensor, o_cuda
            torch.cuda.current_device = o_cur

    return mod, cpu_placement


def load_megatron_rope(cp_size: int, cp_rank: int):
    """The reference's `long_vita_megatron/core/models/common/embeddings/rotary_pos_embedding.py`
    (RotaryEmbedding :50-122, get_pos_emb_on_this_cp_rank :36-47, apply_rotary_pos_emb_bshd :181-204), executed
    from /root/reference with `megatron.core.parallel_state` and `long_vita_megatron.training.utils` satisfied by
    the stand-ins of `load_megatron_training_utils`.  Returns (module, cpu_placement)."""
    utils_mod, cpu_placement = load_megatron_training_utils(cp_size, cp_rank, 0)
    import importlib.machinery

    ps = types.ModuleType("megatron.core.parallel_state")
    ps.__spec__ = importlib.machinery.ModuleSpec("megatron.core.parallel_state", loader=None)
    ps.get_context_parallel_world_size = lambda: cp_size
    ps.get_context_parallel_rank = lambda: cp_rank
    sys.modules["megatron.core.parallel_state"] = ps
    sys.modules["megatron.core"].parallel_state = ps
    for name in ("long_vita_megatron", "long_vita_megatron.training"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules["long_vita_megatron.training.utils"] = utils_mod
    path = os.path.join(REF_ROOT, "long_vita_megatron", "core", "models", "common", "embeddings", "rotary_pos_embedding.py")
    spec = importlib.util.spec_from_file_location("lv_ref_megatron_rope", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, cpu_placement


def load_megatron_embedding():
    """The reference's `LanguageModelEmbedding` (core/models/common/embeddings/language_model_embedding.py:13-174:
    vocabulary gather + the three image-feature merge modes `indices` / `pre_len` / `src_indices`+`tgt_indices`,
    transpose to [s, b, h]), executed from /root/reference.  Megatron base classes are stand-ins
    (`MegatronModule` = torch.nn.Module holding `config`; cp world size 1).  Returns the class."""
    import importlib.machinery

    import torch

    load_megatron_training_utils(1, 0, 0)             # installs the megatron.* stand-in modules

    def mk(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
            m.__path__ = []
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    class MegatronModule(torch.nn.Module):
        def __init__(self, config=None):
            super().__init__()
            self.config = config

    tp = mk("megatron.core.tensor_parallel", VocabParallelEmbedding=None)
    mk("megatron.core", tensor_parallel=tp)
    mk("megatron.core.transformer")
    mk("megatron.core.transformer.module", MegatronModule=MegatronModule)
    mk("megatron.core.transformer.transformer_config", TransformerConfig=object)
    path = os.path.join(REF_ROOT, "long_vita_megatron", "core", "models", "common", "embeddings", "language_model_embedding.py")
    spec = importlib.util.spec_from_file_location("lv_ref_megatron_embedding", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.LanguageModelEmbedding


def load_megatron_masked_linear():
    """The reference's `LinearWithGradAccumulationAndAsyncCommunication` (core/tensor_parallel/layers.py:365-534:
    forward / backward of the linear layer with `logit_mask` - masked_select of the rows before the GEMM,
    masked_scatter of dX after it, dW = dY^T sel), executed from /root/reference.  The module's Megatron imports
    are stand-ins; the only one that runs on this path (tp = 1, no sequence parallelism, no fused wgrad) is
    `megatron.core.utils.prepare_input_tensors_for_wgrad_compute`, restated from Megatron-LM core_r0.7.0
    (flatten [M, b, *] to 2-D).  Returns the autograd Function class."""
    import importlib.machinery

    load_megatron_training_utils(1, 0, 0)

    def mk(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
            m.__path__ = []
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    def prepare_input_tensors_for_wgrad_compute(grad_output, all_gathered_input):
        if grad_output.dim() == 3:
            grad_output = grad_output.contiguous().view(grad_output.shape[0] * grad_output.shape[1], grad_output.shape[2])
            all_gathered_input = all_gathered_input.contiguous().view(
                all_gathered_input.shape[0] * all_gathered_input.shape[1], all_gathered_input.shape[2])
        return grad_output, all_gathered_input

    none = lambda *a, **k: None   # noqa: E731
    mk("megatron.core.model_parallel_config", ModelParallelConfig=object)
    mk("megatron.core.parallel_state", get_global_memory_buffer=none, get_tensor_and_expert_parallel_rank=lambda: 0,
       get_tensor_and_expert_parallel_world_size=lambda: 1, get_tensor_model_parallel_group=none,
       get_tensor_model_parallel_rank=lambda: 0, get_tensor_model_parallel_world_size=lambda: 1)
    mk("megatron.core.dist_checkpointing")
    mk("megatron.core.dist_checkpointing.mapping", ShardedStateDict=dict)
    mk("megatron.core.transformer")
    mk("megatron.core.transformer.utils", make_sharded_tensors_for_checkpoint=none)
    mk("megatron.core.utils", make_tp_sharded_tensor_for_checkpoint=none,
       prepare_input_tensors_for_wgrad_compute=prepare_input_tensors_for_wgrad_compute)
    mk("megatron.core.tensor_parallel.mappings", copy_to_tensor_model_parallel_region=none,
       gather_from_sequence_parallel_region=none, gather_from_tensor_model_parallel_region=none,
       reduce_from_tensor_model_parallel_region=none, reduce_scatter_to_sequence_parallel_region=none,
       scatter_to_tensor_model_parallel_region=none)
    mk("megatron.core.tensor_parallel.random", get_cuda_rng_tracker=none, get_expert_parallel_rng_tracker_name=none)
    mk("megatron.core.tensor_parallel.utils", VocabUtility=object, divide=lambda a, b: a // b, split_tensor_along_last_dim=none)
    path = os.path.join(REF_ROOT, "long_vita_megatron", "core", "tensor_parallel", "layers.py")
    spec = importlib.util.spec_from_file_location("lv_ref_megatron_layers", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.LinearWithGradAccumulationAndAsyncCommunication


def load_megatron_generation(cp_size: int, cp_rank: int, group=None):
    """The reference's `long_vita_megatron/inference/text_generation/generation.py` (its inference-side
    `get_batch_on_this_cp_rank` :517-539 and `sync_output` :542-566: all-gather of the ranks' zig-zag shards and
    the un-permutation to global order), executed from /root/reference; `mpu` stand-ins report `cp_size`,
    `cp_rank` and `group`.  Returns the module."""
    import importlib.machinery

    utils_mod, _ = load_megatron_training_utils(cp_size, cp_rank, 0)

    def mk(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
            m.__path__ = []
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    none = lambda *a, **k: None   # noqa: E731
    mk("megatron.training", get_tokenizer=none)
    mpu = sys.modules["megatron.core.mpu"]
    mpu.get_context_parallel_group = lambda: group
    mk("megatron.inference")
    mk("megatron.inference.text_generation")
    mk("megatron.inference.text_generation.communication", copy_from_last_to_first_pipeline_stage=none,
       broadcast_from_last_pipeline_stage=none, broadcast_from_last_to_first_pipeline_stage=none)
    mk("megatron.inference.text_generation.forward_step", ForwardStep=object)
    mk("megatron.inference.text_generation.beam_utils", BeamHypotheses=object)
    mk("megatron.inference.text_generation.generation", _build_attention_mask_and_position_ids=none)
    for name in ("long_vita_megatron", "long_vita_megatron.training"):
        mk(name)
    sys.modules["long_vita_megatron.training.utils"] = utils_mod
    path = os.path.join(REF_ROOT, "long_vita_megatron", "inference", "text_generation", "generation.py")
    spec = importlib.util.spec_from_file_location("lv_ref_megatron_generation", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_checkpoint_converter():
    """`convert_checkpoint_from_megatron_to_transformers` and `safe_copy` of the reference's
    tools/hf2mcore_long_vita.py (:120-124, :373-510).  The script is a concatenation of files whose mid-file
    imports need a full Megatron install, so only these two function definitions are taken: their source text is
    read from /root/reference at call time (never stored) and executed in a namespace that holds `torch`.
    Returns the converter function."""
    import ast

    import torch

    path = os.path.join(REF_ROOT, "tools", "hf2mcore_long_vita.py")
    src = open(path).read()
    tree = ast.parse(src)
    wanted = {"safe_copy", "convert_checkpoint_from_megatron_to_transformers"}
    lines = src.splitlines()
    ns = {"torch": torch}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in wanted:
            start = min([node.lineno] + [d.lineno for d in node.decorator_list]) - 1
            code = "\n".join(lines[start : node.end_lineno])
            exec(compile(code, f"{path}:{node.name}", "exec"), ns)     # noqa: S102 - the reference's own function
    return ns["convert_checkpoint_from_megatron_to_transformers"]


def module_tree_from_state_dict(sd):
    """A torch.nn.Module tree whose attribute paths are the dotted keys of `sd` (numeric path components become
    ModuleLists) - the shape of object the reference's converter walks (`mgmodel.decoder.layers[i]...`)."""
    import torch

    class Box(torch.nn.Module):
        pass

    root = Box()
    for key, t in sd.items():
        parts = key.split(".")
        cur = root
        for p in parts[:-1]:
            if p not in cur._modules:
                cur.add_module(p, Box())
            cur = cur._modules[p]
        cur.register_parameter(parts[-1], torch.nn.Parameter(t.clone(), requires_grad=False))

    def listify(mod):
        for name, child in list(mod._modules.items()):
            listify(child)
            names = list(child._modules)
            if names and all(n.isdigit() for n in names) and not child._parameters:
                mod._modules[name] = torch.nn.ModuleList([child._modules[str(i)] for i in range(len(names))])

    listify(root)
    return root


def load_class_methods(rel_path: str, class_name: str, method_names, namespace=None):
    """Methods of a class of the reference, taken by name from the file's syntax tree (source text read from
    /root/reference at call time, never stored) and executed as plain functions in a namespace holding `torch` -
    for files whose module-level imports need a full Megatron install.  Returns {name: function(self, ...)}."""
    import ast

    import torch

    path = os.path.join(REF_ROOT, rel_path)
    src = open(path).read()
    lines = src.splitlines()
    ns = {"torch": torch}
    ns.update(namespace or {})          # names the methods' annotations / bodies expect at module level
    out = {}
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == class_name:
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name in method_names:
                    body = "\n".join(lines[item.lineno - 1 : item.end_lineno])
                    import textwrap

                    exec(compile(textwrap.dedent(body), f"{path}:{class_name}.{item.name}", "exec"), ns)   # noqa: S102
                    out[item.name] = ns[item.name]
    return out


def load_module_functions(rel_path: str, names, namespace=None):
    """Module-level functions of a reference file, taken by name from its syntax tree like `load_class_methods`
    (for files whose imports cannot be satisfied here).  Returns {name: function}; the functions share one namespace,
    so they can call each other."""
    import ast
    import textwrap

    path = os.path.join(REF_ROOT, rel_path)
    src = open(path).read()
    lines = src.splitlines()
    ns = dict(namespace or {})
    out = {}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            body = "\n".join(lines[node.lineno - 1 : node.end_lineno])
            exec(compile(textwrap.dedent(body), f"{path}:{node.name}", "exec"), ns)   # noqa: S102
            out[node.name] = ns[node.name]
    return out
