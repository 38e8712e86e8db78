"""oracle/ref_loader.py -- TEST INFRASTRUCTURE ONLY; works only where /root/reference exists.

Imports the UNMODIFIED reference modules that define the hot path

    /root/reference/models/sampling.py
    /root/reference/models/unigen.py
    /root/reference/training/prompting_utils.py

in this container.  `models/unigen.py` depends on `diffusers` (absent here) only for
`ConfigMixin` / `register_to_config` / `ModelMixin` (models/modeling_utils.py:26-45) and on
the SigLIP builder; those two modules are replaced by the minimal stand-ins below, the
reference files themselves are executed as they are.  Used by oracle/make_golden.py to produce
tests/golden/*.npz and by tests that pin the oracle when the reference tree is mounted.
"""
from __future__ import annotations

import functools
import importlib
import importlib.util
import inspect
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("UNIGEN_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "unigen.py"))


class _Config(dict):
    """dict with attribute access -- the slice of diffusers' FrozenDict the path uses."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e


def _install_stubs():
    import torch

    if "models.unigen" in sys.modules and getattr(sys.modules["models"], "_ug_stub", False):
        return
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)          # utils.checkpoint_registry, components.core

    models = types.ModuleType("models")
    models.__path__ = [os.path.join(REFERENCE_ROOT, "models")]
    models._ug_stub = True
    sys.modules["models"] = models

    mu = types.ModuleType("models.modeling_utils")

    class ConfigMixin:
        def register_to_config(self, **kw):
            if not hasattr(self, "_cfg"):
                self._cfg = _Config()
            self._cfg.update(kw)

        @property
        def config(self):
            return self._cfg

    def register_to_config(init):
        @functools.wraps(init)
        def wrapper(self, *a, **kw):
            sig = inspect.signature(init)
            bound = sig.bind(self, *a, **kw)
            bound.apply_defaults()
            cfg = {k: v for k, v in bound.arguments.items() if k not in ("self", "kwargs")}
            cfg.update(bound.arguments.get("kwargs", {}))
            ConfigMixin.register_to_config(self, **cfg)
            init(self, *a, **kw)
        return wrapper

    mu.ConfigMixin = ConfigMixin
    mu.ModelMixin = torch.nn.Module
    mu.register_to_config = register_to_config
    sys.modules["models.modeling_utils"] = mu

    me = types.ModuleType("models.multimodal_encoder")
    me.__path__ = [os.path.join(REFERENCE_ROOT, "models", "multimodal_encoder")]
    sys.modules["models.multimodal_encoder"] = me
    bld = types.ModuleType("models.multimodal_encoder.builder")

    def get_vision_tower(*a, **k):  # understanding branch: never reached on the T2I path
        raise NotImplementedError("vision tower is outside the T2I decoding path")

    bld.get_vision_tower = get_vision_tower
    sys.modules["models.multimodal_encoder.builder"] = bld


def load_sampling():
    """models/sampling.py (torch only; importable standalone)."""
    spec = importlib.util.spec_from_file_location(
        "_ug_ref_sampling", os.path.join(REFERENCE_ROOT, "models", "sampling.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_unigen():
    """models/unigen.py, unmodified, under the stand-ins above.  Returns the module."""
    _install_stubs()
    return importlib.import_module("models.unigen")


def load_prompting():
    """training/prompting_utils.py, unmodified."""
    spec = importlib.util.spec_from_file_location(
        "_ug_ref_prompting", os.path.join(REFERENCE_ROOT, "training", "prompting_utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def qwen_config_dir(path: str, *, hidden_size=1536, intermediate_size=8960, num_hidden_layers=28,
                    num_attention_heads=12, num_key_value_heads=2, vocab_size=151936,
                    max_position_embeddings=32768, rope_theta=1000000.0, rms_norm_eps=1e-6,
                    tie_word_embeddings=True) -> str:
    """Write a Qwen2 config.json (public Qwen2.5-1.5B-Instruct values by default) for
    AutoConfig.from_pretrained(llm_model_path) at models/unigen.py:52."""
    import json
    os.makedirs(path, exist_ok=True)
    cfg = dict(architectures=["Qwen2ForCausalLM"], model_type="qwen2", hidden_act="silu",
               hidden_size=hidden_size, intermediate_size=intermediate_size,
               num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
               num_key_value_heads=num_key_value_heads, vocab_size=vocab_size,
               max_position_embeddings=max_position_embeddings, rope_theta=rope_theta,
               rms_norm_eps=rms_norm_eps, tie_word_embeddings=tie_word_embeddings,
               attention_dropout=0.0, initializer_range=0.02, use_sliding_window=False,
               sliding_window=None, max_window_layers=num_hidden_layers, torch_dtype="float32",
               bos_token_id=151643, eos_token_id=151645, use_cache=False)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(cfg, f)
    return path


def build_reference_unigen(cfg_dir: str, vocab_size: int, codebook_size: int, num_vq_tokens: int):
    """Instantiate the reference's UniGen with a random-init Qwen2 backbone (load_from_pretrained=True
    takes the `model_fn(config)` branch, models/unigen.py:58-65)."""
    ug = load_unigen()
    model = ug.UniGen(w_und_encoder=False, vocab_size=vocab_size, llm_vocab_size=vocab_size - codebook_size - 1,
                      llm_model_path=cfg_dir, codebook_size=codebook_size, num_vq_tokens=num_vq_tokens,
                      load_from_pretrained=True)
    model.llm.config._attn_implementation = "sdpa"
    return model.eval()


def load_magvitv2():
    """models/multimodal_encoder/magvitv2.py (+ common_modules.py), unmodified, under the stand-ins above.
    Returns the module (VQGANDecoder, LFQuantizer, MAGVITv2).  `models/misc.py` (config helpers over omegaconf /
    jaxtyping / typeguard, absent here and unused by the decoder) is replaced by an empty stand-in."""
    _install_stubs()
    if "models.misc" not in sys.modules:
        sys.modules["models.misc"] = types.ModuleType("models.misc")
    return importlib.import_module("models.multimodal_encoder.magvitv2")
