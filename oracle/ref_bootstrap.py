"""Import the *real* reference (Dwaynekj/xdiffusion) in this authoring container.

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/make_golden.py`` (fixture
generation) and by optional CPU tests that are skipped when ``/root/reference``
is absent (it never exists on the GPU box).  Nothing under ``xdiffusion_b200/``
imports this file.

The reference imports ``accelerate``, ``torchinfo`` and ``soundfile`` at module
top (xdiffusion/diffusion/ddpm.py:8,13; xdiffusion/utils.py:10) although the
sampling path never uses them; they are not installed here, so empty stub
modules are registered.
"""
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("XDIFFUSION_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "xdiffusion"))


def bootstrap():
    """Make ``import xdiffusion`` resolve to the reference tree."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    # transformers resolves these lazily with find_spec: import them BEFORE stubbing.
    from transformers import (  # noqa: F401
        AutoModelForCausalLM, AutoProcessor, AutoTokenizer, CLIPTextModel,
        CLIPTextModelWithProjection, CLIPTokenizer, CLIPVisionModelWithProjection,
        T5EncoderModel, T5Tokenizer)
    for name in ("accelerate", "torchinfo", "soundfile"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)
            sys.modules[name] = m
    sys.modules["torchinfo"].summary = lambda *a, **k: ""
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load_reference_model(config_relpath: str, patch_config=None):
    """Build the reference GaussianDiffusion_DDPM from one of its own YAML files."""
    bootstrap()
    from xdiffusion.diffusion.ddpm import GaussianDiffusion_DDPM
    from xdiffusion.utils import DotConfig, load_yaml
    cfg = load_yaml(os.path.join(REFERENCE_ROOT, config_relpath))
    if patch_config is not None:
        d = cfg.to_dict()
        patch_config(d)
        cfg = DotConfig(d)
    return GaussianDiffusion_DDPM(cfg).eval()


def strip_t5_from_pixart(d):
    """C4: the T5 tokenizer/encoder need HF weights (no network).  Drop those two
    projections and feed ``text_embeddings`` directly."""
    cond = d["diffusion"]["score_network"]["params"]["conditioning"]
    cond["signals"] = ["timestep", "classes"]
    for k in ("text_tokens", "text_prompts"):
        cond["projections"].pop(k, None)
    cond["context_transformer_head"] = [
        c for c in cond["context_transformer_head"]
        if c["params"].get("projection_key") not in ("text_tokens", "text_prompts")]


def strip_t5_from_imagen(d):
    """C7 (configs/image/mnist/imagen_base.yaml): the T5 tokenizer (context preprocessor) and encoder (``text_tokens``
    projection + TextTokenProjectionAdapter) need HF weights.  Drop them and feed ``text_embeddings`` (B, 77, 768)
    directly; PooledTextEmbeddingsToTimestep and every SpatialCrossAttention(context_dim=768) stay as they are."""
    diff = d["diffusion"]
    diff["context_preprocessing"] = [{"target": "xdiffusion.context.IgnoreContextAdapter", "params": {}}]
    cond = diff["score_network"]["params"]["conditioning"]
    cond["signals"] = ["timestep"]
    cond["projections"].pop("text_tokens", None)
    cond["context_transformer_head"] = [c for c in cond["context_transformer_head"]
                                        if not c["target"].endswith("TextTokenProjectionAdapter")]
