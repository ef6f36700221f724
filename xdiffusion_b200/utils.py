"""Config plumbing kept API-compatible with the reference (xdiffusion/utils.py:25-54,207-260):
``DotConfig``, ``load_yaml``, ``instantiate_from_config`` and friends.  YAML ``target:`` paths that
name reference classes (``xdiffusion.…``) resolve to this package's drop-in classes, so the
reference's own config files work unchanged."""
import importlib
from functools import partial
from typing import Any

import torch
import yaml

_PREFIX_FROM, _PREFIX_TO = "xdiffusion.", "xdiffusion_b200."


class DotConfig:
    """Attribute / item / ``in`` access over a nested dict (reference: utils.py:25-48)."""

    def __init__(self, cfg):
        self._cfg = cfg

    def __getattr__(self, k) -> Any:
        if k.startswith("__"):
            raise AttributeError(k)
        v = self._cfg[k]
        return DotConfig(v) if isinstance(v, dict) else v

    def __getitem__(self, k) -> Any:
        return self.__getattr__(k)

    def __contains__(self, k) -> bool:
        return k in self._cfg

    def to_dict(self):
        return self._cfg


def load_yaml(yaml_path: str) -> DotConfig:
    with open(yaml_path, "r") as fp:
        loader = getattr(yaml, "CLoader", yaml.SafeLoader)
        return DotConfig(yaml.load(fp, loader))


def resolve_target(path: str) -> str:
    return _PREFIX_TO + path[len(_PREFIX_FROM):] if path.startswith(_PREFIX_FROM) else path


def get_obj_from_str(string: str, reload: bool = False):
    module, cls = resolve_target(string).rsplit(".", 1)
    try:
        mod = importlib.import_module(module)
    except ModuleNotFoundError as e:
        raise NotImplementedError(
            f"`{string}` is outside the sampling hot path covered by xdiffusion_b200 (see DESIGN.md)") from e
    if reload:
        importlib.reload(mod)
    if not hasattr(mod, cls):
        raise NotImplementedError(
            f"`{string}` is outside the sampling hot path covered by xdiffusion_b200 (see DESIGN.md)")
    return getattr(mod, cls)


def _as_dict(config):
    return config.to_dict() if isinstance(config, DotConfig) else config


def instantiate_from_config(config, use_config_struct: bool = False) -> Any:
    cfg = _as_dict(config)
    if "target" not in cfg:
        if cfg in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate.")
    cls = get_obj_from_str(cfg["target"])
    params = cfg.get("params", dict()) or dict()
    if use_config_struct or cfg.get("instantiate_with_config_struct", False):
        return cls(DotConfig(params) if not isinstance(params, DotConfig) else params)
    return cls(**params)


def instantiate_partial_from_config(config, use_config_struct: bool = False) -> Any:
    cfg = _as_dict(config)
    if "target" not in cfg:
        if cfg in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate.")
    cls = get_obj_from_str(cfg["target"])
    params = cfg.get("params", dict()) or dict()
    if use_config_struct:
        return partial(cls, DotConfig(params))
    return partial(cls, **params)


def normalize_to_neg_one_to_one(img):
    return img * 2 - 1


def unnormalize_to_zero_to_one(t: torch.Tensor) -> torch.Tensor:
    """(clamp(t,-1,1)+1)/2 on the device kernel (reference: utils.py:62-64)."""
    out = torch.empty_like(t)
    torch.ops.xdb200.unnormalize(t.contiguous(), out)
    return out


def broadcast_from_left(x, shape):
    assert len(shape) >= x.ndim
    return torch.broadcast_to(x.reshape(x.shape + (1,) * (len(shape) - x.ndim)), shape)
