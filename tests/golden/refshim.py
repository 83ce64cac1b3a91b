"""Import shim for running the UNMODIFIED reference (/root/reference) on CPU in the build container.

Only used by ``tests/golden/make_golden.py`` (fixture generation) and by the optional
``tests/test_oracle_vs_reference.py`` cross-check, both of which are skipped when
``/root/reference`` is absent (it never exists on the GPU box).  Nothing on the product path
imports this module.

What it does:
  * injects a minimal ``yacs.config.CfgNode`` (attribute dict + yaml/list merge) because yacs is
    not installed in this image;
  * injects empty stub modules for ``matplotlib``, ``librosa``, ``ffmpeg``, ``cv2``, ``tensorboard``
    bits that the reference imports at module scope but never calls on the hot path;
  * patches ``Tensor.cuda`` / ``Module.cuda`` to identity while no GPU is visible
    (reference hard-codes ``.cuda()``: core/pipelines/voice2pose.py:86-90);
  * aliases ``np.float`` (removed in numpy>=1.24; reference uses it at
    core/datasets/gesture_dataset.py:175,195).
No reference source is copied: the reference is imported from where it lies.
"""
import copy
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SDT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "core", "networks"))


class CfgNode(dict):
    """Tiny stand-in for yacs.config.CfgNode: nested attribute dict with merge helpers."""

    def __init__(self, init=None):
        super().__init__()
        self.__dict__["_frozen"] = False
        if init:
            for k, v in init.items():
                self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.__dict__.get("_frozen"):
            raise AttributeError("CfgNode is frozen")
        self[name] = value

    def clone(self):
        return copy.deepcopy(self)

    def freeze(self):
        self.__dict__["_frozen"] = True
        for v in self.values():
            if isinstance(v, CfgNode):
                v.freeze()

    def defrost(self):
        self.__dict__["_frozen"] = False
        for v in self.values():
            if isinstance(v, CfgNode):
                v.defrost()

    def _merge(self, other):
        for k, v in other.items():
            if k not in self:
                raise KeyError("Non-existent config key: %s" % k)
            if isinstance(v, dict) and isinstance(self[k], CfgNode):
                self[k]._merge(v)
            else:
                if isinstance(v, str):      # yacs decodes string leaves with literal_eval ('1e-4' -> float)
                    import ast
                    try:
                        v = ast.literal_eval(v)
                    except (ValueError, SyntaxError):
                        pass
                self[k] = v

    def merge_from_file(self, path):
        import yaml
        with open(path) as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_list(self, opts):
        import ast
        assert len(opts) % 2 == 0
        for key, val in zip(opts[0::2], opts[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                node = node[p]
            if isinstance(val, str):
                try:
                    val = ast.literal_eval(val)
                except (ValueError, SyntaxError):
                    pass
            if parts[-1] not in node:
                raise KeyError("Non-existent config key: %s" % key)
            node[parts[-1]] = val


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


def install():
    """Make ``import core...`` / ``import configs...`` resolve to the reference, on CPU."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    try:
        import yacs.config  # noqa: F401
    except ImportError:
        yacs = _stub("yacs")
        yacs.config = _stub("yacs.config", CfgNode=CfgNode)

    for name in ("matplotlib", "librosa", "ffmpeg"):
        try:
            __import__(name)
        except ImportError:
            _stub(name)
    try:
        import matplotlib.pyplot  # noqa: F401
    except ImportError:
        sys.modules["matplotlib"].pyplot = _stub("matplotlib.pyplot")
    try:
        import cv2  # noqa: F401
    except ImportError:
        _stub("cv2")
    try:
        import torch.utils.tensorboard  # noqa: F401
    except Exception:
        tb = _stub("torch.utils.tensorboard", SummaryWriter=object)
        import torch.utils
        torch.utils.tensorboard = tb

    import numpy as np
    if not hasattr(np, "float"):
        np.float = float

    import torch
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self


def get_cfg(config_name, opts=()):
    """Defaults (configs/default.py) + one of the reference's YAML overlays + KEY VALUE overrides."""
    install()
    from configs.default import get_cfg_defaults
    cfg = get_cfg_defaults()
    cfg.merge_from_file(os.path.join(REFERENCE_ROOT, "configs", config_name + ".yaml"))
    cfg.merge_from_list(list(opts))
    return cfg


def make_dataset_stub(cfg):
    """A GestureDataset instance without its csv-reading __init__ (gesture_dataset.py:15-48)."""
    install()
    from core.datasets.gesture_dataset import GestureDataset
    ds = GestureDataset.__new__(GestureDataset)
    ds.cfg = cfg.DATASET
    ds.root_node, ds.hand_root_l, ds.hand_root_r, ds.head_root = 1, 6, 3, 39
    return ds
