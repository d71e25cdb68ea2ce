"""Python-file configs -> attribute dict, as det3d/torchie/utils/config.py:12-29,51-162 exposes them,
plus det3d/utils/config_tool.py:39-53 (`get_downsample_factor`, imported *by the config files*)."""
import importlib.util
import os

import numpy as np


class ConfigDict(dict):
    """dict with attribute access; nested dicts (also inside lists/tuples) are converted recursively."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError("'%s' object has no attribute '%s'" % (type(self).__name__, k))

    def __delattr__(self, k):
        del self[k]

    def setdefault(self, k, default=None):
        if k not in self:
            self[k] = default
        return self[k]

    def update(self, *args, **kwargs):
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    def copy(self):
        return ConfigDict(self)

    def to_dict(self):
        def un(v):
            if isinstance(v, dict):
                return {k: un(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return type(v)(un(x) for x in v)
            return v
        return un(self)


class Config:
    """`Config.fromfile('x.py')` executes the file as a module and exposes its public names."""

    def __init__(self, cfg_dict=None, filename=None):
        cfg_dict = {} if cfg_dict is None else cfg_dict
        if not isinstance(cfg_dict, dict):
            raise TypeError("cfg_dict must be a dict, but got %s" % type(cfg_dict))
        object.__setattr__(self, "_cfg_dict", ConfigDict(cfg_dict))
        object.__setattr__(self, "_filename", filename)
        text = ""
        if filename:
            with open(filename, "r") as f:
                text = f.read()
        object.__setattr__(self, "_text", text)

    @staticmethod
    def fromfile(filename):
        filename = os.path.abspath(os.path.expanduser(filename))
        if not os.path.isfile(filename):
            raise FileNotFoundError('file "%s" does not exist' % filename)
        if not filename.endswith(".py"):
            raise IOError("Only py type is supported")
        name = os.path.basename(filename)[:-3]
        if "." in name:
            raise ValueError("Dots are not allowed in config file path.")
        import sys
        if "det3d" not in sys.modules:      # reference configs import det3d.utils.config_tool at load time
            from .compat import install
            install()
        spec = importlib.util.spec_from_file_location("_fdcfg_" + name, filename)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        import types
        cfg = {k: v for k, v in vars(mod).items() if not k.startswith("__") and not isinstance(v, types.ModuleType)
               and not callable(v)}
        return Config(cfg, filename=filename)

    filename = property(lambda self: self._filename)
    text = property(lambda self: self._text)

    def __repr__(self):
        return "Config (path: %s): %r" % (self._filename, self._cfg_dict)

    def __len__(self):
        return len(self._cfg_dict)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __setattr__(self, name, value):
        self._cfg_dict[name] = value

    def __setitem__(self, name, value):
        self._cfg_dict[name] = value

    def __iter__(self):
        return iter(self._cfg_dict)

    def get(self, key, default=None):
        return self._cfg_dict.get(key, default)


def get_downsample_factor(model_config):
    """prod(ds_layer_strides) / us_layer_strides[-1] * backbone.ds_factor  (config_tool.py:39-53)."""
    if "neck" not in model_config:
        model_config = model_config["first_stage_cfg"]
    neck = model_config["neck"]
    factor = float(np.prod(neck.get("ds_layer_strides", [1])))
    us = neck.get("us_layer_strides", [])
    if len(us) > 0:
        factor /= us[-1]
    factor = int(factor * model_config["backbone"]["ds_factor"])
    assert factor > 0
    return factor
