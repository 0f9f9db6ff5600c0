"""Minimal `yacs.config.CfgNode` (what /root/reference/config.py:10-160 uses: attribute tree, clone, defrost / freeze,
merge_from_file)."""
import copy

import yaml


class CfgNode(dict):
    def __init__(self, init=None):
        super().__init__()
        self.__dict__["_frozen"] = False
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        if self.__dict__.get("_frozen"):
            raise AttributeError(f"config is frozen: cannot set {k}")
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def _set_frozen(self, f):
        self.__dict__["_frozen"] = f
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(f)

    def defrost(self):
        self._set_frozen(False)

    def freeze(self):
        self._set_frozen(True)

    def _merge(self, other, path=""):
        for k, v in other.items():
            if k not in self:
                raise KeyError(f"Non-existent config key: {path}{k}")
            if isinstance(self[k], CfgNode) and isinstance(v, dict):
                self[k]._merge(v, path + k + ".")
            else:
                old = self[k]
                if isinstance(old, tuple) and isinstance(v, list):
                    v = tuple(v)
                elif isinstance(old, float) and isinstance(v, int) and not isinstance(v, bool):
                    v = float(v)
                elif isinstance(old, float) and isinstance(v, str):
                    v = float(v)          # yaml reads 1e-8 as a string
                self[k] = v

    def merge_from_file(self, path):
        with open(path) as f:
            self._merge(yaml.safe_load(f) or {})

    def __deepcopy__(self, memo):
        n = CfgNode()
        for k, v in self.items():
            dict.__setitem__(n, k, copy.deepcopy(v, memo))
        n.__dict__["_frozen"] = self.__dict__.get("_frozen", False)
        return n
