"""Model registry with the reference's interface (lib/model_zoo/common/get_model.py:53-111):
`get_model()(cfg)` builds `registry[cfg.type](**cfg.args)` and optionally loads `cfg.pretrained`;
`@register(name, version)` adds a class.  cfg may be an attribute-style dict (EasyDict) or a plain dict."""
import copy

import torch


def _get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def load_state_dict(net, model_path):
    """lib/model_zoo/common/get_model.py:10-22: parameters missing from the file keep their current values."""
    paras = torch.load(model_path, map_location=torch.device('cpu'))
    merged = net.state_dict()
    merged.update(paras)
    net.load_state_dict(merged)


def save_state_dict(net, path):
    net = net.module if hasattr(net, 'module') else net
    torch.save(net.state_dict(), path)


class _Registry:
    def __init__(self):
        self.model = {}
        self.version = {}

    def register(self, model, name, version='x'):
        self.model[name] = model
        self.version[name] = version

    def __call__(self, cfg):
        if cfg is None:
            return None
        t = _get(cfg, 'type')
        if t not in self.model:
            raise KeyError(f'model type {t!r} is not provided by shgan_b200 (generator-forward path only); '
                           f'known types: {sorted(self.model)}')
        args = copy.deepcopy(dict(_get(cfg, 'args') or {}))
        net = self.model[t](**args)
        pretrained = _get(cfg, 'pretrained')
        if pretrained is not None:
            load_state_dict(net, pretrained)
        return net

    def get_version(self, name):
        return self.version[name]


_instance = _Registry()


def get_model():
    return _instance


def register(name, version='x'):
    def wrapper(cls):
        _instance.register(cls, name, version)
        return cls
    return wrapper
