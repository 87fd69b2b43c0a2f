"""Registry surface of the reference (mmcv / mmdet / mmdet3d are not installed here).

Mirrors ``mmdet3d/models/registry.py:1-5`` (VOXEL_ENCODERS, MIDDLE_ENCODERS, FUSION_LAYERS),
``mmdet3d/models/builder.py:8-68`` (build_* helpers) and the mmcv pieces the hot path
touches (``Registry``, ``build_from_cfg``, ``build_conv_layer``, ``build_norm_layer``,
``Config.fromfile`` for the flat python config files), so that
``configs/MSMDFusion_nusc_voxel_LC.py`` and ``configs/transfusion_nusc_voxel_L.py`` load
unchanged.
"""
import copy
import os

from torch import nn


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def __len__(self):
        return len(self._module_dict)

    def __contains__(self, key):
        return self.get(key) is not None

    def __repr__(self):
        return f'Registry(name={self._name}, items={sorted(self._module_dict)})'

    def get(self, key):
        return self._module_dict.get(key)

    def _register(self, cls, name=None, force=False):
        name = name or cls.__name__
        if not force and name in self._module_dict:
            raise KeyError(f'{name} is already registered in {self._name}')
        self._module_dict[name] = cls

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register(module, name, force)
            return module

        def _wrap(cls):
            self._register(cls, name, force)
            return cls

        return _wrap


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict):
        raise TypeError(f'cfg must be a dict, got {type(cfg)}')
    if 'type' not in cfg and not (default_args and 'type' in default_args):
        raise KeyError(f'`cfg` or `default_args` must contain the key "type", got {cfg}')
    args = copy.deepcopy(dict(cfg))
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop('type')
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError(f'{obj_type} is not in the {registry.name} registry')
    else:
        obj_cls = obj_type
    return obj_cls(**args)


DETECTORS = Registry('detector')
BACKBONES = Registry('backbone')
NECKS = Registry('neck')
HEADS = Registry('head')
VOXEL_ENCODERS = Registry('voxel_encoder')      # mmdet3d/models/registry.py:3
MIDDLE_ENCODERS = Registry('middle_encoder')    # :4
FUSION_LAYERS = Registry('fusion_layer')        # :5 (registry object only; no layer on the path)
CONV_LAYERS = Registry('conv layer')            # mmcv.cnn CONV_LAYERS (bug_fix/conv.py:22)
NORM_LAYERS = Registry('norm layer')

CONV_LAYERS.register_module('Conv1d', module=nn.Conv1d)
CONV_LAYERS.register_module('Conv2d', module=nn.Conv2d)
CONV_LAYERS.register_module('Conv3d', module=nn.Conv3d)
CONV_LAYERS.register_module('Conv', module=nn.Conv2d)
NORM_LAYERS.register_module('BN', module=nn.BatchNorm2d)
NORM_LAYERS.register_module('BN1d', module=nn.BatchNorm1d)
NORM_LAYERS.register_module('BN2d', module=nn.BatchNorm2d)
NORM_LAYERS.register_module('BN3d', module=nn.BatchNorm3d)
NORM_LAYERS.register_module('LN', module=nn.LayerNorm)

UPSAMPLE_LAYERS = Registry('upsample layer')
UPSAMPLE_LAYERS.register_module('nearest', module=nn.Upsample)
UPSAMPLE_LAYERS.register_module('bilinear', module=nn.Upsample)
UPSAMPLE_LAYERS.register_module('deconv', module=nn.ConvTranspose2d)

_NORM_ABBR = {'BN': 'bn', 'BN1d': 'bn', 'BN2d': 'bn', 'BN3d': 'bn', 'LN': 'ln'}


def build_conv_layer(cfg, *args, **kwargs):
    """mmcv.cnn.build_conv_layer: cfg None -> Conv2d; extra cfg keys become kwargs."""
    cfg_ = dict(type='Conv2d') if cfg is None else dict(cfg)
    layer_type = cfg_.pop('type')
    cls = CONV_LAYERS.get(layer_type)
    if cls is None:
        raise KeyError(f'Unrecognized conv type {layer_type}')
    return cls(*args, **kwargs, **cfg_)


def build_norm_layer(cfg, num_features, postfix=''):
    """mmcv.cnn.build_norm_layer -> (name, layer)."""
    cfg_ = dict(cfg)
    layer_type = cfg_.pop('type')
    cls = NORM_LAYERS.get(layer_type)
    if cls is None:
        raise KeyError(f'Unrecognized norm type {layer_type}')
    requires_grad = cfg_.pop('requires_grad', True)
    cfg_.setdefault('eps', 1e-5)
    layer = cls(num_features, **cfg_)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return _NORM_ABBR.get(layer_type, 'norm') + str(postfix), layer


def build_upsample_layer(cfg, *args, **kwargs):
    """mmcv.cnn.build_upsample_layer: 'deconv' -> ConvTranspose2d, 'nearest' / 'bilinear' -> Upsample(mode=type)."""
    cfg_ = dict(cfg)
    layer_type = cfg_.pop('type')
    cls = UPSAMPLE_LAYERS.get(layer_type)
    if cls is None:
        raise KeyError(f'Unrecognized upsample type {layer_type}')
    if cls is nn.Upsample:
        cfg_['mode'] = layer_type
    return cls(*args, **kwargs, **cfg_)


def build_voxel_encoder(cfg):
    return build_from_cfg(cfg, VOXEL_ENCODERS)


def build_middle_encoder(cfg):
    return build_from_cfg(cfg, MIDDLE_ENCODERS)


def build_fusion_layer(cfg):
    return build_from_cfg(cfg, FUSION_LAYERS)


def build_backbone(cfg):
    return build_from_cfg(cfg, BACKBONES)


def build_neck(cfg):
    return build_from_cfg(cfg, NECKS)


def build_head(cfg):
    return build_from_cfg(cfg, HEADS)


def build_detector(cfg, train_cfg=None, test_cfg=None):
    return build_from_cfg(cfg, DETECTORS, dict(train_cfg=train_cfg, test_cfg=test_cfg))


class ConfigDict(dict):
    """Attribute-access dict (mmcv.ConfigDict)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value


def _to_cfgdict(v):
    if isinstance(v, dict):
        return ConfigDict({k: _to_cfgdict(x) for k, x in v.items()})
    if isinstance(v, list):
        return [_to_cfgdict(x) for x in v]
    if isinstance(v, tuple):
        return tuple(_to_cfgdict(x) for x in v)
    return v


class Config(ConfigDict):
    """Loader for the flat (no ``_base_``) python config files of the hot path."""

    @staticmethod
    def fromfile(filename):
        filename = os.path.abspath(os.path.expanduser(filename))
        with open(filename) as f:
            src = f.read()
        scope = {'__file__': filename}
        exec(compile(src, filename, 'exec'), scope)
        if '_base_' in scope:
            raise NotImplementedError('config inheritance (_base_) is outside the hot path')
        cfg = {k: v for k, v in scope.items()
               if not k.startswith('__') and not callable(v) and not isinstance(v, type(os))}
        return Config(_to_cfgdict(cfg))
