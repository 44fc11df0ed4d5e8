"""Import the reference's pure-PyTorch modules in THIS container (no mmcv/mmdet installed).

Used only by the golden-vector generators under tools/ (test infrastructure).  It installs
throw-away stand-ins for the mmcv / mmdet / mmdet3d names the reference files import at
module level, and maps a synthetic package `vidar_ref` onto
/root/reference/projects/mmdet3d_plugin/bevformer so that files are loaded one by one
(relative imports work, the plugin's heavy __init__ chain is not executed).
`mmcv.ops.multi_scale_deform_attn.multi_scale_deformable_attn_pytorch` -- the reference's CPU
path for MSDA, whose source lives in the un-vendored mmcv-full 1.4.0 -- is provided by the
same-lineage implementation in `transformers` (NOT by this repo's oracle, so the goldens stay
independent of it).
"""
import importlib
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_BEVFORMER = "/root/reference/projects/mmdet3d_plugin/bevformer"


class _Anything:
    """Placeholder usable as decorator, base class factory, callable or namespace."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _Permissive(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _Registry:
    def __init__(self):
        self.modules = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.modules[name or cls.__name__] = cls
            return cls
        return deco(module) if module is not None else deco

    def build(self, cfg):
        cfg = dict(cfg)
        return self.modules[cfg.pop("type")](**cfg)


def _passthrough_decorator(*dargs, **dkw):
    if len(dargs) == 1 and callable(dargs[0]) and not dkw:
        return dargs[0]

    def deco(fn):
        return fn
    return deco


def _msda_pytorch(value, value_spatial_shapes, sampling_locations, attention_weights):
    from transformers.models.deformable_detr.modeling_deformable_detr import MultiScaleDeformableAttention
    shapes = [(int(h), int(w)) for h, w in value_spatial_shapes.tolist()]
    return MultiScaleDeformableAttention().forward(value, value_spatial_shapes, shapes, None,
                                                   sampling_locations, attention_weights, 64)


def install():
    if "vidar_ref" in sys.modules:
        return sys.modules["vidar_ref"]
    attention, layer, seq = _Registry(), _Registry(), _Registry()

    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
            self.init_cfg = init_cfg

    def constant_init(m, val, bias=0):
        nn.init.constant_(m.weight, val)
        if m.bias is not None:
            nn.init.constant_(m.bias, bias)

    def xavier_init(m, gain=1, bias=0, distribution="normal"):
        if m is None:     # the reference calls xavier_init(self.output_proj) with output_proj=None
            return
        (nn.init.xavier_uniform_ if distribution == "uniform" else nn.init.xavier_normal_)(m.weight, gain=gain)
        if m.bias is not None:
            nn.init.constant_(m.bias, bias)

    def build_attention(cfg):
        return attention.build(cfg)

    def mod(name, **attrs):
        m = _Permissive(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("mmcv", __path__=[])
    mod("mmcv.cnn", __path__=[], Linear=nn.Linear, xavier_init=xavier_init, constant_init=constant_init,
        bias_init_with_prob=lambda p: float(-torch.log(torch.tensor((1 - p) / p))))
    mod("mmcv.cnn.bricks", __path__=[])
    mod("mmcv.cnn.bricks.registry", ATTENTION=attention, TRANSFORMER_LAYER=layer,
        TRANSFORMER_LAYER_SEQUENCE=seq)
    mod("mmcv.cnn.bricks.transformer", build_attention=build_attention,
        TransformerLayerSequence=BaseModule)
    mod("mmcv.runner", __path__=[], force_fp32=_passthrough_decorator, auto_fp16=_passthrough_decorator,
        BaseModule=BaseModule)
    mod("mmcv.runner.base_module", BaseModule=BaseModule, ModuleList=nn.ModuleList, Sequential=nn.Sequential)
    mod("mmcv.utils", __path__=[], ext_loader=types.SimpleNamespace(load_ext=lambda *a, **k: _Anything()),
        TORCH_VERSION=torch.__version__, digit_version=lambda v: (1, 10, 0),
        deprecated_api_warning=lambda *a, **k: _passthrough_decorator)
    mod("mmcv.ops", __path__=[])
    mod("mmcv.ops.multi_scale_deform_attn", multi_scale_deformable_attn_pytorch=_msda_pytorch)
    for name in ("mmdet", "mmdet.models", "mmdet.models.utils", "mmdet3d", "mmdet3d.models",
                 "mmdet3d.models.losses"):
        mod(name, __path__=[])
    # decoder.py imports plotting libraries it never uses on this path
    for name in ("cv2", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except ImportError:
                mod(name, __path__=[])

    pkg = types.ModuleType("vidar_ref")
    pkg.__path__ = [REF_BEVFORMER]
    sys.modules["vidar_ref"] = pkg
    for sub in ("modules", "dense_heads"):
        m = types.ModuleType(f"vidar_ref.{sub}")
        m.__path__ = [os.path.join(REF_BEVFORMER, sub)]
        sys.modules[f"vidar_ref.{sub}"] = m
    rays = types.ModuleType("vidar_ref.modules.ray_operations")
    rays.__path__ = [os.path.join(REF_BEVFORMER, "modules", "ray_operations")]
    sys.modules["vidar_ref.modules.ray_operations"] = rays
    # `from ..utils import e2e_predictor_utils` JIT-compiles CUDA at import: stub it out
    utils = _Permissive("vidar_ref.utils")
    utils.__path__ = []
    sys.modules["vidar_ref.utils"] = utils
    sys.modules["vidar_ref.utils.e2e_predictor_utils"] = _Permissive("vidar_ref.utils.e2e_predictor_utils")
    pkg.registries = dict(ATTENTION=attention)
    rays.LatentRendering = importlib.import_module(
        "vidar_ref.modules.ray_operations.latent_rendering").LatentRendering
    return pkg


def load_functions(relpath, names):
    """Execute ONLY the named top-level functions of a reference file (plus its plain imports that
    resolve here) and return them as a dict -- for files whose module body cannot run in this
    container (e2e_predictor_utils.py JIT-compiles CUDA extensions at import).  The source is read
    from /root/reference at call time; nothing is copied into the repo."""
    import ast
    path = os.path.join(REF_BEVFORMER, relpath)
    with open(path) as fh:
        tree = ast.parse(fh.read(), filename=path)
    ns = {"__name__": "vidar_ref._functions"}
    for node in tree.body:
        if isinstance(node, (ast.Import, ast.ImportFrom)) and not getattr(node, "level", 0):
            try:
                exec(compile(ast.Module([node], []), path, "exec"), ns)
            except ImportError:
                pass
        elif isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return {n: ns[n] for n in names}


def install_e2e_utils():
    """Put the pure helpers of utils/e2e_predictor_utils.py on the stub module the heads import."""
    install()
    stub = sys.modules["vidar_ref.utils.e2e_predictor_utils"]
    fns = load_functions(os.path.join("utils", "e2e_predictor_utils.py"),
                         ("coords_to_voxel_grids", "get_bev_grids", "get_bev_grids_3d", "get_inside_mask"))
    for k, v in fns.items():
        setattr(stub, k, v)
    sys.modules["vidar_ref.utils"].e2e_predictor_utils = stub      # `from ..utils import e2e_predictor_utils`
    return stub


def load(dotted):
    """e.g. load('modules.spatial_cross_attention') -> module object of the reference file."""
    install()
    return importlib.import_module(f"vidar_ref.{dotted}")
