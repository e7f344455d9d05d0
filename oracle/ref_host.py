"""Host the UNMODIFIED reference classes (BaseModel, TIMMModel, AVTh) from /root/reference on CPU.

Only usable in the authoring container (the GPU box has no /root/reference). Installs stub modules for the
packages the reference imports but this image lacks (hydra, omegaconf, submitit, pretrainedmodels, timm);
`timm.create_model` is served by oracle.vit (the restated timm-0.4.12 ViT). TEST INFRASTRUCTURE.
"""
import importlib
import importlib.machinery
import os
import sys
import types

REF = os.environ.get("AVT_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "models"))


def _locate(path):
    mod, _, attr = path.rpartition(".")
    return getattr(importlib.import_module(mod), attr)


class _Cfg(dict):
    """Minimal OmegaConf-like node: attribute access + `in` + item assignment."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return _Cfg(v) if isinstance(v, dict) and not isinstance(v, _Cfg) else v

    def __setattr__(self, k, v):
        self[k] = v


def _instantiate(cfg, *args, **kwargs):
    cfg = dict(cfg)
    target = cfg.pop("_target_")
    kwargs.pop("_recursive_", None)
    cfg.update(kwargs)
    return _locate(target)(*args, **cfg)


def install_stubs():
    import transformers  # noqa: F401  (import before the timm stub exists)
    from . import vit

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    hydra = mod("hydra")
    hydra.utils = mod("hydra.utils", instantiate=_instantiate, call=_instantiate)
    hydra.types = mod("hydra.types", TargetConf=dict)
    mod("omegaconf", OmegaConf=_Cfg)
    mod("submitit")
    mod("pretrainedmodels", bninception=lambda *a, **k: None)
    mod("timm", create_model=vit.create_model)
    if REF not in sys.path:
        sys.path.insert(0, REF)


def model_cfg(model_type="vit_base_patch16_224", backbone_dim=768, dropout=0.2, head=None):
    """The model sub-config of expts/01_ek100_avt.txt over conf/config.yaml defaults."""
    hk = dict(n_head=4, n_layer=6, output_len=1, inter_dim=2048, return_past_too=True,
              future_pred_loss={"_target_": "torch.nn.MSELoss"}, future_pred_loss_wt=1.0, avg_last_n=1)
    hk.update(head or {})
    return _Cfg(
        backbone={"_target_": "models.video_classification.TIMMModel", "model_type": model_type},
        backbone_last_n_modules_to_drop=0, backbone_dim=backbone_dim, intermediate_featdim=None,
        temporal_aggregator={"_target_": "models.temporal_aggregation.Identity"}, same_temp_agg_dim=False,
        future_predictor=dict({"_target_": "models.future_prediction.AVTh"}, **hk),
        project_dim_for_nce=None,
        temporal_aggregator_after_future_pred={"_target_": "models.temporal_aggregation.Identity"},
        dropout=dropout, classifier={"_target_": "torch.nn.Linear"}, use_cls_mappings=False,
        add_regression_head=False, bn=_Cfg(eps=1e-3, mom=0.1), classifier_on_past=True)


def build_reference_model(num_classes=3806, **kw):
    install_stubs()
    from models.base_model import BaseModel
    return BaseModel(model_cfg(**kw), {"action": num_classes}, {})


def build_reference_avth(in_features, **kw):
    install_stubs()
    from models.future_prediction import AVTh
    return AVTh(in_features, **kw)
