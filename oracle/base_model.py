"""Restatement of the reference glue around the hot path for the AVT end-to-end configuration
(expts/01_ek100_avt.txt): `models/base_model.py:140-220` (BaseModel.forward_singlecrop),
`models/video_classification.py:213-227` (process_each_frame) and the loss arithmetic of
`func/train_eval_ops.py:57-85` + `func/train.py:207-217`. TEST INFRASTRUCTURE (and the CPU baseline of
bench.py, kind "port").
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import vit
from .avth import AVTh


class TIMMModel(nn.Module):
    """models/video_classification.py:249-257 + process_each_frame (:213-227)."""

    def __init__(self, num_classes=1, model_type="vit_base_patch16_224", drop_cls=True):
        super().__init__()
        self.model = vit.create_model(model_type, num_classes=0 if drop_cls else num_classes)

    def forward(self, video):
        B, T = video.size(0), video.size(2)
        flat = video.transpose(1, 2).flatten(0, 1)
        feats = self.model(flat)
        return feats.view((B, T) + feats.shape[1:]).transpose(1, 2).unsqueeze(-1).unsqueeze(-1)


class BaseModel(nn.Module):
    """backbone=avt_b, temporal_aggregator=identity, future_predictor=avth, classifier=linear,
    classifier_on_past=true, single task ('action')."""

    def __init__(self, model_type="vit_base_patch16_224", backbone_dim=768, num_classes=3806, dropout=0.2,
                 head_kwargs=None, backbone=None, future_predictor=None):
        super().__init__()
        self.backbone = backbone if backbone is not None else TIMMModel(1, model_type)
        hk = dict(n_head=4, n_layer=6, output_len=1, inter_dim=2048, return_past_too=True, future_pred_loss="mse",
                  avg_last_n=1)
        hk.update(head_kwargs or {})
        self.future_predictor = future_predictor if future_predictor is not None else AVTh(backbone_dim, **hk)
        self.dropout = nn.Dropout(dropout)
        self.classifiers = nn.ModuleDict({"action": nn.Linear(backbone_dim, num_classes)})
        self._initialize_weights()

    def _initialize_weights(self):  # models/base_model.py:110-127
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward(self, video, target_shape=None):   # models/base_model.py:239-273 (crops averaged)
        if video.ndim == 6:
            crops = [video]
        elif video.ndim == 7 and video.size(2) == 1:
            crops = [video.squeeze(2)]
        elif video.ndim == 7:
            crops = torch.unbind(video, dim=2)
        else:
            raise NotImplementedError("Unsupported size %s" % (video.shape,))
        feats, losses = zip(*[self.forward_singlecrop(c, target_shape) for c in crops])
        feats = {k: torch.mean(torch.stack([d[k] for d in feats], dim=0), dim=0) for k in feats[0]}
        losses = {k: torch.mean(torch.stack([d[k] for d in losses], dim=0), dim=0) for k in losses[0]}
        return feats, losses

    def forward_singlecrop(self, video, target_shape=None):
        B, num_clips = video.size(0), video.size(1)
        feats = self.backbone(video.flatten(0, 1))                 # :153-154   (B*T, C, 1, 1, 1)
        feats = torch.mean(feats, [-1, -2]).permute((0, 2, 1))     # :157,166   (B*T, 1, C)
        feats = feats.reshape((B, num_clips) + feats.shape[1:]).flatten(1, 2)  # :183-191  (B, T, C)
        past, future, losses, _ = self.future_predictor(feats, target_shape)   # :196-197
        out = {"past": past, "future": future}
        out["past_logits/action"] = self.classifiers["action"](self.dropout(past))      # :203-207
        out["logits/action"] = self.classifiers["action"](self.dropout(future))        # :215-216
        return out, losses


def training_loss(outputs, aux_losses, target, target_subclips):
    """CE(future) + CE(past vs per-frame mode label, ignore_index -1) + mean MSE feat, weights 1/1/1
    (func/train_eval_ops.py:57-85, loss_fn/multidim_xentropy.py:10-25, func/train.py:207-217, expts/01:1-2)."""
    losses = {}
    losses["cls_action"] = F.cross_entropy(outputs["logits/action"], target, ignore_index=-1, reduction="none")
    past_tgt = torch.mode(target_subclips, -1)[0]                   # train_eval_ops.py:74-77  (B, T)
    pl = outputs["past_logits/action"]
    losses["past_cls_action"] = F.cross_entropy(pl.flatten(0, 1), past_tgt.flatten(), ignore_index=-1,
                                                reduction="none").view(past_tgt.shape)
    losses.update(aux_losses)
    return sum(torch.mean(v) for v in losses.values())
