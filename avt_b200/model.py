"""Host-side mirror of the reference glue around the two hot-path modules, for running the AVT end-to-end
configuration (expts/01_ek100_avt.txt) without Hydra: `models/base_model.py:140-220` (forward_singlecrop) with
backbone = avt_b200.backbone.TIMMModel, temporal_aggregator = Identity, future_predictor =
avt_b200.future_prediction.AVTh, classifier = nn.Linear, classifier_on_past = true; plus the loss arithmetic of
`func/train_eval_ops.py:57-85` / `func/train.py:207-217`. Inside the reference itself none of this is needed:
BaseModel instantiates the two modules through their `_target_` paths.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .backbone import TIMMModel
from .future_prediction import AVTh
from .loss_head import FusedClassifierLoss

EXPTS01_HEAD = dict(n_head=4, n_layer=6, output_len=1, inter_dim=2048, return_past_too=True,
                    future_pred_loss={"_target_": "torch.nn.MSELoss"}, future_pred_loss_wt=1.0, avg_last_n=1)


class AVTModel(nn.Module):
    def __init__(self, model_type="vit_base_patch16_224", backbone_dim=768, num_classes=3806, dropout=0.2,
                 head_kwargs=None):
        super().__init__()
        self.backbone = TIMMModel(1, model_type)
        hk = dict(EXPTS01_HEAD)
        hk.update(head_kwargs or {})
        self.future_predictor = AVTh(backbone_dim, **hk)
        self.dropout = nn.Dropout(dropout)
        self.classifiers = nn.ModuleDict({"action": nn.Linear(backbone_dim, num_classes)})
        self._initialize_weights()
        self._loss_head = None

    def _initialize_weights(self):  # models/base_model.py:110-127
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward(self, video, target_shape=None):
        """video (B, #clips, C, T', H, W) or (B, #clips, #crops, C, T', H, W) -> (outputs, aux_losses), each averaged over
        the crops: the test-time-augmentation loop of models/base_model.py:239-273 (3 crops x flip at evaluation,
        common/transforms.py:254-296). Every crop is one full forward; in train mode each keeps its own activation workspace
        until the backward (engine.Lease)."""
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

    def set_precision(self, precision):
        """'bf16' (tensor-core product path) or 'fp32' (inference-only validation path on the CUDA cores, 1e-5 vs the oracle)."""
        assert precision in ("bf16", "fp32")
        self.backbone.model.precision = precision
        self.future_predictor.precision = precision
        return self

    def _features(self, video, target_shape):
        B, num_clips = video.size(0), video.size(1)
        feats = self.backbone(video.flatten(0, 1))                      # base_model.py:153-154
        feats = torch.mean(feats, [-1, -2]).permute((0, 2, 1))          # :157, :166
        feats = feats.reshape((B, num_clips) + feats.shape[1:]).flatten(1, 2)   # :183-191
        return self.future_predictor(feats, target_shape)               # :196-197

    def training_losses(self, video, target, past_target):
        """The training step's forward with the classifier + cross-entropy head fused (avt_b200.loss_head): the same
        losses `forward()` + `training_loss()` give, without materialising the logits dictionaries.
        Returns ({'cls_action', 'past_cls_action', 'feat'}: scalar means as in func/train.py:207-209, {'acc1/action',
        'acc5/action'}: top-k accuracies of the future logits in percent, func/train_eval_ops.py:61-63)."""
        past, future, aux, _ = self._features(video, target.shape)
        if self._loss_head is None or self._loss_head.linear is not self.classifiers["action"]:
            self._loss_head = FusedClassifierLoss(self.classifiers["action"])
        p = self.dropout.p if self.training else 0.0
        lf, lp, acc1, acc5 = self._loss_head(past, future, past_target, target, p)
        losses = {"cls_action": lf, "past_cls_action": lp}
        losses.update({k: torch.mean(v) for k, v in aux.items()})
        return losses, {"acc1/action": acc1, "acc5/action": acc5}

    def attach_loss_head_to(self, optimizer):
        """Let a FlatSGD keep the classifier's bf16 copy current in its own update pass (valid after one training_losses())."""
        head = self._loss_head
        if head is not None and head._wb is not None:
            w = self.classifiers["action"].weight
            optimizer.attach_shadow(w, head._wb[:head.classes].view(-1), head.shadow_written_by_optimizer)

    def forward_singlecrop(self, video, target_shape=None):
        """video (B, #clips=T, C, T'=1, H, W) -> (outputs dict, aux_losses dict)  [models/base_model.py:140-220]"""
        past, future, losses, _ = self._features(video, target_shape)
        out = {"past": past, "future": future}
        out["past_logits/action"] = self.classifiers["action"](self.dropout(past))     # :203-207
        out["logits/action"] = self.classifiers["action"](self.dropout(future))       # :215-216
        return out, losses


def past_targets(target_subclips):
    """Per-frame label of the past-prediction loss: the mode over the sub-clip's frame labels
    (func/train_eval_ops.py:70-75). Label preparation only — kept out of the graph-captured part of the step because
    torch.mode synchronises."""
    return torch.mode(target_subclips, -1)[0]


def training_loss(outputs, aux_losses, target, target_subclips=None, past_tgt=None):
    """sum_k mean(loss_k), weights 1/1/1: CE(future), CE(past vs per-frame mode label, ignore_index -1), MSE feat
    (func/train_eval_ops.py:57-85, loss_fn/multidim_xentropy.py:10-25, func/train.py:207-217, expts/01:1-2)."""
    losses = {"cls_action": F.cross_entropy(outputs["logits/action"], target, ignore_index=-1, reduction="none")}
    if past_tgt is None:
        past_tgt = past_targets(target_subclips)
    pl = outputs["past_logits/action"]
    losses["past_cls_action"] = F.cross_entropy(pl.flatten(0, 1), past_tgt.flatten(), ignore_index=-1,
                                                reduction="none").view(past_tgt.shape)
    losses.update(aux_losses)
    return sum(torch.mean(v) for v in losses.values())


def accuracy(output, target, topk=(1,)):
    """Top-k accuracies in percent, as `common/utils.py:17-44` computes them every training iteration
    (func/train_eval_ops.py:61-63). The reference's early-out `if torch.all(target < 0)` is a device synchronisation; it
    returns zeros, which is also what the general formula gives (a negative label never equals a predicted class), so the
    branch is dropped and the step stays capturable in a CUDA graph."""
    with torch.no_grad():
        output = output.flatten(0, -2)
        target = target.flatten()
        _, pred = output.topk(max(topk), 1, True, True)
        correct = pred.t().eq(target[None])
        return [correct[:k].flatten().sum(dtype=torch.float32) * (100.0 / target.size(0)) for k in topk]
