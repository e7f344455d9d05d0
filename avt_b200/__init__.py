"""avt_b200: Blackwell-native (sm_100a) implementation of the AVT training hot path.

Drop-in modules for the reference's Hydra `_target_` slots:
  avt_b200.backbone.TIMMModel           (replaces models.video_classification.TIMMModel)
  avt_b200.future_prediction.AVTh       (replaces models.future_prediction.AVTh)
All arithmetic runs in hand-written CUDA kernels reached through the C-ABI in include/avt_b200.h.
"""
__version__ = "0.1.0"
