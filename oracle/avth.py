"""Restatement of the reference AVT head, `models/future_prediction.py:51-258` (class AVTh), for the
configurations every shipped AVT experiment uses (`nn.Linear` encoder; no k-means assignment, no
quantized rollout, drop_last_n == 0). TEST INFRASTRUCTURE.
"""
import torch
import torch.nn as nn

from .gpt2 import GPT2Model


class AVTh(nn.Module):
    def __init__(self, in_features, output_len=-1, output_len_eval=-1, avg_last_n=-1, inter_dim=768,
                 future_pred_loss=None, return_past_too=False, **gpt_kwargs):
        super().__init__()
        self.encoder = nn.Linear(in_features, inter_dim, bias=False)      # future_prediction.py:80
        self.decoder = nn.Linear(inter_dim, in_features, bias=False)      # :81
        gpt_kwargs.pop("future_pred_loss_wt", None)                       # inert kwarg in expts/01:21
        self.gpt_model = GPT2Model(n_embd=inter_dim, **gpt_kwargs)        # :89-95 (wte deleted)
        self.output_len, self.output_len_eval = output_len, output_len_eval
        self.avg_last_n, self.inter_dim, self.in_features = avg_last_n, inter_dim, in_features
        # `future_pred_loss` is a Hydra TargetConf in the reference (:101-105); here: None | "mse"
        self.future_pred_loss = nn.MSELoss(reduction="none") if future_pred_loss else None
        self.return_past_too = return_past_too

    def forward(self, feats, target_shape):
        if feats.ndim == 2:
            feats = feats.unsqueeze(1)                                    # :119-121
        if len(target_shape) == 3:                                        # :123-130
            output_len = target_shape[1]
        elif self.training or self.output_len_eval < 0:
            output_len = self.output_len
        else:
            output_len = self.output_len_eval
        full_orig_feats = inp_feats = feats
        orig_feats_len = feats.size(1)
        feats = self.encoder(feats)                                       # :163
        past, all_outputs, all_outputs_decoded = None, [], []
        for _ in range(output_len):                                       # :168-202
            pred_so_far = sum(el.size(1) for el in all_outputs)
            position_ids = torch.arange(pred_so_far, pred_so_far + feats.size(1), dtype=torch.long,
                                        device=feats.device)
            last_hidden_state, past = self.gpt_model(feats, past, position_ids)
            all_outputs.append(last_hidden_state)
            all_outputs_decoded.append(self.decoder(last_hidden_state))   # :190
            feats = last_hidden_state[:, -1:, :]                          # :202
        all_outputs = torch.cat(all_outputs_decoded, dim=1)               # :227-229 (decoded branch)
        losses = {}
        if self.future_pred_loss is not None:                             # :207-215
            n = min(full_orig_feats.size(1), all_outputs.size(1))
            losses = {"feat": self.future_pred_loss(all_outputs[:, :n - 1], full_orig_feats[:, 1:n])}
        prev = inp_feats
        if self.return_past_too:                                          # :232-236
            final = torch.cat((prev, all_outputs[:, orig_feats_len - 1:, :]), dim=1)
        elif output_len > 0:
            final = all_outputs[:, -output_len:]
        else:
            final = all_outputs
        if self.avg_last_n > 0:                                           # :241-242
            final = torch.mean(final[:, -self.avg_last_n:, :], dim=1)
        updated_past = torch.cat([prev[:, :1, :], all_outputs[:, :orig_feats_len - 1]], dim=1)  # :249-250
        return updated_past, final, losses, {}

    @property
    def output_dim(self):
        return self.in_features
