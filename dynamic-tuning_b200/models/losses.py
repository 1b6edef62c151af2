"""Drop-in for the reference's `models.losses` (imported by main_image.py:37, main_vtab.py:35,
main_video.py:37): `AdaLoss` with the reference's constructor keywords and return value, so the
entry scripts run against this package without copying a reference file.

Reference models/losses.py:17-84: loss = base_criterion(prediction, y) + token_loss_ratio *
token_loss, token_loss = (mean(token_select) - token_target_ratio)^2 [+ token_minimal_weight *
sum(clamp(token_minimal - mean_over_last_dim(token_select), 0))].  The layer_* keywords are accepted
and unused, as in the reference (its layer loss is commented out, :53).  Host-side glue on the
[B, classes] logits and the [B, L, N-1, 1] masks; the model forward / backward are the kernels.
"""
import torch.nn as nn


class AdaLoss(nn.Module):
    def __init__(self, base_criterion, layer_target_ratio=0.5, layer_loss_ratio=2., layer_diverse_ratio=0.1,
                 layer_entropy_weight=0.1, layer_minimal_weight=0., layer_minimal=0., token_target_ratio=0.5,
                 token_loss_ratio=2., token_minimal=0.1, token_minimal_weight=1.):
        super().__init__()
        self.base_criterion = base_criterion
        self.token_target_ratio = token_target_ratio
        self.token_loss_ratio = token_loss_ratio
        self.token_minimal = token_minimal
        self.token_minimal_weight = token_minimal_weight

    def forward(self, outputs, y):
        x, token_select = outputs["prediction"], outputs["token_select"]
        base_loss = self.base_criterion(x, y)
        token_loss = self.token_loss_ratio * self._get_token_loss(x, token_select)
        return base_loss + token_loss, dict(base_loss=base_loss, token_loss=token_loss)

    def _get_token_loss(self, x, token_select):
        if token_select is None:
            return x.new_zeros(1).mean()
        loss = ((token_select.mean() - self.token_target_ratio) ** 2).mean()
        if self.token_minimal_weight > 0:
            loss = loss + self.token_minimal_weight * (
                self.token_minimal - token_select.mean(-1)).clamp(min=0.).sum()
        return loss
