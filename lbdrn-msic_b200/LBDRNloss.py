"""Training loss of LBDRN: mean squared error over batch x bands (reference LBDRNloss.py:8-11).
The fused training kernel evaluates the same expression on the device; this module keeps the reference's
class name for callers that build the loss object."""
from torch import nn


class LBDRNLoss(nn.Module):
    def forward(self, y_pred, y):
        return nn.functional.mse_loss(y_pred, y)
