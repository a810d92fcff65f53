"""Vision models for image-shaped list columns (``feature_shapes=[(C, H, W)]``)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class SmallConvNet(nn.Module):
    """28x28 single-channel classifier (the role of the reference example's
    ``Net``), written for channels-last bf16 execution."""

    def __init__(self, num_classes: int = 10):
        super().__init__()
        self.features = nn.Sequential(
            nn.Conv2d(1, 16, 3, padding=1), nn.ReLU(inplace=True), nn.MaxPool2d(2),
            nn.Conv2d(16, 32, 3, padding=1), nn.ReLU(inplace=True), nn.MaxPool2d(2))
        self.head = nn.Sequential(nn.Flatten(), nn.Linear(32 * 7 * 7, 64),
                                  nn.ReLU(inplace=True), nn.Linear(64, num_classes))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.log_softmax(self.head(self.features(x)), dim=1)


def build_resnet50(num_classes: int = 1000, channels_last: bool = True) -> nn.Module:
    """Random-init torchvision ResNet-50 (BASELINE.json's "ResNet-50 torch
    trainer fed by TorchShufflingDataset" config; no pretrained weights - there
    is no network)."""
    try:
        from torchvision.models import resnet50
    except Exception as e:  # pragma: no cover
        raise RuntimeError("torchvision is required for the ResNet-50 example") from e
    model = resnet50(weights=None, num_classes=num_classes)
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
    return model
