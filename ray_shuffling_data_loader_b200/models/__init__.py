"""Trainer-side models used by the examples and the smoke test.

The reference's example builds a small MNIST CNN and never runs it (its training
step is ``time.sleep``; reference ``examples/horovod/ray_torch_shuffle.py:124-140,
209-218``). These models are actually trained in ``examples/ddp/torch_shuffle.py``.
"""
from ray_shuffling_data_loader_b200.models.tabular import TabularMLP, EmbeddingTabularNet
from ray_shuffling_data_loader_b200.models.vision import SmallConvNet, build_resnet50

__all__ = ["TabularMLP", "EmbeddingTabularNet", "SmallConvNet", "build_resnet50"]
