"""B200-native per-epoch shuffling data loader.

Same public surface as ``ray_shuffling_data_loader`` (reference
``ray_shuffling_data_loader/__init__.py:1-7``): ``TorchShufflingDataset``,
``ShufflingDataset`` and ``shuffle``.
"""
from ray_shuffling_data_loader_b200.torch_dataset import TorchShufflingDataset
from ray_shuffling_data_loader_b200.dataset import ShufflingDataset
from ray_shuffling_data_loader_b200.shuffle import shuffle

__all__ = ["TorchShufflingDataset", "ShufflingDataset", "shuffle"]

__version__ = "0.1.0"
