"""Packaging (component C20). ``pip install -e .`` builds the sm_100a extension
in-tree via ``ray_shuffling_data_loader_b200._build`` (nvcc + g++, no torch
headers needed)."""
from setuptools import find_packages, setup
from setuptools.command.build_py import build_py


class BuildNative(build_py):
    def run(self):
        try:
            from ray_shuffling_data_loader_b200 import _build
            _build.build()
        except Exception as e:  # CPU-only installs still get the numpy backend
            print(f"[setup] native extension not built: {e}")
        super().run()


setup(
    name="ray_shuffling_data_loader_b200",
    version="0.1.0",
    description="A B200-native data loader with pipelined per-epoch shuffling.",
    long_description=(
        "Per-epoch shuffling and loading of Parquet training data for distributed "
        "training: fused TMA scatter kernels over NVLink instead of a Ray "
        "map/reduce shuffle, with the ray_shuffling_data_loader API."),
    install_requires=["numpy", "pandas", "pyarrow", "torch"],
    extras_require={"stats-s3": ["fsspec"], "test": ["pytest", "pytest-timeout"]},
    packages=find_packages(include=["ray_shuffling_data_loader_b200*"]),
    package_data={"ray_shuffling_data_loader_b200": ["csrc/*", "_C*.so"]},
    cmdclass={"build_py": BuildNative},
    python_requires=">=3.10",
)
