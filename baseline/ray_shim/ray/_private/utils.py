"""``ray._private.utils.get_num_cpus`` (reference dataset.py:7)."""
import os


def get_num_cpus():
    return os.cpu_count() or 1
