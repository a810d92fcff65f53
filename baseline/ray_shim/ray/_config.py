"""``ray._config`` knobs read at import time by the reference's stats.py:648."""


def max_grpc_message_size():
    return 512 * 1024 * 1024
