"""BatchQueue: the intent of the reference's 13 queue tests
(ray_shuffling_data_loader/tests/test_batch_queue.py) on the Ray-free queue,
plus the epoch-window back-pressure and by-name connect the reference never tested."""
import asyncio
import os
import threading
import time

import pytest

from ray_shuffling_data_loader_b200.batch_queue import (BatchQueue, Empty, Full,
                                                        QueueActorError,
                                                        connect_queue_actor)


def _q(**kw):
    kw.setdefault("num_epochs", 1)
    kw.setdefault("num_trainers", 1)
    kw.setdefault("max_concurrent_epochs", 1)
    return BatchQueue(**kw)


def test_simple_usage():
    q = _q()
    items = list(range(10))
    for item in items:
        q.put(rank=0, epoch=0, item=item)
    for item in items:
        assert item == q.get(rank=0, epoch=0)


def test_get():
    q = _q()
    q.put(rank=0, epoch=0, item=0)
    assert q.get(rank=0, epoch=0, block=False) == 0
    q.put(rank=0, epoch=0, item=1)
    assert q.get(rank=0, epoch=0, timeout=0.2) == 1
    with pytest.raises(ValueError):
        q.get(rank=0, epoch=0, timeout=-1)
    with pytest.raises(Empty):
        q.get_nowait(rank=0, epoch=0)
    with pytest.raises(Empty):
        q.get(rank=0, epoch=0, timeout=0.2)


def test_get_async():
    async def body():
        q = _q()
        await q.put_async(rank=0, epoch=0, item=0)
        assert await q.get_async(rank=0, epoch=0, block=False) == 0
        await q.put_async(rank=0, epoch=0, item=1)
        assert await q.get_async(rank=0, epoch=0, timeout=0.2) == 1
        with pytest.raises(ValueError):
            await q.get_async(rank=0, epoch=0, timeout=-1)
        with pytest.raises(Empty):
            await q.get_async(rank=0, epoch=0, block=False)
        with pytest.raises(Empty):
            await q.get_async(rank=0, epoch=0, timeout=0.2)
    asyncio.run(body())


def test_put():
    q = _q(maxsize=1)
    q.put(rank=0, epoch=0, item=0, block=False)
    assert q.get(rank=0, epoch=0) == 0
    q.put(rank=0, epoch=0, item=1, timeout=0.2)
    assert q.get(rank=0, epoch=0) == 1
    with pytest.raises(ValueError):
        q.put(rank=0, epoch=0, item=0, timeout=-1)
    q.put(rank=0, epoch=0, item=0)
    with pytest.raises(Full):
        q.put_nowait(rank=0, epoch=0, item=1)
    with pytest.raises(Full):
        q.put(rank=0, epoch=0, item=1, timeout=0.2)
    assert q.full(0, 0) and not q.empty(0, 0)


def test_put_async():
    async def body():
        q = _q(maxsize=1)
        await q.put_async(rank=0, epoch=0, item=0, block=False)
        assert await q.get_async(rank=0, epoch=0) == 0
        await q.put_async(rank=0, epoch=0, item=1, timeout=0.2)
        assert await q.get_async(rank=0, epoch=0) == 1
        with pytest.raises(ValueError):
            await q.put_async(rank=0, epoch=0, item=0, timeout=-1)
        await q.put_async(rank=0, epoch=0, item=0)
        with pytest.raises(Full):
            await q.put_async(rank=0, epoch=0, item=1, block=False)
        with pytest.raises(Full):
            await q.put_async(rank=0, epoch=0, item=1, timeout=0.2)
    asyncio.run(body())


def test_concurrent_get():
    q = _q()
    out = []
    t = threading.Thread(target=lambda: out.append(q.get(rank=0, epoch=0)))
    t.start()
    with pytest.raises(Empty):
        q.get_nowait(rank=0, epoch=0)
    t.join(timeout=0.1)
    assert t.is_alive()          # still blocked
    q.put(rank=0, epoch=0, item=1)
    t.join(timeout=5)
    assert out == [1]


def test_concurrent_put():
    q = _q(maxsize=1)
    q.put(rank=0, epoch=0, item=1)
    t = threading.Thread(target=lambda: q.put(rank=0, epoch=0, item=2))
    t.start()
    with pytest.raises(Full):
        q.put_nowait(rank=0, epoch=0, item=3)
    t.join(timeout=0.1)
    assert t.is_alive()
    assert q.get(rank=0, epoch=0) == 1
    t.join(timeout=5)
    assert q.get(rank=0, epoch=0) == 2


def test_batch():
    q = _q(maxsize=1)
    with pytest.raises(Full):
        q.put_nowait_batch(rank=0, epoch=0, items=[1, 2])
    with pytest.raises(Empty):
        q.get_nowait_batch(rank=0, epoch=0, num_items=1)
    with pytest.raises(TypeError):
        q.put_nowait_batch(rank=0, epoch=0, items=5)
    with pytest.raises(ValueError):
        q.get_nowait_batch(rank=0, epoch=0, num_items=-1)
    with pytest.raises(TypeError):
        q.get_nowait_batch(rank=0, epoch=0, num_items=1.5)
    big = _q(maxsize=100)
    big.put_nowait_batch(rank=0, epoch=0, items=list(range(100)))
    assert big.get_nowait_batch(rank=0, epoch=0, num_items=100) == list(range(100))
    big.put_batch(0, 0, [1, 2, 3])
    assert big.get_nowait_batch(0, 0) == [1, 2, 3]


def test_qsize():
    q = _q()
    items = list(range(10))
    size = 0
    assert q.qsize(rank=0, epoch=0) == size
    for item in items:
        q.put(rank=0, epoch=0, item=item)
        size += 1
        assert q.qsize(rank=0, epoch=0) == size and q.size(0, 0) == size
    assert len(q) == size
    for item in items:
        assert q.get(rank=0, epoch=0) == item
        size -= 1
        assert q.qsize(rank=0, epoch=0) == size


def test_shutdown():
    q = _q()
    blocked = []

    def waiter():
        try:
            q.actor.get(0, 0)
        except QueueActorError as e:
            blocked.append(e)
    t = threading.Thread(target=waiter)
    t.start()
    time.sleep(0.05)
    actor = q.actor
    q.shutdown()
    t.join(timeout=5)
    assert blocked, "blocked getter must be woken with QueueActorError"
    assert q.actor is None
    with pytest.raises(QueueActorError):
        actor.put(0, 0, 1)
    with pytest.raises(QueueActorError):
        q.put(0, 0, 1)


def test_custom_resources():
    # actor_options are accepted for API parity (no Ray scheduler to reserve from)
    q = _q(actor_options={"num_cpus": 1})
    assert q.actor_options == {"num_cpus": 1}
    q.put(0, 0, "x")
    assert q.get(0, 0) == "x"


def test_pulling_streaming():
    """Consumer thread streams get_batch() across epochs while the producer
    dribbles items and producer_done sentinels (reference :231-288)."""
    num_epochs, num_trainers, per_epoch = 5, 1, 6
    q = BatchQueue(num_epochs, num_trainers, max_concurrent_epochs=2, name="stream-test")
    consumed = {e: [] for e in range(num_epochs)}

    def consumer():
        cq = BatchQueue(num_epochs, num_trainers, 2, name="stream-test", connect=True)
        for epoch in range(num_epochs):
            done = False
            while not done:
                items = cq.get_batch(0, epoch)
                if items[-1] is None:
                    done = True
                    items.pop()
                consumed[epoch].extend(items)
                cq.task_done(0, epoch, len(items))
            cq.task_done(0, epoch, 1)
    t = threading.Thread(target=consumer)
    t.start()
    for epoch in range(num_epochs):
        q.new_epoch(epoch)
        for i in range(per_epoch):
            q.put(0, epoch, (epoch, i))
            if i % 2:
                time.sleep(0.005)
        q.producer_done(0, epoch)
    q.wait_until_all_epochs_done()
    t.join(timeout=10)
    assert not t.is_alive()
    for epoch in range(num_epochs):
        assert consumed[epoch] == [(epoch, i) for i in range(per_epoch)]
    q.shutdown()


def test_epoch_window_backpressure():
    """new_epoch blocks while max_concurrent_epochs are in flight until the
    oldest is fully produced AND task_done'd (reference batch_queue.py:395-418)."""
    q = BatchQueue(num_epochs=3, num_trainers=2, max_concurrent_epochs=2)
    q.new_epoch(0)
    q.new_epoch(1)
    for r in range(2):
        q.put(r, 0, "a")
        q.producer_done(r, 0)
    admitted = threading.Event()
    t = threading.Thread(target=lambda: (q.new_epoch(2), admitted.set()))
    t.start()
    assert not admitted.wait(0.15)
    # trainer 0 finishes epoch 0 ("a" + sentinel): still blocked on trainer 1
    assert q.get_batch(0, 0) == ["a", None]
    q.task_done(0, 0, 2)
    assert not admitted.wait(0.15)
    assert q.get_batch(1, 0) == ["a", None]
    q.task_done(1, 0, 2)
    assert admitted.wait(5)
    t.join()
    with pytest.raises(ValueError):
        q.task_done(0, 0, 1)     # over-acknowledging is an error
    with pytest.raises(TimeoutError):
        q.actor.new_epoch(3, timeout=0.05)   # epoch 1 never consumed: clear error, no hang


def test_connect_by_name_other_process(tmp_path, monkeypatch):
    import multiprocessing as mp
    monkeypatch.setenv("RSDL_B200_QUEUE_DIR", str(tmp_path))
    q = BatchQueue(1, 1, 1, name="xproc")
    ctx = mp.get_context("spawn")
    p = ctx.Process(target=_child_put, args=(str(tmp_path),))
    p.start()
    assert q.get(0, 0, timeout=60) == {"hello": [1, 2, 3]}
    p.join(timeout=30)
    assert p.exitcode == 0
    q.shutdown()


def _child_put(qdir):
    import os
    os.environ["RSDL_B200_QUEUE_DIR"] = qdir
    from ray_shuffling_data_loader_b200.batch_queue import BatchQueue
    q = BatchQueue(1, 1, 1, name="xproc", connect=True, connect_backoff_s=0.05)
    q.put(0, 0, {"hello": [1, 2, 3]})
    try:
        q.get(0, 0, timeout=-1)
    except ValueError:
        pass
    else:
        raise SystemExit(3)


def test_connect_retries_then_fails(tmp_path, monkeypatch):
    monkeypatch.setenv("RSDL_B200_QUEUE_DIR", str(tmp_path))
    t0 = time.monotonic()
    with pytest.raises(ValueError, match="Unable to connect"):
        connect_queue_actor("nobody-home", num_retries=3, initial_backoff_s=0.01)
    assert time.monotonic() - t0 >= 0.01 + 0.02 + 0.04 - 1e-3   # exponential back-off


def test_queue_dir_is_private_and_secret_is_not_the_name(tmp_path, monkeypatch):
    """the socket directory must be 0700 and ours, and the handshake key
    must come from a random 0600 secret, not from the public queue name."""
    import stat
    from ray_shuffling_data_loader_b200 import batch_queue as bq
    d = tmp_path / "qdir"
    d.mkdir(mode=0o755)
    monkeypatch.setenv("RSDL_B200_QUEUE_DIR", str(d))
    k1 = bq._authkey("some-queue")
    assert stat.S_IMODE(os.stat(d).st_mode) == 0o700
    sec = d / "secret"
    assert sec.exists() and stat.S_IMODE(os.stat(sec).st_mode) == 0o600
    assert bq._authkey("some-queue") == k1 and bq._authkey("other") != k1
    import hashlib
    assert k1 != hashlib.sha256(b"rsdl-b200:some-queue").digest()
    # a directory owned by somebody else is refused (simulated through os.stat)
    real_stat = os.stat

    def fake_stat(path, *a, **kw):
        st = real_stat(path, *a, **kw)
        if str(path) == str(d):
            vals = list(st)
            vals[4] = st.st_uid + 1
            return os.stat_result(vals)
        return st
    monkeypatch.setattr(bq.os, "stat", fake_stat)
    with pytest.raises(PermissionError):
        bq._socket_path("some-queue")
