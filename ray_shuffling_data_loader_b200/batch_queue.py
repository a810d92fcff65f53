"""BatchQueue: per-(epoch, trainer) FIFO queues with an epoch window (C6 + C7).

API-compatible with the reference's ``BatchQueue``/``_QueueActor``
(``ray_shuffling_data_loader/batch_queue.py:24-509``): sync + async + batched
put/get, ``Empty``/``Full``, ``new_epoch`` back-pressure (at most
``max_concurrent_epochs`` epochs in flight; a new one is admitted only when the
oldest has all producers done *and* every item ``task_done``), ``producer_done``
sentinels, ``wait_until_all_epochs_done``, create-or-connect by name with
exponential back-off, and ``shutdown``.

What is different is the substrate. The reference's queue is an asyncio *Ray
actor* reached over gRPC; the data it carries are ObjectRefs into plasma. Here:

* the actor is an in-process, lock-based object (``_QueueActor``) - the control
  plane of one trainer process; in GPU mode the items it carries are
  ``ShuffledChunk`` handles onto slots of the device epoch ring
  (``runtime/device_engine.py``), so, like the reference, the queue moves
  handles, never rows;
* "named actor" lookup is a process-local registry plus a small pickle-RPC
  server on a unix socket (``multiprocessing.connection``), which is what lets
  another process ``connect=True`` by name without Ray's GCS.
"""
from __future__ import annotations

import asyncio
import collections
import hashlib
import logging
import os
import tempfile
import threading
import time
from collections.abc import Iterable
from multiprocessing.connection import Client, Listener
from typing import Any, Dict, List, Optional

logger = logging.getLogger(__name__)


class Empty(Exception):
    pass


class Full(Exception):
    pass


class QueueActorError(RuntimeError):
    """The queue actor is gone (the analogue of ``ray.exceptions.RayActorError``)."""


# TODO: per-item priorities are not supported (neither does the reference).


class _Channel:
    """One bounded FIFO with ``task_done``/``join`` accounting and ``close``."""

    def __init__(self, maxsize: int, cond: threading.Condition):
        self.maxsize = maxsize
        self.items: collections.deque = collections.deque()
        self.unfinished = 0
        self.cond = cond

    def qsize(self) -> int:
        return len(self.items)

    def full(self) -> bool:
        return 0 < self.maxsize <= len(self.items)


class _QueueActor:
    """Thread-safe queue state machine. One instance per ``BatchQueue`` owner."""

    def __init__(self, max_epochs, num_epochs, num_trainers, maxsize,
                 active_ranks=None):
        # Ranks whose queues take part in the epoch window. In distributed mode
        # every process hosts the queue of its own trainer only.
        self.active_ranks = (list(range(num_trainers)) if active_ranks is None
                             else list(active_ranks))
        self.max_epochs = max_epochs
        self.num_epochs = num_epochs
        self.num_trainers = num_trainers
        self.maxsize = maxsize
        self.curr_epochs = collections.deque()
        self._lock = threading.Lock()
        self._cond = threading.Condition(self._lock)
        self.queues = [[_Channel(maxsize, self._cond) for _ in range(num_trainers)]
                       for _ in range(num_epochs)]
        self.queue_producer_done = [[False] * num_trainers for _ in range(num_epochs)]
        self._dead = False

    # -- helpers ----------------------------------------------------------
    def _check(self):
        if self._dead:
            raise QueueActorError("queue actor has been shut down")

    def _wait_for(self, predicate, timeout=None) -> bool:
        """Wait under ``self._cond`` until predicate() or timeout / shutdown."""
        deadline = None if timeout is None else time.monotonic() + timeout
        while True:
            self._check()
            if predicate():
                return True
            remaining = None if deadline is None else deadline - time.monotonic()
            if remaining is not None and remaining <= 0:
                return False
            self._cond.wait(remaining if remaining is not None else 1.0)

    def _epoch_drained(self, epoch) -> bool:
        return all(self.queue_producer_done[epoch][r]
                   and self.queues[epoch][r].unfinished == 0
                   for r in self.active_ranks)

    # -- epoch window (reference batch_queue.py:395-418) -------------------
    def new_epoch(self, epoch: int, timeout: Optional[float] = None):
        """Admit ``epoch``; blocks while ``max_epochs`` epochs are in flight
        until the oldest one is fully produced and fully ``task_done``."""
        with self._cond:
            self._check()
            if len(self.curr_epochs) == self.max_epochs:
                first_epoch = self.curr_epochs[0]
                if not self._wait_for(lambda: self._epoch_drained(first_epoch), timeout):
                    raise TimeoutError(
                        f"epoch {first_epoch} was not consumed within {timeout}s; "
                        "a trainer has stalled or died")
                self.curr_epochs.popleft()
            self.curr_epochs.append(epoch)

    def producer_done(self, rank: int, epoch: int):
        with self._cond:
            self._check()
            ch = self.queues[epoch][rank]
            self._wait_for(lambda: not ch.full())
            ch.items.append(None)
            ch.unfinished += 1
            self.queue_producer_done[epoch][rank] = True
            self._cond.notify_all()

    def wait_until_all_epochs_done(self, timeout: Optional[float] = None):
        last = self.num_epochs - 1
        with self._cond:
            if not self._wait_for(lambda: self._epoch_drained(last), timeout):
                raise TimeoutError("final epoch was not consumed in time")

    # -- introspection ----------------------------------------------------
    def size(self):
        with self._lock:
            return sum(ch.qsize() for chans in self.queues for ch in chans)

    def qsize(self, rank: int, epoch: int):
        with self._lock:
            return self.queues[epoch][rank].qsize()

    def empty(self, rank: int, epoch: int):
        with self._lock:
            return self.queues[epoch][rank].qsize() == 0

    def full(self, rank: int, epoch: int):
        with self._lock:
            return self.queues[epoch][rank].full()

    # -- put / get --------------------------------------------------------
    def put(self, rank: int, epoch: int, item, timeout=None):
        with self._cond:
            ch = self.queues[epoch][rank]
            if not self._wait_for(lambda: not ch.full(), timeout):
                raise Full
            ch.items.append(item)
            ch.unfinished += 1
            self._cond.notify_all()

    def put_batch(self, rank: int, epoch: int, items, timeout=None):
        for item in items:
            self.put(rank, epoch, item, timeout)

    def get(self, rank: int, epoch: int, timeout=None):
        with self._cond:
            ch = self.queues[epoch][rank]
            if not self._wait_for(lambda: len(ch.items) > 0, timeout):
                raise Empty
            item = ch.items.popleft()
            self._cond.notify_all()
            return item

    def get_batch(self, rank: int, epoch: int):
        """Block for one item, then drain whatever else is queued
        (reference batch_queue.py:468-475)."""
        with self._cond:
            ch = self.queues[epoch][rank]
            self._wait_for(lambda: len(ch.items) > 0)
            batch = list(ch.items)
            ch.items.clear()
            self._cond.notify_all()
            return batch

    def put_nowait(self, rank: int, epoch: int, item):
        with self._cond:
            self._check()
            ch = self.queues[epoch][rank]
            if ch.full():
                raise Full
            ch.items.append(item)
            ch.unfinished += 1
            self._cond.notify_all()

    def put_nowait_batch(self, rank: int, epoch: int, items):
        items = list(items)
        with self._cond:
            self._check()
            ch = self.queues[epoch][rank]
            # If maxsize is 0, queue is unbounded, so no need to check size.
            if self.maxsize > 0 and len(items) + ch.qsize() > self.maxsize:
                raise Full(f"Cannot add {len(items)} items to queue of size "
                           f"{ch.qsize()} and maxsize {self.maxsize}.")
            ch.items.extend(items)
            ch.unfinished += len(items)
            self._cond.notify_all()

    def get_nowait(self, rank: int, epoch: int):
        with self._cond:
            self._check()
            ch = self.queues[epoch][rank]
            if not ch.items:
                raise Empty
            item = ch.items.popleft()
            self._cond.notify_all()
            return item

    def get_nowait_batch(self, rank: int, epoch: int, num_items: int = None):
        with self._cond:
            self._check()
            ch = self.queues[epoch][rank]
            if num_items is None:
                # If num_items isn't specified, get all items in the queue.
                num_items = ch.qsize()
            if num_items > ch.qsize():
                raise Empty(f"Cannot get {num_items} items from queue of size "
                            f"{ch.qsize()}.")
            out = [ch.items.popleft() for _ in range(num_items)]
            self._cond.notify_all()
            return out

    def task_done(self, rank: int, epoch: int, num_items: int = 1):
        with self._cond:
            ch = self.queues[epoch][rank]
            if num_items > ch.unfinished:
                raise ValueError("task_done() called too many times")
            ch.unfinished -= num_items
            self._cond.notify_all()

    def ready(self):
        self._check()

    def shutdown(self):
        with self._cond:
            self._dead = True
            self._cond.notify_all()


# ---------------------------------------------------------------------------
# Named-actor directory: process-local registry + unix-socket RPC server
# ---------------------------------------------------------------------------

_REGISTRY: Dict[str, "_ActorHost"] = {}
_REGISTRY_LOCK = threading.Lock()
_RPC_METHODS = {
    "new_epoch", "producer_done", "wait_until_all_epochs_done", "size", "qsize",
    "empty", "full", "put", "put_batch", "get", "get_batch", "put_nowait",
    "put_nowait_batch", "get_nowait", "get_nowait_batch", "task_done", "ready",
    "shutdown",
}


def _queue_root() -> str:
    """Private per-user directory holding the queue sockets and the RPC secret.

    The RPC payloads are pickles, so whoever can serve (or talk to) the socket
    can run code in the peer. The directory therefore must belong to this user
    and be closed to everybody else: it is created 0700 under
    ``$XDG_RUNTIME_DIR`` (else the temp dir) and an existing directory with
    another owner or group/other permission bits is refused rather than trusted.
    ``RSDL_B200_QUEUE_DIR`` overrides the location (same checks apply)."""
    root = os.environ.get("RSDL_B200_QUEUE_DIR")
    if not root:
        base = os.environ.get("XDG_RUNTIME_DIR")
        if not (base and os.path.isdir(base) and os.access(base, os.W_OK)):
            base = tempfile.gettempdir()
        root = os.path.join(base, f"rsdl_b200_{os.getuid()}")
    try:
        os.mkdir(root, 0o700)
    except FileExistsError:
        pass
    except FileNotFoundError:
        os.makedirs(root, mode=0o700, exist_ok=True)
    st = os.stat(root)
    if st.st_uid != os.getuid():
        raise PermissionError(f"queue directory {root} is owned by uid {st.st_uid}, not by this "
                              "user; refusing to use it (set RSDL_B200_QUEUE_DIR)")
    if st.st_mode & 0o077:
        os.chmod(root, 0o700)           # ours but too open (e.g. a mkdtemp'd test dir)
    return root


def _socket_path(name: str) -> str:
    digest = hashlib.sha1(name.encode()).hexdigest()[:16]
    return os.path.join(_queue_root(), f"q_{digest}.sock")


def _authkey(name: str) -> bytes:
    """HMAC key of the connection handshake: a random secret kept in a 0600 file
    inside the private directory (created on first use), mixed with the queue
    name. Knowing the queue name is not enough to talk to the actor."""
    path = os.path.join(_queue_root(), "secret")
    for _ in range(50):
        try:
            with open(path, "rb") as f:
                secret = f.read()
            if len(secret) == 32:
                break
        except FileNotFoundError:
            pass
        try:
            fd = os.open(path + f".{os.getpid()}", os.O_WRONLY | os.O_CREAT | os.O_EXCL, 0o600)
        except FileExistsError:
            time.sleep(0.01)
            continue
        with os.fdopen(fd, "wb") as f:
            f.write(os.urandom(32))
        try:
            os.link(path + f".{os.getpid()}", path)     # atomic: first writer wins
        except FileExistsError:
            pass
        finally:
            os.unlink(path + f".{os.getpid()}")
    else:
        raise RuntimeError(f"could not read or create the queue secret {path}")
    return hashlib.sha256(secret + b"rsdl-b200:" + name.encode()).digest()


class _ActorHost:
    """Owns a ``_QueueActor`` and (when named) serves it to other processes."""

    def __init__(self, actor: _QueueActor, name: Optional[str]):
        self.actor = actor
        self.name = name
        self._listener = None
        self._thread = None
        self._stopping = False
        if name is not None:
            with _REGISTRY_LOCK:
                _REGISTRY[name] = self
            self._serve(name)

    def _serve(self, name: str):
        path = _socket_path(name)
        try:
            if os.path.exists(path):
                os.unlink(path)
            self._listener = Listener(path, family="AF_UNIX", authkey=_authkey(name))
        except OSError as e:  # e.g. read-only tmp: stay process-local
            logger.info("queue %s is process-local only (%s)", name, e)
            self._listener = None
            return
        self._thread = threading.Thread(target=self._accept_loop, daemon=True,
                                        name=f"BatchQueue[{name}]")
        self._thread.start()

    def _accept_loop(self):
        while not self._stopping:
            try:
                conn = self._listener.accept()
            except Exception:
                if self._stopping:
                    return
                continue
            threading.Thread(target=self._conn_loop, args=(conn,), daemon=True).start()

    def _conn_loop(self, conn):
        try:
            while True:
                try:
                    method, args, kwargs = conn.recv()
                except (EOFError, OSError):
                    return
                if method not in _RPC_METHODS:
                    conn.send(("err", AttributeError(method)))
                    continue
                try:
                    conn.send(("ok", getattr(self.actor, method)(*args, **kwargs)))
                except BaseException as e:  # propagate to the caller
                    try:
                        conn.send(("err", e))
                    except Exception:
                        conn.send(("err", RuntimeError(repr(e))))
        finally:
            conn.close()

    def stop(self):
        self._stopping = True
        self.actor.shutdown()
        if self.name is not None:
            with _REGISTRY_LOCK:
                if _REGISTRY.get(self.name) is self:
                    del _REGISTRY[self.name]
        if self._listener is not None:
            try:
                self._listener.close()
            except Exception:
                pass
            try:
                os.unlink(_socket_path(self.name))
            except OSError:
                pass


class _RemoteActor:
    """Client-side stub: one connection per calling thread (calls may block)."""

    def __init__(self, name: str):
        self._name = name
        self._local = threading.local()
        self._conn()  # fail fast if the server is not there

    def _conn(self):
        conn = getattr(self._local, "conn", None)
        if conn is None:
            conn = Client(_socket_path(self._name), family="AF_UNIX",
                          authkey=_authkey(self._name))
            self._local.conn = conn
        return conn

    def __getattr__(self, method):
        if method.startswith("_") or method not in _RPC_METHODS:
            raise AttributeError(method)

        def call(*args, **kwargs):
            try:
                conn = self._conn()
                conn.send((method, args, kwargs))
                status, value = conn.recv()
            except (EOFError, OSError, ConnectionError) as e:
                self._local.conn = None
                raise QueueActorError(f"queue actor {self._name} is unreachable: {e}")
            if status == "err":
                raise value
            return value
        return call


def connect_queue_actor(name, num_retries=5, initial_backoff_s: float = 1.0):
    """
    Connect to the named actor denoted by `name`, retrying up to
    `num_retries` times. Note that the retry uses exponential backoff.
    If max retries is reached without connecting, an exception is raised.
    (Same contract as reference batch_queue.py:358-380.)
    """
    retries = 0
    sleep_dur = initial_backoff_s
    last_exc = None
    while retries < num_retries:
        with _REGISTRY_LOCK:
            host = _REGISTRY.get(name)
        if host is not None:
            return host.actor
        try:
            return _RemoteActor(name)
        except Exception as e:
            retries += 1
            logger.info(
                f"Couldn't connect to queue actor {name}, trying again in "
                f"{sleep_dur} seconds: {retries} / {num_retries}, error: "
                f"{e!s}")
            time.sleep(sleep_dur)
            sleep_dur *= 2
            last_exc = e
    raise ValueError(f"Unable to connect to queue actor {name} after "
                     f"{num_retries} retries. Last error: {last_exc!s}")


class BatchQueue:
    """A first-in, first-out queue between the shuffle engine and trainers.

    The behavior and use cases are similar to those of the asyncio.Queue class.

    Features both sync and async put and get methods.  Provides the option to
    block until space is available when calling put on a full queue,
    or to block until items are available when calling get on an empty queue.

    Optionally supports batched put and get operations.

    Args:
        num_epochs, num_trainers, max_concurrent_epochs: shape of the queue
            grid and the epoch window.
        maxsize (optional, int): maximum size of each queue. If zero, size is
            unbounded.
        name (optional, str): register (or look up) the queue under this name.
        connect (bool): connect to an existing named queue instead of creating.
        actor_options (optional, Dict): accepted for API compatibility with
            the reference (``name`` is honoured; scheduling resources such as
            ``num_cpus`` have no meaning without Ray and are recorded only).
    """

    def __init__(self,
                 num_epochs: int,
                 num_trainers: int,
                 max_concurrent_epochs: int,
                 maxsize: int = 0,
                 name: str = None,
                 connect: bool = False,
                 actor_options: Optional[Dict] = None,
                 connect_retries: int = 5,
                 connect_backoff_s: float = 1.0,
                 active_ranks=None) -> None:
        self._host = None
        if connect:
            assert actor_options is None
            assert name is not None
            self.actor = connect_queue_actor(name, connect_retries, connect_backoff_s)
        else:
            actor_options = dict(actor_options or {})
            if name is not None:
                actor_options["name"] = name
            self.actor_options = actor_options
            actor = _QueueActor(max_concurrent_epochs, num_epochs, num_trainers,
                                maxsize, active_ranks)
            self._host = _ActorHost(actor, actor_options.get("name"))
            self.actor = actor

    def _actor(self):
        if self.actor is None:
            raise QueueActorError("queue has been shut down")
        return self.actor

    def ready(self):
        """Wait until the queue actor is ready."""
        self._actor().ready()

    def new_epoch(self, epoch: int):
        """Prepare the queue for a new epoch, blocking until the queue has
        capacity for a new epoch (in accordance with max_concurrent_epochs)."""
        self._actor().new_epoch(epoch)

    def producer_done(self, rank: int, epoch: int):
        """Signal that the batch producer for (rank, epoch) is done."""
        self._actor().producer_done(rank, epoch)

    def task_done(self, rank: int, epoch: int, num_items: int = 1):
        """Signal that num_items batches are done being processed by the given
        trainer for the provided epoch."""
        self._actor().task_done(rank, epoch, num_items)

    def wait_until_all_epochs_done(self):
        """Block until all batches for all epochs are done being consumed."""
        self._actor().wait_until_all_epochs_done()

    def __len__(self) -> int:
        return self._actor().size()

    def size(self, rank: int, epoch: int) -> int:
        """The size of the queue."""
        return self._actor().qsize(rank, epoch)

    def qsize(self, rank: int, epoch: int) -> int:
        """The size of the queue."""
        return self.size(rank, epoch)

    def empty(self, rank: int, epoch: int) -> bool:
        """Whether the queue is empty."""
        return self._actor().empty(rank, epoch)

    def full(self, rank: int, epoch: int) -> bool:
        """Whether the queue is full."""
        return self._actor().full(rank, epoch)

    @staticmethod
    def _check_timeout(timeout):
        if timeout is not None and timeout < 0:
            raise ValueError("'timeout' must be a non-negative number")

    def put(self, rank: int, epoch: int, item: Any, block: bool = True,
            timeout: Optional[float] = None) -> None:
        """Adds an item to the queue.

        Raises:
            Full: if the queue is full and blocking is False, or it timed out.
            ValueError: if timeout is negative.
        """
        if not block:
            self._actor().put_nowait(rank, epoch, item)
        else:
            self._check_timeout(timeout)
            self._actor().put(rank, epoch, item, timeout)

    def put_batch(self, rank: int, epoch: int, items: Iterable,
                  block: bool = True, timeout: Optional[float] = None) -> None:
        """Adds a list of items to the queue, in order."""
        if not block:
            self._actor().put_nowait_batch(rank, epoch, list(items))
        else:
            self._check_timeout(timeout)
            self._actor().put_batch(rank, epoch, list(items), timeout)

    async def put_async(self, rank: int, epoch: int, item: Any,
                        block: bool = True,
                        timeout: Optional[float] = None) -> None:
        """Async ``put``; blocking waits run on the event loop's executor."""
        if not block:
            self._actor().put_nowait(rank, epoch, item)
        else:
            self._check_timeout(timeout)
            loop = asyncio.get_running_loop()
            await loop.run_in_executor(
                None, lambda: self._actor().put(rank, epoch, item, timeout))

    def get(self, rank: int, epoch: int, block: bool = True,
            timeout: Optional[float] = None) -> Any:
        """Gets an item from the queue.

        Raises:
            Empty: if the queue is empty and blocking is False, or it timed out.
            ValueError: if timeout is negative.
        """
        if not block:
            return self._actor().get_nowait(rank, epoch)
        self._check_timeout(timeout)
        return self._actor().get(rank, epoch, timeout)

    async def get_async(self, rank: int, epoch: int, block: bool = True,
                        timeout: Optional[float] = None) -> Any:
        """Async ``get``."""
        if not block:
            return self._actor().get_nowait(rank, epoch)
        self._check_timeout(timeout)
        loop = asyncio.get_running_loop()
        return await loop.run_in_executor(
            None, lambda: self._actor().get(rank, epoch, timeout))

    def get_batch(self, rank: int, epoch: int) -> Any:
        return self._actor().get_batch(rank, epoch)

    def put_nowait(self, rank: int, epoch: int, item: Any) -> None:
        """Equivalent to put(item, block=False)."""
        return self.put(rank, epoch, item, block=False)

    def put_nowait_batch(self, rank: int, epoch: int, items: Iterable) -> None:
        """Takes in a list of items and puts them into the queue in order.

        Raises:
            Full: if the items will not fit in the queue
        """
        if not isinstance(items, Iterable):
            raise TypeError("Argument 'items' must be an Iterable")
        self._actor().put_nowait_batch(rank, epoch, list(items))

    def get_nowait(self, rank: int, epoch: int) -> Any:
        """Equivalent to get(block=False)."""
        return self.get(rank, epoch, block=False)

    def get_nowait_batch(self, rank: int, epoch: int,
                         num_items: int = None) -> List[Any]:
        """Gets items from the queue and returns them in a list in order.

        Raises:
            Empty: if the queue does not contain the desired number of items
        """
        if num_items is not None:
            if not isinstance(num_items, int):
                raise TypeError("Argument 'num_items' must be an int")
            if num_items < 0:
                raise ValueError("'num_items' must be nonnegative")
        return self._actor().get_nowait_batch(rank, epoch, num_items)

    def shutdown(self, force: bool = False, grace_period_s: int = 5) -> None:
        """Terminates the underlying queue actor: blocked callers and later
        calls raise ``QueueActorError``. ``force``/``grace_period_s`` are kept
        for API parity; termination here is always immediate and clean."""
        if self.actor is not None:
            if self._host is not None:
                self._host.stop()
            else:
                try:
                    self.actor.shutdown()
                except Exception:
                    pass
        self.actor = None
