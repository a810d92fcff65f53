"""Minimal single-node stand-in for the parts of the Ray API that the UNMODIFIED
reference (``baseline/_ref/ray_shuffling_data_loader``) uses.

Why this exists: the reference arm of ``bench.py`` must run the stock reference
code path, but Ray cannot be installed in this image (no wheel, no network; see
DESIGN.md "Reference arm"). This package supplies ``ray.remote / get / wait /
get_actor / kill / init`` with Ray's semantics so that the reference's own
``shuffle``, ``shuffle_map``, ``shuffle_reduce``, ``_QueueActor`` and
``TorchShufflingDataset`` run unchanged:

* tasks run on a pool of worker *processes* (like Ray workers), scheduled by
  the head process once their ObjectRef arguments are ready;
* task outputs live in a shared-memory object store (one pickle-protocol-5 file
  per object under /dev/shm - a put copy and a get copy, roughly what plasma
  costs without its zero-copy read path);
* actors are hosted by the head process, one thread + asyncio loop each, and are
  reachable from every process through a unix-socket RPC (``get_actor`` by name).

It is NOT part of the product and none of the product's code runs on the
reference arm. Compatibility patches applied here (environment, not reference
code): ``np.object`` alias (removed in numpy 1.24; reference
``torch_dataset.py:212``) and ``asyncio.wait`` accepting bare coroutines
(removed in Python 3.11; reference ``batch_queue.py:409-414,425-433``).
"""
from __future__ import annotations

import asyncio
import atexit
import functools
import inspect
import itertools
import multiprocessing as mp
import os
import mmap
import pickle
import struct
import shutil
import sys
import threading
import time
import traceback
import uuid
from multiprocessing.connection import Client, Listener

import numpy as np

from . import exceptions  # noqa: F401
from . import _config     # noqa: F401

__version__ = "0.0.0-shim"

if "object" not in np.__dict__:
    np.object = object      # environment compat, see module docstring


def _patch_asyncio_wait():
    orig = asyncio.wait
    if getattr(orig, "_shim", False):
        return

    async def wait(fs, *a, **kw):
        fs = [asyncio.ensure_future(f) if inspect.iscoroutine(f) else f for f in fs]
        return await orig(fs, *a, **kw)
    wait._shim = True
    asyncio.wait = wait


_patch_asyncio_wait()

_STATE = {"session": None, "role": None, "head": None, "conn_local": threading.local()}
_AUTH = b"ray-shim"


# ---------------------------------------------------------------------------
# object store
# ---------------------------------------------------------------------------
class ObjectRef:
    __slots__ = ("id",)

    def __init__(self, oid=None):
        self.id = oid or uuid.uuid4().hex

    def __reduce__(self):
        return (ObjectRef, (self.id,))

    def __hash__(self):
        return hash(self.id)

    def __eq__(self, other):
        return isinstance(other, ObjectRef) and other.id == self.id

    def __repr__(self):
        return f"ObjectRef({self.id[:8]})"

    # ``await ref`` (used by the reference's async queue API)
    def __await__(self):
        loop = asyncio.get_event_loop()
        return loop.run_in_executor(None, get, self).__await__()


def _session():
    if _STATE["session"] is None:
        raise RuntimeError("ray.init() has not been called")
    return _STATE["session"]


def _path(oid):
    return os.path.join(_session(), "objects", oid)


_MAGIC = b"RSHIM5\0\0"


def _store(oid, value, is_error=False):
    """Object file = header | pickle (protocol 5) | out-of-band buffers, 64-byte
    aligned. Large numpy / pandas blocks travel out of band so that ``_load`` can map
    them instead of copying them - what plasma gives real Ray (zero-copy reads of
    immutable objects)."""
    tmp = _path(oid) + ".tmp" + uuid.uuid4().hex[:6]
    bufs = []
    try:
        payload = pickle.dumps((is_error, value), protocol=5, buffer_callback=bufs.append)
        raws = [b.raw() for b in bufs]
    except (BufferError, ValueError, TypeError):
        payload, raws = pickle.dumps((is_error, value), protocol=5), []   # in-band fallback
    with open(tmp, "wb") as f:
        f.write(_MAGIC + struct.pack("<QQ", len(payload), len(raws))
                + b"".join(struct.pack("<Q", r.nbytes) for r in raws))
        f.write(payload)
        pos = f.tell()
        for r in raws:
            pad = (-pos) % 64
            f.write(b"\0" * pad)
            f.write(r)
            pos += pad + r.nbytes
    os.replace(tmp, _path(oid))


def _read_object(path):
    """-> (is_error, value); buffers are copy-on-write views of the mapped file."""
    with open(path, "rb") as f:
        size = os.fstat(f.fileno()).st_size
        mm = mmap.mmap(f.fileno(), size, access=mmap.ACCESS_COPY) if size else None
    view = memoryview(mm)
    if bytes(view[:8]) != _MAGIC:
        raise ValueError(f"{path}: not a ray_shim object")
    plen, nbuf = struct.unpack_from("<QQ", view, 8)
    sizes = struct.unpack_from(f"<{nbuf}Q", view, 24) if nbuf else ()
    pos = 24 + 8 * nbuf
    payload = view[pos:pos + plen]
    pos += plen
    buffers = []
    for n in sizes:
        pos += (-pos) % 64
        buffers.append(view[pos:pos + n])
        pos += n
    return pickle.loads(payload, buffers=buffers)


def _ready(oid):
    return os.path.exists(_path(oid))


_DELETE_ON_CONSUME = os.environ.get("RAY_SHIM_DELETE_ON_CONSUME", "1") == "1"


def _load(oid, consume=False):
    is_error, value = _read_object(_path(oid))      # (the mapping outlives the unlink below)
    if consume and _DELETE_ON_CONSUME:
        # No distributed ref-counting here: objects in this workload have exactly
        # one consumer (mapper partition -> its reducer, reducer output -> its
        # trainer), so the store frees an object once it has been fetched.
        try:
            os.unlink(_path(oid))
        except OSError:
            pass
    if is_error:
        raise value
    return value


def put(value):
    ref = ObjectRef()
    _store(ref.id, value)
    return ref


def get(refs, timeout=None, _consume=True):
    single = isinstance(refs, ObjectRef)
    lst = [refs] if single else list(refs)
    deadline = None if timeout is None else time.monotonic() + timeout
    out = []
    for r in lst:
        sleep = 0.0002
        while not _ready(r.id):
            if deadline is not None and time.monotonic() > deadline:
                raise exceptions.GetTimeoutError("Get timed out")
            time.sleep(sleep)
            sleep = min(sleep * 1.5, 0.001)
        out.append(_load(r.id, consume=_consume))
    return out[0] if single else out


def wait(refs, num_returns=1, timeout=None, fetch_local=True):
    refs = list(refs)
    deadline = None if timeout is None else time.monotonic() + timeout
    sleep = 0.0002
    while True:
        ready = [r for r in refs if _ready(r.id)]
        if len(ready) >= num_returns or (deadline is not None and time.monotonic() > deadline):
            ready = ready[:num_returns]
            rs = set(ready)
            return ready, [r for r in refs if r not in rs]
        time.sleep(sleep)
        sleep = min(sleep * 1.5, 0.001)


# ---------------------------------------------------------------------------
# RPC to the head
# ---------------------------------------------------------------------------
def _head_call(msg):
    if _STATE["role"] == "head":
        return _STATE["head"].handle(msg)
    loc = _STATE["conn_local"]
    conn = getattr(loc, "conn", None)
    if conn is None:
        conn = Client(os.path.join(_session(), "head.sock"), family="AF_UNIX", authkey=_AUTH)
        loc.conn = conn
    conn.send(msg)
    status, value = conn.recv()
    if status == "err":
        raise value
    return value


# ---------------------------------------------------------------------------
# tasks
# ---------------------------------------------------------------------------
def _resolve_args(args, kwargs):
    args = [get(a) if isinstance(a, ObjectRef) else a for a in args]
    kwargs = {k: get(v) if isinstance(v, ObjectRef) else v for k, v in kwargs.items()}
    return args, kwargs


def _run_task(fn_blob, args, kwargs, out_ids):
    try:
        fn = pickle.loads(fn_blob)
        if isinstance(fn, RemoteFunction):   # module-level @ray.remote functions
            fn = fn._fn
        args, kwargs = _resolve_args(args, kwargs)
        result = fn(*args, **kwargs)
        if len(out_ids) == 1:
            _store(out_ids[0], result)
        else:
            result = list(result)
            assert len(result) == len(out_ids), "num_returns mismatch"
            for oid, val in zip(out_ids, result):
                _store(oid, val)
    except BaseException as e:  # noqa
        err = exceptions.RayTaskError(f"{type(e).__name__}: {e}\n{traceback.format_exc()}")
        try:
            pickle.dumps(e)
            err = e
        except Exception:
            pass
        for oid in out_ids:
            _store(oid, err, is_error=True)


def _worker_main(session, task_q):
    _STATE["session"] = session
    _STATE["role"] = "worker"
    parent = os.getppid()

    def _watch_parent():        # never outlive the head process
        while True:
            time.sleep(1.0)
            if os.getppid() != parent:
                os._exit(0)
    threading.Thread(target=_watch_parent, daemon=True).start()
    while True:
        item = task_q.get()
        if item is None:
            return
        _run_task(*item)


class RemoteFunction:
    def __init__(self, fn, options=None):
        self._fn = fn
        self._options = dict(options or {})
        functools.update_wrapper(self, fn)

    def options(self, **opts):
        return RemoteFunction(self._fn, {**self._options, **opts})

    def remote(self, *args, **kwargs):
        n = int(self._options.get("num_returns", 1))
        refs = [ObjectRef() for _ in range(n)]
        blob = _pickle_fn(self._fn)
        _head_call(("task", blob, list(args), kwargs, [r.id for r in refs]))
        return refs[0] if n == 1 else refs

    def __call__(self, *a, **k):
        raise TypeError("Remote functions cannot be called directly; use .remote()")


def _pickle_fn(fn):
    try:
        return pickle.dumps(fn)
    except Exception:
        import cloudpickle
        return cloudpickle.dumps(fn)


# ---------------------------------------------------------------------------
# actors
# ---------------------------------------------------------------------------
class _ActorHost:
    def __init__(self, cls, args, kwargs):
        self.loop = asyncio.new_event_loop()
        self.ready = threading.Event()
        self.error = None
        self.dead = False
        self.thread = threading.Thread(target=self._run, args=(cls, args, kwargs), daemon=True)
        self.thread.start()
        self.ready.wait()
        if self.error is not None:
            raise self.error

    def _run(self, cls, args, kwargs):
        asyncio.set_event_loop(self.loop)
        try:
            self.instance = cls(*args, **kwargs)
        except BaseException as e:  # noqa
            self.error = e
            self.ready.set()
            return
        self.ready.set()
        self.loop.run_forever()

    def submit(self, method, args, kwargs, oid):
        if self.dead:
            _store(oid, exceptions.RayActorError("actor is dead"), is_error=True)
            return

        async def runner():
            try:
                if method == "__ray_terminate__":
                    self.dead = True
                    _store(oid, None)
                    self.loop.stop()
                    return
                a, k = _resolve_args(args, kwargs)
                res = getattr(self.instance, method)(*a, **k)
                if inspect.isawaitable(res):
                    res = await res
                _store(oid, res)
            except BaseException as e:  # noqa
                try:
                    pickle.dumps(e)
                except Exception:
                    e = exceptions.RayTaskError(repr(e))
                _store(oid, e, is_error=True)
        asyncio.run_coroutine_threadsafe(runner(), self.loop)

    def kill(self):
        self.dead = True
        try:
            self.loop.call_soon_threadsafe(self.loop.stop)
        except Exception:
            pass


class ActorMethod:
    def __init__(self, handle, name):
        self._handle, self._name = handle, name

    def remote(self, *args, **kwargs):
        ref = ObjectRef()
        _head_call(("actor_call", self._handle._actor_id, self._name, list(args), kwargs, ref.id))
        return ref


class ActorHandle:
    def __init__(self, actor_id):
        self._actor_id = actor_id

    def __getattr__(self, name):
        if name.startswith("_") and name != "__ray_terminate__":
            raise AttributeError(name)
        return ActorMethod(self, name)

    def __reduce__(self):
        return (ActorHandle, (self._actor_id,))


class ActorClass:
    def __init__(self, cls, options=None):
        self._cls = cls
        self._options = dict(options or {})

    def options(self, **opts):
        return ActorClass(self._cls, {**self._options, **opts})

    def remote(self, *args, **kwargs):
        actor_id = uuid.uuid4().hex
        _head_call(("actor_create", actor_id, _pickle_fn(self._cls), list(args), kwargs,
                    self._options.get("name")))
        return ActorHandle(actor_id)


def remote(*args, **options):
    if len(args) == 1 and not options and (inspect.isfunction(args[0]) or inspect.isclass(args[0])):
        target = args[0]
        return ActorClass(target) if inspect.isclass(target) else RemoteFunction(target)

    def deco(target):
        return (ActorClass(target, options) if inspect.isclass(target)
                else RemoteFunction(target, options))
    return deco


def get_actor(name):
    aid = _head_call(("get_actor", name))
    if aid is None:
        raise ValueError(f"Failed to look up actor with name '{name}'")
    return ActorHandle(aid)


def kill(actor, no_restart=True):
    _head_call(("actor_kill", actor._actor_id))


# ---------------------------------------------------------------------------
# head: scheduler + actor directory + RPC server
# ---------------------------------------------------------------------------
class _Head:
    def __init__(self, session, num_workers):
        self.session = session
        self.actors = {}
        self.names = {}
        self.lock = threading.Lock()
        self.pending = []
        self.cv = threading.Condition()
        ctx = mp.get_context("fork")
        self.task_q = ctx.Queue()
        self.workers = [ctx.Process(target=_worker_main, args=(session, self.task_q), daemon=True)
                        for _ in range(num_workers)]
        for w in self.workers:
            w.start()
        self.stop = False
        self.listener = Listener(os.path.join(session, "head.sock"), family="AF_UNIX",
                                 authkey=_AUTH)
        threading.Thread(target=self._accept, daemon=True).start()
        threading.Thread(target=self._schedule, daemon=True).start()

    # -- scheduling: dispatch a task once its ObjectRef args exist --------------
    def _deps(self, args, kwargs):
        return [a.id for a in itertools.chain(args, kwargs.values()) if isinstance(a, ObjectRef)]

    def _schedule(self):
        while not self.stop:
            with self.cv:
                if not self.pending:
                    self.cv.wait(0.05)
                batch, self.pending = self.pending, []
            still = []
            for item in batch:
                if all(_ready(d) for d in item[0]):
                    self.task_q.put(item[1])
                else:
                    still.append(item)
            if still:
                with self.cv:
                    self.pending = still + self.pending
                time.sleep(0.0005)

    def handle(self, msg):
        kind = msg[0]
        if kind == "task":
            _, blob, args, kwargs, out_ids = msg
            with self.cv:
                self.pending.append((self._deps(args, kwargs), (blob, args, kwargs, out_ids)))
                self.cv.notify()
            return None
        if kind == "actor_create":
            _, aid, blob, args, kwargs, name = msg
            cls = pickle.loads(blob)
            host = _ActorHost(cls, args, kwargs)
            with self.lock:
                self.actors[aid] = host
                if name:
                    self.names[name] = aid
            return None
        if kind == "actor_call":
            _, aid, method, args, kwargs, oid = msg
            host = self.actors.get(aid)
            if host is None:
                _store(oid, exceptions.RayActorError("unknown actor"), is_error=True)
            else:
                host.submit(method, args, kwargs, oid)
            return None
        if kind == "get_actor":
            return self.names.get(msg[1])
        if kind == "actor_kill":
            host = self.actors.get(msg[1])
            if host:
                host.kill()
            return None
        raise ValueError(kind)

    def _accept(self):
        while not self.stop:
            try:
                conn = self.listener.accept()
            except Exception:
                return
            threading.Thread(target=self._serve, args=(conn,), daemon=True).start()

    def _serve(self, conn):
        try:
            while True:
                try:
                    msg = conn.recv()
                except (EOFError, OSError):
                    return
                try:
                    conn.send(("ok", self.handle(msg)))
                except BaseException as e:  # noqa
                    conn.send(("err", RuntimeError(repr(e))))
        finally:
            conn.close()

    def shutdown(self):
        self.stop = True
        for _ in self.workers:
            self.task_q.put(None)
        for w in self.workers:
            w.join(timeout=2)
            if w.is_alive():
                w.terminate()
        try:
            self.listener.close()
        except Exception:
            pass


def _default_session():
    tag = os.environ.get("RAY_SHIM_SESSION") or f"{os.getuid()}_{os.environ.get('MASTER_PORT', 'local')}"
    root = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    return os.path.join(root, f"rayshim_{tag}")


def is_initialized():
    return _STATE["session"] is not None


def init(address=None, num_cpus=None, object_store_memory=None, resources=None,
         _system_config=None, ignore_reinit_error=True, **_):
    if is_initialized():
        return
    session = _default_session()
    if address in (None, "local"):
        shutil.rmtree(session, ignore_errors=True)
        os.makedirs(os.path.join(session, "objects"))
        _STATE["session"] = session
        _STATE["role"] = "head"
        n = num_cpus or int(os.environ.get("RAY_SHIM_WORKERS", max(2, (os.cpu_count() or 4) - 2)))
        _STATE["head"] = _Head(session, n)
        atexit.register(shutdown)
    else:
        deadline = time.monotonic() + 120
        while not os.path.exists(os.path.join(session, "head.sock")):
            if time.monotonic() > deadline:
                raise ConnectionError("could not find a running ray-shim head")
            time.sleep(0.05)
        _STATE["session"] = session
        _STATE["role"] = "driver"


def shutdown():
    head = _STATE.get("head")
    session = _STATE.get("session")
    if head is not None:
        head.shutdown()
        shutil.rmtree(session, ignore_errors=True)
    _STATE.update(session=None, role=None, head=None)


def object_store_bytes():
    d = os.path.join(_session(), "objects")
    return sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))
