"""The reference arm's substrate: the UNMODIFIED reference package
(baseline/_ref, pip-installed from /root/reference) must run on the Ray
stand-in (baseline/ray_shim). Runs in a subprocess so the stand-in `ray` never
leaks into this interpreter. Skipped when baseline/_ref is not installed."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "ray_shuffling_data_loader")

SCRIPT = r"""
import os, sys, tempfile
sys.path.insert(0, os.path.join(ROOT, "baseline", "ray_shim"))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, ROOT)
import numpy as np
import ray
from ray_shuffling_data_loader import TorchShufflingDataset          # the reference
from ray_shuffling_data_loader_b200.data_generation import generate_data, float_spec
d = tempfile.mkdtemp()
files, _ = generate_data(60_000, 3, 2, 0.0, d, data_spec=float_spec(6, np.float32), seed=0)
ray.init(num_cpus=3)
ds = TorchShufflingDataset(files, 3, 1, 7_000, 0, num_reducers=2, max_concurrent_epochs=2,
                           feature_columns=["key"] + [f"f{i}" for i in range(5)],
                           label_column="labels")
for epoch in range(3):
    ds.set_epoch(epoch)
    keys = []
    for data, target in ds:
        assert len(data) == 6 and data[1].shape[1] == 1
        keys.append(data[0][:, 0].long())
    import torch
    keys = torch.cat(keys).numpy()
    assert len(keys) == 60_000 and np.array_equal(np.sort(keys), np.arange(60_000)), len(keys)
print("REFERENCE-OK")
ray.shutdown()
os._exit(0)
"""


@pytest.mark.timeout(280)
@pytest.mark.skipif(not os.path.isdir(REF), reason="baseline/_ref not installed")
def test_unmodified_reference_runs_on_the_shim(tmp_path):
    # the installed reference is byte-identical to the mounted one
    src = "/root/reference/ray_shuffling_data_loader"
    if os.path.isdir(src):
        for name in ("shuffle.py", "dataset.py", "torch_dataset.py", "batch_queue.py", "stats.py"):
            assert open(os.path.join(src, name)).read() == open(os.path.join(REF, name)).read()
    script = tmp_path / "run_ref.py"
    script.write_text(f"ROOT = {ROOT!r}\n" + SCRIPT)
    log = tmp_path / "log.txt"
    with open(log, "w") as f:   # never pipe: worker processes may outlive the parent briefly
        rc = subprocess.run([sys.executable, str(script)], stdout=f, stderr=subprocess.STDOUT,
                            timeout=250).returncode
    out = log.read_text()
    assert rc == 0 and "REFERENCE-OK" in out, out[-3000:]


def test_shim_object_store_roundtrip_is_zero_copy(tmp_path, monkeypatch):
    """The shim's object store writes numpy / pandas blocks out of band and maps them
    back copy-on-write (plasma's zero-copy reads): values round-trip, arrays are
    backed by the mapped file (not owned copies) and stay valid after the object is
    consumed (unlinked)."""
    import os
    import sys
    import numpy as np
    import pandas as pd
    shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "ray_shim")
    monkeypatch.syspath_prepend(shim)
    monkeypatch.setenv("RAY_SHIM_SESSION", f"pytest_store_{os.getpid()}")
    for m in [m for m in sys.modules if m == "ray" or m.startswith("ray.")]:
        monkeypatch.delitem(sys.modules, m)
    import ray
    ray.init(num_cpus=1)
    try:
        df = pd.DataFrame({"a": np.arange(100_000, dtype=np.int64), "b": np.random.rand(100_000)})
        arr = np.arange(1 << 16, dtype=np.float32).reshape(256, 256)
        ref = ray.put({"df": df, "arr": arr, "note": "x"})
        out = ray.get(ref)
        assert out["note"] == "x" and out["df"].equals(df) and np.array_equal(out["arr"], arr)
        assert not out["arr"].flags.owndata            # a view of the mapped object file
        out["arr"][0, 0] = -1.0                        # copy-on-write: private to this reader
        assert out["arr"][0, 0] == -1.0
        assert float(out["df"]["b"].sum()) == float(df["b"].sum())   # still readable after unlink
    finally:
        ray.shutdown()
