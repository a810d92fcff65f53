"""CLI-level smoke tests on the CPU backend: the benchmark harness (C15), the
smoke drivers (C18) and the trainer example (C17) - the reference's CI only
checked "no exception" (run_ci_examples.sh); here outputs are asserted."""
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=280):
    env = dict(os.environ, PYTHONPATH=ROOT)
    res = subprocess.run([sys.executable] + args, cwd=ROOT, env=env, timeout=timeout,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout[-3000:]
    return res.stdout


@pytest.mark.timeout(300)
def test_benchmark_cli_writes_reference_csvs(tmp_path):
    out = _run(["benchmarks/benchmark.py", "--num-rows", "40000", "--num-files", "4",
                "--num-row-groups-per-file", "2", "--num-reducers", "4", "--num-trainers", "2",
                "--num-epochs", "3", "--max-concurrent-epochs", "2", "--batch-size", "1000",
                "--num-trials", "2", "--data-dir", str(tmp_path / "d"), "--stats-dir",
                str(tmp_path / "s"), "--backend", "cpu", "--quiet", "--seed", "3",
                "--utilization-sample-period", "0.05"])
    assert "Mean throughput over 2 trials" in out
    trial = pd.read_csv(tmp_path / "s" / "trial_stats_40K_rows_1K_batch_size.csv")
    epoch = pd.read_csv(tmp_path / "s" / "epoch_stats_40K_rows_1K_batch_size.csv")
    assert len(trial) == 2 and len(epoch) == 6
    assert list(trial.columns[:7]) == ["num_files", "num_row_groups_per_file", "num_reducers",
                                      "num_trainers", "num_epochs", "max_concurrent_epochs",
                                      "trial"]
    assert np.allclose(trial["row_throughput"], 3 * 40000 / trial["duration"])
    assert np.allclose(trial["batch_throughput_per_trainer"], trial["batch_throughput"] / 2)
    # validation errors of the reference CLI are kept
    env = dict(os.environ, PYTHONPATH=ROOT)
    bad = subprocess.run([sys.executable, "benchmarks/benchmark.py", "--num-trials", "1",
                          "--trials-timeout", "1"], cwd=ROOT, env=env, capture_output=True,
                         text=True, timeout=120)
    assert bad.returncode != 0 and "Only one of --num-trials" in bad.stderr


@pytest.mark.timeout(300)
def test_example_trains_and_reports_wait_times(tmp_path):
    out = _run(["examples/ddp/torch_shuffle.py", "--no-cuda", "--num-rows", "12000",
                "--num-files", "2", "--num-columns", "8", "--batch-size", "1500", "--epochs", "2",
                "--num-reducers", "2", "--data-dir", str(tmp_path), "--log-interval", "4"])
    assert "Mean batch wait time" in out and "Done consuming batches." in out
    assert out.count("stats over 8 steps") == 2       # 12000 / 1500 batches per epoch


@pytest.mark.timeout(300)
def test_self_launching_example_with_reference_flags(tmp_path):
    """examples/horovod/ray_torch_shuffle.py: the reference's CLI, spawning its
    own workers (2 gloo ranks here) like the RayExecutor launcher did."""
    out = _run(["examples/horovod/ray_torch_shuffle.py", "--num-workers", "2", "--no-cuda",
                "--num-rows", "16000", "--num-files", "4", "--num-columns", "8",
                "--batch-size", "2000", "--epochs", "2", "--num-reducers", "4",
                "--mock-train-step-time", "0.001", "--cpus-per-worker", "2",
                "--data-dir", str(tmp_path)])
    assert "--cpus-per-worker only apply to Ray/Horovod" in out
    assert out.count("Done consuming batches on worker") == 2
    assert out.count("stats over 4 steps") == 4       # 2 workers x 2 epochs, 8000 rows each
    bad = subprocess.run([sys.executable, "examples/horovod/ray_torch_shuffle.py",
                          "--num-workers", "2", "--num-hosts", "1"], cwd=ROOT,
                         env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True,
                         timeout=120)
    assert bad.returncode != 0 and "either --num-workers or" in bad.stderr


def test_models_forward_backward():
    import torch
    from ray_shuffling_data_loader_b200 import models
    from ray_shuffling_data_loader_b200.data_generation import DATA_SPEC
    mlp = models.TabularMLP(15, hidden=(32, 16))
    x = torch.rand(64, 15)
    mlp(x).sum().backward()
    card = {c: hi for c, (lo, hi, dt) in DATA_SPEC.items() if c != "labels"}
    net = models.EmbeddingTabularNet(card, embedding_dim=4, hidden=(16,), max_rows=1000)
    feats = [torch.randint(0, hi, (32, 1)) for hi in card.values()]
    out = net(feats)
    assert out.shape == (32, 1)
    out.sum().backward()
    cnn = models.SmallConvNet()
    assert cnn(torch.rand(4, 1, 28, 28)).shape == (4, 10)


def test_native_numa_helpers():
    """The C++ runtime's NUMA placement helpers work without a GPU."""
    import os
    from ray_shuffling_data_loader_b200 import _C
    assert _C.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert _C.parse_cpulist("") == []
    assert _C.gpu_numa_node(0) == -1 or _C.gpu_numa_node(0) >= 0
    before = os.sched_getaffinity(0)
    assert _C.bind_thread_to_numa_node(-1) == 0          # unknown node: untouched
    assert os.sched_getaffinity(0) == before
    cpus = _C.numa_node_cpus(0)
    if cpus:
        n = _C.bind_thread_to_numa_node(0)
        after = os.sched_getaffinity(0)
        assert after <= before                            # never widens the mask
        assert n == 0 or n == len(after)
        os.sched_setaffinity(0, before)


def test_bench_effective_counts_and_exactly_once():
    import bench
    # the driver's flags (--steps 20 --warmup 5) with 50 batches per epoch and window 2
    assert bench.effective_counts(20, 5, 50, 2) == (2, 1)
    assert bench.effective_counts(20, 5, 50, 2, min_timed_epochs=20) == (2, 20)
    assert bench.effective_counts(200, 100, 50, 2) == (2, 4)
    assert bench.effective_counts(120, 260, 50, 3) == (6, 3)
    ok = bench.exactly_once([10.0, 10.0 + 1e-12], 10.0)
    assert ok["ok"] and ok["epochs_checked"] == 2
    assert not bench.exactly_once([10.0, 10.5], 10.0)["ok"]
