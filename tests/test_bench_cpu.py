"""Host-side pieces of bench.py that must not depend on a GPU being present."""
import argparse
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_clock_sampler_degrades_without_nvml_or_nvidia_smi():
    """No driver in the CPU container: prepare/start/stop must still return a well-formed `clocks` object."""
    b = _bench()
    s = b.ClockSampler(0)
    s.prepare()
    s.start()
    clocks = s.stop()
    assert set(clocks) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert isinstance(clocks["reasons"], list)


def test_workload_table_and_synthetic_batch_follow_the_dataset_contract():
    """dataset.py:332-346: data (B,1,S,S,S) f32; label (B,L+1,S,S,S) u8 with the class map last."""
    b = _bench()
    assert b.WORKLOADS["cfg3"]["batch"] == 8 and b.WORKLOADS["cfg3"]["edge"] == 128          # BASELINE.json configs[1]
    wl = dict(b.WORKLOADS["cfg2"], batch=1, edge=16)
    batch = b.synthetic_batch(wl, 0, "cpu")
    assert batch["data"].shape == (1, 1, 16, 16, 16) and str(batch["data"].dtype) == "torch.float32"
    assert batch["label"].shape == (1, 9, 16, 16, 16) and str(batch["label"].dtype) == "torch.uint8"
    assert int(batch["label"][:, -1].max()) < wl["classes"]
    hp = b.hparams_for(wl)
    assert isinstance(hp, argparse.Namespace) and hp.out_channels == 10 and len(hp.loss_regression_weight) == 8
