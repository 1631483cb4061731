"""Host-side pieces of bench.py that need no GPU: the one-poller-per-job clock sampler (against a stand-in NVML) and the config dict
both arms print."""
import importlib
import sys
import time
import types

import pytest


@pytest.fixture()
def bench(monkeypatch):
    fake = types.ModuleType("pynvml")
    fake.NVML_CLOCK_SM = 1
    fake.nvmlClocksEventReasonSwPowerCap = 0x4
    fake.nvmlClocksEventReasonHwSlowdown = 0x8
    fake.nvmlClocksEventReasonSwThermalSlowdown = 0x20
    fake.nvmlClocksEventReasonHwThermalSlowdown = 0x40
    fake.nvmlClocksEventReasonHwPowerBrakeSlowdown = 0x80
    fake.calls = []
    fake.nvmlInit = lambda: None
    fake.nvmlDeviceGetHandleByIndex = lambda i: ("gpu", i)
    fake.nvmlDeviceGetMaxClockInfo = lambda h, k: 1965
    fake.nvmlDeviceGetClockInfo = lambda h, k: 1965 - 15 * h[1]
    fake.nvmlDeviceGetCurrentClocksEventReasons = lambda h: 0x4 if h[1] == 2 else 0
    monkeypatch.setitem(sys.modules, "pynvml", fake)
    sys.modules.pop("bench", None)
    return importlib.import_module("bench")


def test_rank0_watches_every_gpu_and_other_ranks_do_not_poll(bench, monkeypatch):
    monkeypatch.delenv("CUDA_VISIBLE_DEVICES", raising=False)
    s = bench.ClockSampler([0, 1, 2, 3])
    assert s.ok and len(s.handles) == 4
    t0 = time.perf_counter()
    s.start()
    time.sleep(0.05)
    s.stop_flag = True
    s.join(timeout=1)
    out = s.summary(t0, time.perf_counter())
    assert out["samples"] >= 3 and out["gpus_watched"] == 4
    assert out["sm_mhz_per_gpu"] == [1965.0, 1950.0, 1935.0, 1920.0] and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]          # GPU 2's reason is reported for the job
    idle = bench.ClockSampler([])
    idle.start(); idle.join(timeout=1)
    assert not idle.ok and idle.summary(0, 1)["sm_mhz"] is None


def test_nvml_index_follows_cuda_visible_devices(bench, monkeypatch):
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "4,5,7")
    assert [bench._nvml_index(i) for i in range(4)] == [4, 5, 7, 3]
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "GPU-8f6d")
    assert bench._nvml_index(1) == 1
    s = bench.ClockSampler([0])
    assert s.ok


def test_both_arms_print_the_same_config(bench):
    """The driver compares the `config` of the two arms: one function makes it, keyed only by the GPU count and the series."""
    for n in (1, 2, 8):
        for series in ("weak", "strong"):
            c = bench.config_dict(n, series)
            assert c == bench.config_dict(n, series) and c["n_gpus"] == n and c["scaling"] == series
            assert "workload" in c and "model" not in c
    assert bench.config_dict(1, "weak")["rows"] == 100_000_000
