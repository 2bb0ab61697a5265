"""bench.py pieces that run without a GPU: the nvidia-smi clock sampler's bookkeeping (samples are attributed to the timed region
by their arrival time; a timed region shorter than one polling period falls back to the warm-up samples and says so)."""
import importlib.util
import os

from helpers import ROOT


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class FakeProc:
    def terminate(self):
        pass

    def wait(self, timeout=None):
        return 0

    def kill(self):
        pass


def row(sm, mx, power_cap="Not Active", thermal="Not Active"):
    return "0, %d, %d, 400.0, 0x0, Not Active, Not Active, %s, %s\n" % (sm, mx, thermal, power_cap)


def sampler(bench, rows, t0, t1):
    s = bench.ClockSampler(0)
    s.p = FakeProc()
    s.rows = rows
    s.t0, s.t1 = t0, t1
    return s.stop()


def test_clock_sampler_uses_the_samples_of_the_timed_region():
    bench = load_bench()
    rows = [(0.5, row(1200, 1965)), (1.1, row(1965, 1965)), (1.2, row(1950, 1965, power_cap="Active")), (9.0, row(300, 1965))]
    c = sampler(bench, rows, 1.0, 1.5)
    assert c["samples"] == 2 and c["sm_mhz"] == 1957.5 and c["sm_max_mhz"] == 1965 and c["reasons"] == ["sw_power_cap"]
    assert c["window"] == "timed region"


def test_clock_sampler_falls_back_to_the_warm_up_for_a_short_region():
    bench = load_bench()
    rows = [(0.5, row(1965, 1965)), (0.9, row(1965, 1965, thermal="Active")), (5.0, row(300, 1965))]
    c = sampler(bench, rows, 1.0, 1.05)
    assert c["samples"] == 2 and c["sm_mhz"] == 1965 and c["reasons"] == ["sw_thermal_slowdown"]
    assert c["window"].startswith("warm-up steps")


def test_clock_sampler_without_nvidia_smi():
    bench = load_bench()
    c = bench.ClockSampler(0).stop()
    assert c["sm_mhz"] is None and c["reasons"] == ["nvidia-smi unavailable"]
