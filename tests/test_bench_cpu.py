"""Host-side pieces of bench.py that need no GPU: the nvidia-smi clock sampler's windowing (the `clocks` object of the
bench line must describe the timed region, not the warm-up, and must never come back empty for a short run)."""
import importlib.util
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class _Proc:
    def terminate(self):
        pass


def _row(sm, cap="Not Active"):
    return [str(sm), "1965", "900.0", "Not Active", "Not Active", "Not Active", cap]


def test_clock_sampler_reports_only_samples_of_the_timed_region():
    b = _bench()
    s = b.ClockSampler(0)
    s.proc = _Proc()
    t = time.monotonic()
    s.rows = [(t - 2.0, _row(1965)), (t - 1.0, _row(1900))]          # warm-up samples
    s.t_mark = t
    s.rows += [(t + 0.1, _row(1700, "Active")), (t + 0.2, _row(1680, "Active")), (t + 0.3, _row(1720, "Active"))]
    out = s.stop()
    assert out["samples"] == 3 and out["sm_mhz"] == 1700 and out["sm_max_mhz"] == 1965
    assert out["reasons"] == ["sw_power_cap"] and "note" not in out


def test_clock_sampler_short_region_falls_back_to_last_warmup_sample():
    b = _bench()
    s = b.ClockSampler(0)
    s.proc = _Proc()
    t = time.monotonic()
    s.rows = [(t - 2.0, _row(1965)), (t - 0.05, _row(1800, "Active"))]
    s.t_mark = t
    out = s.stop()
    assert out["samples"] == 1 and out["sm_mhz"] == 1800 and out["reasons"] == ["sw_power_cap"]
    assert "note" in out


def test_clock_sampler_without_nvidia_smi():
    b = _bench()
    s = b.ClockSampler(0)
    out = s.stop()
    assert out["sm_mhz"] is None and out["reasons"] == ["nvidia-smi unavailable"]
