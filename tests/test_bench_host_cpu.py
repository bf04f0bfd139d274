"""Host-side pieces of bench.py that need no GPU: the byte model of the roofline and the clock-sample parser."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


class _Counts:
    def __init__(self, ne, nt, nv):
        self.n_elements, self.n_traditional, self.n_vertices = ne, nt, nv


def test_algorithmic_bytes_is_the_survey_model():
    # 304 Ne + 248 Nt + 148 Nv + 28 A; a garment mesh (Ne = 2 Nv) is 252 B per particle
    assert bench.algorithmic_bytes(_Counts(332928, 0, 167040), 0) == 304 * 332928 + 148 * 167040
    assert bench.algorithmic_bytes(_Counts(2, 0, 1), 0) == 252 * 3
    assert bench.algorithmic_bytes(_Counts(0, 10, 0), 7) == 2480 + 196


def test_clock_sampler_parses_nvidia_smi_rows():
    s = bench.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0, "kill": lambda self: None})()
    s.rows = ["1965, 1965, Not Active, Not Active, Not Active, Not Active",
              "1350, 1965, Not Active, Not Active, Not Active, Active",
              "1965, 1965, Not Active, Not Active, Not Active, Not Active",
              "garbage"]
    out = s.stop()
    assert out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 3
    assert out["reasons"] == ["sw_power_cap"]


def test_disabled_sampler_reports_nothing():
    s = bench.ClockSampler(3, enabled=False)
    s.start()
    assert s.stop() is None
