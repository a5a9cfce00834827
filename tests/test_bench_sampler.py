"""bench.py's clocks sampler: nvidia-smi CSV lines -> the `clocks` object of the JSON line (no GPU, no nvidia-smi)."""
import datetime
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _row(ts, idx, sm, mx, cap="Not Active"):
    t = datetime.datetime.fromtimestamp(ts).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
    return f"{t}, {idx}, {sm}, {mx}, 512.3, 0x0000000000000000, Not Active, Not Active, Not Active, {cap}".split(", ")


def test_samples_inside_the_timed_region_are_selected():
    B = _bench()
    s = B.ClockSampler(3)
    os.unlink(s.f.name)
    s.t0, s.t1 = 1000.0, 1000.4
    rows = [_row(999.0, 3, 345, 1965), _row(1000.1, 3, 1965, 1965), _row(1000.2, 3, 1950, 1965, "Active"),
            _row(1000.2, 2, 100, 1965), _row(1002.0, 3, 345, 1965), ["garbage"]]
    c = s.summarise(rows)
    assert c["samples"] == 2 and c["window"] == "timed region"
    assert c["sm_mhz"] == 1957.5 and c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"]


def test_falls_back_to_the_warmup_samples_and_says_so():
    B = _bench()
    s = B.ClockSampler(0)
    os.unlink(s.f.name)
    s.t0, s.t1 = 2000.0, 2000.2
    c = s.summarise([_row(1999.0, 0, 1965, 1965), _row(1999.5, 0, 1965, 1965)])
    assert c["samples"] == 2 and c["window"].startswith("warm-up") and c["sm_mhz"] == 1965.0
    assert s.summarise([])["reasons"] == ["no samples"]
