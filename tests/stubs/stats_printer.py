"""TEST STUB: the reference's own stats_printer.py does not import on Python >= 3.11 (it uses a
mutable dataclass default, python/stats_printer.py:180-181).  Minimal stand-in with the interface the
pipe uses (measure_time / add_metric / count / log, SingleTimer).  Not product code."""
import time
from contextlib import contextmanager


class StatsPrinter:
    def __init__(self):
        self.times, self.metrics, self.counts = {}, {}, {}

    @contextmanager
    def measure_time(self, key):
        t0 = time.perf_counter_ns()
        try:
            yield
        finally:
            self.times.setdefault(key, []).append(time.perf_counter_ns() - t0)

    def add_metric(self, key, value):
        self.metrics.setdefault(key, []).append(value)

    def count(self, key, n=1):
        self.counts[key] = self.counts.get(key, 0) + n

    def add_time_measure_ns(self, key, ns):
        self.times.setdefault(key, []).append(ns)

    def log(self, msg):
        pass

    def print_stats(self):
        pass

    def toggle_silence(self):
        pass

    def reset(self):
        self._t0 = time.perf_counter_ns()

    def start_time_ns(self):
        return getattr(self, "_t0", time.perf_counter_ns())


class SingleTimer:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
