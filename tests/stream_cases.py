"""Seeded inputs shared by tests/golden/generate_golden_stream.py (which runs the real reference on
them) and the tests (which run the oracle and the CUDA path on them).  No reference imports here."""
import numpy as np

from oracle import xmaps_oracle as orc

FILTER_CASES = [  # (name, seed, n, p_on)
    ("uniform_20k", 2, 20_000, 1.0),
    ("mixed_polarity_30k", 11, 30_000, 0.5),
    ("dense_collisions", 3, 50_000, 1.0),
    ("tiny", 5, 7, 1.0),
]
TRIGGER_CASES = [  # (name, seed, frames, events per frame, glitch_every, chunk sizes cycled)
    ("steady_1500", 0, 40, 5000, 7, [1500]),
    ("ragged", 1, 60, 3000, 5, [3500, 1900, 4200, 2600, 5000]),
    ("sparse", 2, 30, 1200, 0, [1900, 700]),
    ("glitchy", 3, 50, 4000, 4, [4100, 3900]),
]
FILTER_KEYS = ["first_yt", "first_xy", "last_xy", "mean_xy"]


def filter_inputs(name, seed, n, p_on):
    ev = orc.synth_events(seed, n, 160, 120, p_on=p_on)
    if name == "dense_collisions":
        rng = np.random.default_rng(seed)
        ev["x"] = rng.integers(60, 70, n)
        ev["y"] = rng.integers(50, 58, n)
    ev["t"] += 3_000_000_000  # beyond int32: the reference's int32 images wrap the timestamps
    return ev


def yt_subset(ev, lut_x):
    """Positive events whose rectified x stays inside the reference's (y, x_rect) image: the reference
    indexes it with the raw int16 coordinate, so a value below -(max + 1) raises IndexError
    (frame_event_filter.py:79); values in [-(max + 1), 0) wrap like any negative NumPy index."""
    pos = ev[ev["p"] == 1]
    xp = lut_x[pos["y"], pos["x"]]
    return pos[xp >= -300]


def trigger_stream(seed, frames, per_frame, glitch):
    return orc.synth_projector_stream(seed, frames, per_frame, 160, 120, glitch_every=glitch)


def chunked(stream, chunks):
    i, k = 0, 0
    while i < len(stream):
        c = chunks[k % len(chunks)]
        yield stream[i : i + c]
        i += c
        k += 1
