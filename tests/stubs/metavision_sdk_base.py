"""TEST STUB of the proprietary Metavision SDK (not installed anywhere this repo runs): just enough
for the reference's unchanged modules to import.  Not product code."""
import numpy as np

EventCD = np.dtype({"names": ["x", "y", "p", "t"], "formats": ["<u2", "<u2", "<i2", "<i8"], "offsets": [0, 2, 4, 8], "itemsize": 16})


class EventCDBuffer:
    def __init__(self, arr=None):
        self._arr = np.zeros(0, dtype=EventCD) if arr is None else arr

    def numpy(self):
        return self._arr
