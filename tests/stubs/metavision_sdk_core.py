"""TEST STUB (see metavision_sdk_base.py)."""
from metavision_sdk_base import EventCDBuffer


class PolarityFilterAlgorithm:
    def __init__(self, polarity):
        self.polarity = polarity

    @staticmethod
    def get_empty_output_buffer():
        return EventCDBuffer()

    def process_events(self, evs, out):
        arr = evs.numpy() if hasattr(evs, "numpy") else evs
        out._arr = arr[arr["p"] == self.polarity]
