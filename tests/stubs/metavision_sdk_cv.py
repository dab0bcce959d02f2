"""TEST STUB (see metavision_sdk_base.py): identity activity filter."""


class ActivityNoiseFilterAlgorithm:
    def __init__(self, width, height, threshold_us):
        self.args = (width, height, threshold_us)

    def process_events(self, evs, out):
        out._arr = evs.numpy() if hasattr(evs, "numpy") else evs
