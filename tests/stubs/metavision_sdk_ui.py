"""TEST STUB (see metavision_sdk_base.py)."""


class BaseWindow:
    class RenderMode:
        BGR = 0


class MTWindow:
    def __init__(self, *a, **k):
        self.frames = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def show_async(self, frame):
        self.frames.append(frame)

    def should_close(self):
        return False

    def set_keyboard_callback(self, cb):
        pass


class UIAction:
    RELEASE = 0


class UIKeyEvent:
    KEY_ESCAPE, KEY_Q, KEY_E, KEY_S = range(4)


class EventLoop:
    @staticmethod
    def poll_and_dispatch():
        pass
