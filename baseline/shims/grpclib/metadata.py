import time


class Deadline:
    def __init__(self, timestamp: float) -> None:
        self._timestamp = timestamp

    @classmethod
    def from_timeout(cls, timeout: float) -> "Deadline":
        return cls(time.monotonic() + timeout)

    def time_remaining(self) -> float:
        return max(0.0, self._timestamp - time.monotonic())
