class PowerOfTwoFragment:
    def __init__(self, low, high, default=None):
        self.low, self.high = low, high
        self.default = default if default is not None else low
