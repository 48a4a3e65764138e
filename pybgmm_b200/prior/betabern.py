"""Beta(a, b) prior on the inclusion probability of a dimension: the struct of `pybgmm/prior/betabern.py:8-18`."""


class BetaBern(object):
    def __init__(self, a, b):
        self.name = 'Beta'
        assert a >= 0, "a must larger or equal to 0"
        self.a = a
        self.b = b
