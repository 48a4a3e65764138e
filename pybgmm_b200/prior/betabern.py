"""Hyper-parameters of the Beta prior on a dimension's inclusion probability (SubCRPMM): the two fields `a`, `b` and the
tag `name` that `pybgmm/prior/betabern.py:8-18` carries; a negative `a` is refused as there (AssertionError)."""


class BetaBern(object):
    name = 'Beta'

    def __init__(self, a, b):
        if not a >= 0:
            raise AssertionError("a must larger or equal to 0")
        self.a, self.b = a, b

    def mean(self):
        """Prior mean a / (a + b): the inclusion probability SubCRPMM starts from (subcrpmm.py:46-48)."""
        return 1. * self.a / (self.a + self.b)

    def posterior(self, included, excluded):
        """(a, b) of the Beta posterior after `included` ones and `excluded` zeros (subcrpmm.py:261-262)."""
        return self.a + included, self.b + excluded
