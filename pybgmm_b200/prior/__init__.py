from .niw import NIW
from .betabern import BetaBern

__all__ = ["NIW", "BetaBern"]
