"""`keras.random` stand-in: draws come from the queue filled by the fixture generator (see _core.push_draw)."""
from refshim_core import pop_draw


def uniform(shape, minval=0.0, maxval=1.0, dtype=None, seed=None):
    assert minval == 0.0 and maxval == 1.0
    return pop_draw(shape)
