from . import seeding  # noqa: F401


class EzPickle(object):
    """gym.utils.EzPickle: pickle by constructor arguments (only subclassed by the Box2D envs, never instantiated here)."""

    def __init__(self, *args, **kwargs):
        self._ezpickle_args, self._ezpickle_kwargs = args, kwargs


def colorize(string, color=None, bold=False, highlight=False):
    return string
