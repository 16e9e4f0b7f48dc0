"""Permissive stand-ins for third-party packages the reference imports at module scope but never touches on the
MultiGrid path (TEST INFRASTRUCTURE ONLY): Box2D / pyglet / shapely / geopandas / imageio (the Box2D car-racing and
bipedal-walker envs, out of scope), pyvirtualdisplay (train.py:24-26 starts a virtual display for screenshots),
treelib (teachDeepRL's RIAC teacher), gym.envs.box2d, the matplotlib sub-modules of the plotting helpers.

`install()` registers a meta-path finder that serves an empty module for every name in STUBBED (and any sub-module
of one); attribute access on such a module yields a do-nothing class that can be subclassed, called and indexed,
so `from Box2D.b2 import polygonShape`, `class Car(gym.envs.box2d.car_dynamics.Car)` etc. import cleanly.  Nothing
here implements behaviour: code that really needs one of these packages still fails, just later and louder."""
import importlib.abc
import importlib.machinery
import sys
import types

STUBBED = ('Box2D', 'pyglet', 'pyvirtualdisplay', 'shapely', 'geopandas', 'imageio', 'treelib', 'gym.envs.box2d',
           'matplotlib.patches', 'matplotlib.colors', 'matplotlib.colorbar', 'matplotlib.backends', 'cv2')


class _Meta(type):
    """Class-level indexing / item assignment (`pyglet.options['debug_gl'] = False`) and attribute access."""

    def __getitem__(cls, k):
        return _Anything()

    def __setitem__(cls, k, v):
        pass

    def __getattr__(cls, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Anything()


class _Anything(object, metaclass=_Meta):
    """Class returned for any attribute of a stub module."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Anything()

    def __getitem__(self, k):
        return _Anything()

    def __setitem__(self, k, v):
        pass

    def __iter__(self):
        return iter(())

    def start(self):
        return self

    def stop(self):
        return self


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Meta(name, (_Anything,), {'__module__': self.__name__})


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        for s in STUBBED:
            if fullname == s or fullname.startswith(s + '.'):
                return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def install():
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.append(_Finder())  # last: a really installed package always wins
