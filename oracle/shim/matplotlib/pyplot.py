"""Empty stand-in for matplotlib.pyplot (oracle test infrastructure): every function is a no-op (train.py:232 calls
plt.close() after saving a screenshot with torchvision; envs/multigrid/window.py only plots with --render)."""


def __getattr__(name):
    if name.startswith('__'):
        raise AttributeError(name)

    def _noop(*a, **k):
        return None
    return _noop
