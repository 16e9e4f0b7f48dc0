class _Registry(object):
    env_specs = {}


registry = _Registry()
