class DummyVecEnv(object):
    pass
