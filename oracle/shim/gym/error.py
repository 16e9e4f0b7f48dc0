class Error(Exception):
    pass


class RetriesExceededError(Error):
    pass


class DeprecatedEnv(Error):
    pass


class UnregisteredEnv(Error):
    pass
