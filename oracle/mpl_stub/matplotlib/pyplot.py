"""pyplot stand-in: every attribute is a no-op callable (figures are out of scope)."""


def __getattr__(name):
    def _noop(*args, **kwargs):
        return None
    return _noop
