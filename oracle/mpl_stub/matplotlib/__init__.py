"""Minimal matplotlib stand-in so the unmodified reference module imports in an image without
matplotlib (it only calls ``mpl.use('Agg')`` at import, vapor_vali/Simple_function.pyx:6-8).  Test infrastructure."""


def use(*args, **kwargs):
    return None
