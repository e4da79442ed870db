from . import classification  # noqa: F401
