from .engine import Engine, Events, State  # noqa: F401
