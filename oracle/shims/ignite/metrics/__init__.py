from .metric import Metric  # noqa: F401
