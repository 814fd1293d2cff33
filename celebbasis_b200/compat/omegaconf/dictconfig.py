from . import DictConfig  # noqa: F401
