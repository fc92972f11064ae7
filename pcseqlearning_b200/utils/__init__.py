from .common import EasyDict, Timer, filter_dict  # noqa: F401
