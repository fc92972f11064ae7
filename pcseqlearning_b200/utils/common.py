"""Small host helpers the preprocessors share (mirrors pcdet/utils/common_utils.py:67-78 filter_dict and
pcdet/utils/timer.py; EasyDict falls back to a local equivalent when the `easydict` package is absent)."""
import time

import numpy as np
import torch

try:  # the reference environment has easydict; this image does not
    from easydict import EasyDict  # type: ignore
except ImportError:  # pragma: no cover - exercised in this image

    class EasyDict(dict):
        """dict with attribute access (nested dicts are converted on assignment)."""

        def __init__(self, d=None, **kwargs):
            super().__init__()
            d = dict(d or {})
            d.update(kwargs)
            for k, v in d.items():
                setattr(self, k, v)

        @staticmethod
        def _wrap(v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                return EasyDict(v)
            if isinstance(v, (list, tuple)):
                return type(v)(EasyDict._wrap(x) for x in v)
            return v

        def __setattr__(self, name, value):
            value = EasyDict._wrap(value)
            super().__setattr__(name, value)
            super().__setitem__(name, value)

        __setitem__ = __setattr__

        def update(self, e=None, **f):
            d = dict(e or {})
            d.update(f)
            for k, v in d.items():
                setattr(self, k, v)

        def pop(self, k, *args):
            if k in self.__dict__:
                delattr(self, k)
            return super().pop(k, *args)


def filter_dict(data_dict, mask, ignore_keys=()):
    """Index every entry of a dict with the same mask / index (common_utils.py:67-78)."""
    out = {}
    for key, val in data_dict.items():
        if key in ignore_keys:
            out[key] = val
            continue
        if isinstance(mask, (torch.Tensor, np.ndarray)) and mask.dtype in (torch.bool, np.bool_):
            assert mask.shape[0] == len(val), f"MisMatch for key={key}, mask.shape={mask.shape}, data.shape={len(val)}"
        out[key] = val[mask]
    return out


class Timer:
    """Wall-clock context manager printing elapsed seconds (pcdet/utils/timer.py).  Unlike the reference it
    synchronises the device on both sides when `sync=True`, so the printed time is the work's time."""

    def __init__(self, name="default", sync=False, verbose=True):
        self.name, self.sync, self.verbose = name, sync, verbose

    def __enter__(self):
        if self.sync and torch.cuda.is_available():
            torch.cuda.synchronize()
        self.start = time.time()
        return self

    def __exit__(self, *args):
        if self.sync and torch.cuda.is_available():
            torch.cuda.synchronize()
        self.time_elapsed = time.time() - self.start
        if self.verbose:
            print(f"Elapsed Time in {self.name} = {self.time_elapsed:.4f} s")
