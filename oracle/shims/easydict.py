"""ORACLE SHIM (test infrastructure, never on the product path).

Minimal stand-in for the un-installable `easydict` package: an attribute
dictionary that recursively wraps nested dicts on assignment, which is the only
behaviour the reference relies on (e.g. robust_e_nerf/models/robust_e_nerf.py:302,
323-336 `batch.diff = {}` followed by `batch.diff.ts_diff = ...`).
"""


class EasyDict(dict):
    def __init__(self, d=None, **kwargs):
        super().__init__()
        if d is None:
            d = {}
        if kwargs:
            d = dict(d, **kwargs)
        for k, v in d.items():
            self[k] = v

    @classmethod
    def _wrap(cls, value):
        if isinstance(value, dict) and not isinstance(value, EasyDict):
            return cls(value)
        if isinstance(value, (list, tuple)) and any(
            isinstance(x, dict) for x in value
        ):
            return type(value)(cls._wrap(x) for x in value)
        return value

    def __setitem__(self, key, value):
        super().__setitem__(key, self._wrap(value))

    def __setattr__(self, key, value):
        self[key] = value

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __delattr__(self, key):
        try:
            del self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def update(self, *args, **kwargs):
        for k, v in dict(*args, **kwargs).items():
            self[k] = v
