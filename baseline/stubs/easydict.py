"""Minimal stand-in for the `easydict` package (absent from this image; the reference imports it in config.py:5 and
main.py:10 for an attribute-access dict). Test / benchmark infrastructure only."""


class EasyDict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__
