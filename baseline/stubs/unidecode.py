"""Minimal stand-in for the `unidecode` package (absent from this image; the reference's torchlight/utils.py:9 imports it
for a string helper the hot path never calls). Test / benchmark infrastructure only."""


def unidecode(s):
    return s
