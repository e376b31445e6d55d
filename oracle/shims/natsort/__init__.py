"""natsort shim -- TEST INFRASTRUCTURE (oracle): natsorted() with natsort's default algorithm
(unsigned integer chunks compared numerically, text chunks lexically)."""
import re

_NUM = re.compile(r"(\d+)")


def natsort_key(s):
    parts = _NUM.split(str(s))
    return tuple(int(p) if i % 2 else p for i, p in enumerate(parts))


def natsorted(seq, key=None, reverse=False, **kw):
    k = natsort_key if key is None else (lambda x: natsort_key(key(x)))
    return sorted(seq, key=k, reverse=reverse)
