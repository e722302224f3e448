"""Stand-in for opt_einsum.parser (test tooling only)."""
import numpy as np

_symbols = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


def convert_interleaved_input(operands):
    tmp = list(operands)
    out = tmp.pop() if len(tmp) % 2 else None
    tensors = tmp[0::2]
    subs = tmp[1::2]
    keys = []
    for s in subs:
        for k in s:
            if k not in keys:
                keys.append(k)
    if out is not None:
        for k in out:
            if k not in keys:
                keys.append(k)
    try:
        keys_sorted = sorted(keys)
    except TypeError:
        keys_sorted = keys
    m = {k: _symbols[i] for i, k in enumerate(keys_sorted)}
    s = ",".join("".join(m[k] for k in sub) for sub in subs)
    if out is not None:
        s += "->" + "".join(m[k] for k in out)
    return s, tensors
