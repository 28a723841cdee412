"""Record the numpy legacy-RNG draws the reference makes (TEST INFRASTRUCTURE).

The reference draws from the *global* legacy stream through module attributes looked up at
call time (`np.random.random()` examples/treasurehunt/entities.py:73, `np.random.choice`
entities.py:75 and env.py:141, `np.random.randint` models/base_model.py:111,
`np.random.choice` buffers.py:109), so patching the attributes on `numpy.random` is enough.
"""
from __future__ import annotations

import contextlib

import numpy as np


class DrawLog:
    """Flat log of (api, args-summary, result) tuples in call order."""

    def __init__(self):
        self.calls: list[tuple] = []

    def __len__(self):
        return len(self.calls)


@contextlib.contextmanager
def record(log: DrawLog):
    orig_random, orig_choice, orig_randint = np.random.random, np.random.choice, np.random.randint

    def random(*a, **k):
        r = orig_random(*a, **k)
        log.calls.append(("random", None, r))
        return r

    def choice(a, size=None, replace=True, p=None):
        # Object-array choice (spawn kind): draw an index with the identical stream usage
        # (legacy choice with replace=True,p=None is randint(0, pop_size)) so we can log it.
        if isinstance(a, np.ndarray) and a.dtype == object and size is None and replace and p is None:
            idx = orig_choice(len(a))
            log.calls.append(("choice_obj", len(a), int(idx)))
            return a[idx]
        r = orig_choice(a, size=size, replace=replace, p=p)
        log.calls.append(("choice", (a if np.isscalar(a) else len(a), size, replace), np.array(r).copy()))
        return r

    def randint(*a, **k):
        r = orig_randint(*a, **k)
        log.calls.append(("randint", a, r))
        return r

    np.random.random, np.random.choice, np.random.randint = random, choice, randint
    try:
        yield log
    finally:
        np.random.random, np.random.choice, np.random.randint = orig_random, orig_choice, orig_randint
