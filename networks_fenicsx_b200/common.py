"""Timer registry with the interface of ``dolfinx.common.timed`` / ``dolfinx.common.timing`` that
the reference uses to instrument its public methods (mesh.py:29,117,138,425; assembly.py:28,120,
164,328; solver.py:107; network_generation.py:41,157) and that demos/demo_perf.py:85-150 reads."""

from __future__ import annotations

import datetime
import functools
import time

_registry: dict[str, list] = {}


def timed(task: str):
    def decorator(fn):
        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            t0 = time.perf_counter()
            try:
                return fn(*args, **kwargs)
            finally:
                rec = _registry.setdefault(task, [0, 0.0])
                rec[0] += 1
                rec[1] += time.perf_counter() - t0

        return wrapper

    return decorator


def timing(task: str) -> tuple[int, datetime.timedelta]:
    """(call count, accumulated wall time) -- same shape as dolfinx.common.timing."""
    count, total = _registry[task]
    return count, datetime.timedelta(seconds=total)


def list_timings() -> dict[str, tuple[int, float]]:
    return {k: (v[0], v[1]) for k, v in _registry.items()}
