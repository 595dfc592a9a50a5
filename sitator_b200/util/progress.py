"""Progress-bar switch with the reference's env-var semantics (``util/progress.py:1-14``)."""
import os

_flag = os.getenv("SITATOR_PROGRESSBAR", "true").lower()
enabled = _flag in ("true", "yes", "on")


def tqdm(iterable, **kwargs):
    if enabled:
        try:
            from tqdm.auto import tqdm as _tqdm
            return _tqdm(iterable, **kwargs)
        except Exception:
            pass
    return iterable
