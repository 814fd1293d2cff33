"""rank_zero_only / rank_zero_info (main_id_embed.py:18-19, ddpm.py:21): rank from the torchrun environment."""
import functools
import os


def _rank():
    return int(os.environ.get("RANK", os.environ.get("LOCAL_RANK", "0")))


def rank_zero_only(fn):
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        if _rank() == 0:
            return fn(*args, **kwargs)
        return None
    return wrapped


@rank_zero_only
def rank_zero_info(*args, **kwargs):
    print(*args, **kwargs)
