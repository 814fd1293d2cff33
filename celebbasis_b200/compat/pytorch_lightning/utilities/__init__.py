from .distributed import rank_zero_only, rank_zero_info  # noqa: F401
