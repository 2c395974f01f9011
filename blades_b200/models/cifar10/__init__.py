from .cct import CCTNet  # noqa: F401
