from .dnn import MLP  # noqa: F401
