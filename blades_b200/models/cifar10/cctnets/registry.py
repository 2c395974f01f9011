"""Minimal model registry (reference cctnets/registry.py falls back to this when timm is absent)."""
from .core import MODEL_REGISTRY, register_model  # noqa: F401
