"""LEAF-style federated data tooling (the reference vendors LEAF's scripts under models/utils/ --
SURVEY M4; unused by its package).  Re-implemented as one importable module (``leaf``) plus thin CLIs."""
