"""Display helpers (mirror of reference ``iactrace/viz/plotting.py``); matplotlib is optional."""
from .plotting import hexshow, squareshow

__all__ = ["hexshow", "squareshow"]
