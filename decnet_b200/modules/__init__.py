"""Host-side mirror of the reference's module API for the hot path.

Same class names, argument order and meaning as
  modules/SparseMatching/modules/SpaMat.py:12-28, functions/SpaMat.py:8-50
  modules/SparseVar/modules/SpaVar.py:12-28,     functions/SpaVar.py:8-52
"""
from .SpaMat import SpaMat, SpaMatFunction
from .SpaVar import SpaVar, SpaVarFunction

__all__ = ["SpaMat", "SpaMatFunction", "SpaVar", "SpaVarFunction"]
