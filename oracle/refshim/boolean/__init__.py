"""Minimal stand-in for boolean.py 4.0 -- TEST INFRASTRUCTURE ONLY.

The reference pins ``boolean.py==4.0`` (``/root/reference/requirements.txt``),
absent from this image.  Only the surface used by
``cyberbattle/simulation/actions.py:111,158-171`` and ``model.py:44,217-223``
is provided: ``BooleanAlgebra().parse``, ``Expression.get_symbols / subs /
simplify / __eq__``.  Importable both as ``import boolean`` and
``from boolean import boolean``.
"""
from . import boolean  # noqa: F401
from .boolean import BooleanAlgebra, Expression, Symbol  # noqa: F401
