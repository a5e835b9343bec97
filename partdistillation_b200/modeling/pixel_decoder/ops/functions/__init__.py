"""Operator boundary of the reference's native op (ops/functions/__init__.py): same names."""
from ....functional import MSDeformAttnFunction, ms_deform_attn  # noqa: F401
