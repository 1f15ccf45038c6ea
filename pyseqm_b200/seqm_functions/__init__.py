"""Operator-level interface (SURVEY 8(b) level B) under the reference's module names."""
