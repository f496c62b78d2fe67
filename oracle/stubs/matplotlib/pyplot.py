"""Stub: the parity harness never plots."""
