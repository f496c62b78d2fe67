"""Stub so the real reference's `pyradClasses` imports without matplotlib (oracle harness only)."""
