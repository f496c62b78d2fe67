"""Stub shadowing the reference's menu module, whose import never returns
(reference pyradInteractive.py:761-762 runs `while True: menuMain()` at import)."""
