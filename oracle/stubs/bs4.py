"""Stub so the real reference's `pyradUtilities` imports without bs4 (oracle harness only)."""
BeautifulSoup = None
