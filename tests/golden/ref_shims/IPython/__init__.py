"""Stub: the reference imports IPython.display for notebook output only."""
