"""Stub: the reference imports matplotlib.pyplot only for SQdata.plot()."""
