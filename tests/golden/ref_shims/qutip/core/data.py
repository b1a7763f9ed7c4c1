"""Data-layer type tags (the reference only compares ``qobj.dtype`` to them)."""


class CSR:
    pass


class Dia:
    pass


class Dense:
    pass
