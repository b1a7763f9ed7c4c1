"""Errors raised by the package (names as in reference exceptions.py)."""

UNIT_ERROR = ('The input unit is not valid. Look at the API documentation for '
              'the allowed unit formats.')


def raise_unit_error():
    raise ValueError(UNIT_ERROR)


class ModeError(Exception):
    """Feature not available with the active numerical engine."""

    def __init__(self, correct_engine):
        self.correct_engine = correct_engine

    def __str__(self):
        from .settings import get_engine
        want = self.correct_engine
        if isinstance(want, (list, tuple)):
            want = ' or '.join(f"'{w}'" for w in want)
        else:
            want = f"'{want}'"
        return (f"The current operation is not supported by the '{get_engine()}' engine. "
                f'Please set the engine to {want}.')


class CircuitStateError(Exception):
    """The circuit is not in a state that allows the requested call."""


def raise_optim_error_if_needed():
    from .settings import get_optim_mode
    if not get_optim_mode():
        raise ModeError('PyTorch')
