def display(*a, **k):
    pass


def Math(x):
    return x
