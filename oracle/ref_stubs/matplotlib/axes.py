class Axes:
    pass
