class GridSpec:
    def __init__(self, *a, **k):
        pass

    def __getitem__(self, item):
        return None
