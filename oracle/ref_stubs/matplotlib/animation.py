class FuncAnimation:
    def __init__(self, *a, **k):
        pass
