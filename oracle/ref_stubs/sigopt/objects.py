class Assignments:
    pass
