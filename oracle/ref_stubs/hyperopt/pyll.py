class Apply:
    pass
