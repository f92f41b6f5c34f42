def suggest(*a, **k):
    raise NotImplementedError('hyperopt is not installed in this container')
