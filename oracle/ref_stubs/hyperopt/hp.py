def _space(kind):
    def make(name, *args):
        return (kind, name) + tuple(args)
    return make


uniform = _space('uniform')
loguniform = _space('loguniform')
quniform = _space('quniform')
qloguniform = _space('qloguniform')
