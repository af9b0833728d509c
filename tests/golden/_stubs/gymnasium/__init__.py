class Env:
    pass
