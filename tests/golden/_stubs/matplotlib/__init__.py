def use(*a, **k):
    pass
