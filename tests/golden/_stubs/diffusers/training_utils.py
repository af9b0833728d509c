class EMAModel:
    pass
