class DDPMScheduler:
    pass
