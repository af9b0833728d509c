class MujocoEnv:
    pass
