def register(id, entry_point, **kw):
    import gym
    gym._registry[id] = entry_point
