def host_subplot(*a, **k):
    return None
