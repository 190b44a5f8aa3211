def mean_squared_error(a, b):
    raise NotImplementedError("shim")
