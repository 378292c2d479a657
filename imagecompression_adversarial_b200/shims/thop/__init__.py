def profile(*a, **k):  # coder.py:13 imports it, nothing calls it
    return 0, 0
