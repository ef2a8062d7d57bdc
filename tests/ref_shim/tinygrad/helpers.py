class GlobalCounters:  # python/dump.py:16 imports it; unused on the paths exercised here
    global_ops = 0
    global_mem = 0
