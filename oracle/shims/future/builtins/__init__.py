range = range
