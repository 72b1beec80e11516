def parallel_api_test(env, num_cycles=10):
    return None
