def get_policy(s):
    raise NotImplementedError("mixed precision policy is not exercised under the numpy shim")
