from theano import SharedVariable  # noqa: F401


class MonitorMode(object):
    def __init__(self, pre_func=None, post_func=None):
        self.pre_func, self.post_func = pre_func, post_func
