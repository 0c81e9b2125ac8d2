class SparseType(object):
    def __init__(self, format='csr', dtype='float32'):
        self.format, self.dtype, self.ndim = format, dtype, 2
