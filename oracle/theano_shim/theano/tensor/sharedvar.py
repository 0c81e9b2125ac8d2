from theano import SharedVariable as TensorSharedVariable  # noqa: F401
