from collections import OrderedDict

import numpy as np
import theano
import theano.tensor as T

from lasagne import init, nonlinearities


class Layer(object):
    def __init__(self, incoming, name=None):
        if isinstance(incoming, tuple):
            self.input_shape, self.input_layer = incoming, None
        else:
            self.input_shape, self.input_layer = incoming.output_shape, incoming
        self.name = name
        self.params = OrderedDict()

    @property
    def output_shape(self):
        return self.get_output_shape_for(self.input_shape)

    def get_output_shape_for(self, input_shape):
        return input_shape

    def get_output_for(self, input, **kwargs):
        raise NotImplementedError

    def add_param(self, spec, shape, name=None, **tags):
        if name is not None and self.name is not None:
            name = '%s.%s' % (self.name, name)
        if isinstance(spec, theano.SharedVariable):
            param = spec
        else:
            value = spec(shape) if callable(spec) else np.asarray(spec)
            assert tuple(value.shape) == tuple(shape), (value.shape, shape)
            param = theano.shared(value, name=name)
        tags['trainable'] = tags.get('trainable', True)
        tags['regularizable'] = tags.get('regularizable', True)
        self.params[param] = set(tag for tag, on in tags.items() if on)
        return param

    def get_params(self, **tags):
        result = list(self.params.keys())
        only = set(tag for tag, on in tags.items() if on)
        if only:
            result = [p for p in result if not (only - self.params[p])]
        exclude = set(tag for tag, on in tags.items() if not on)
        if exclude:
            result = [p for p in result if not (self.params[p] & exclude)]
        return result


class InputLayer(Layer):
    def __init__(self, shape, input_var=None, name=None):
        self.shape, self.input_var, self.name = shape, input_var, name
        self.input_layer, self.params = None, OrderedDict()

    @property
    def output_shape(self):
        return self.shape


class ReshapeLayer(Layer):
    def __init__(self, incoming, shape, **kwargs):
        super(ReshapeLayer, self).__init__(incoming, **kwargs)
        self.shape = tuple(shape)

    def get_output_shape_for(self, input_shape):
        return self.shape

    def get_output_for(self, input, **kwargs):
        return input.reshape(self.shape)


class DenseLayer(Layer):
    def __init__(self, incoming, num_units, W=init.GlorotUniform(), b=init.Constant(0.),
                 nonlinearity=nonlinearities.linear, **kwargs):
        super(DenseLayer, self).__init__(incoming, **kwargs)
        self.nonlinearity = nonlinearity
        self.num_units = num_units
        num_inputs = int(np.prod(self.input_shape[1:]))
        self.W = self.add_param(W, (num_inputs, num_units), name='W')
        self.b = self.add_param(b, (num_units,), name='b', regularizable=False)

    def get_output_shape_for(self, input_shape):
        return (input_shape[0], self.num_units)

    def get_output_for(self, input, **kwargs):
        if input.ndim > 2:
            input = input.flatten(2)
        activation = T.dot(input, self.W)
        if self.b is not None:
            activation = activation + self.b.dimshuffle('x', 0)
        return self.nonlinearity(activation)


def get_all_layers(layer):
    chain = []
    while layer is not None:
        chain.append(layer)
        layer = layer.input_layer
    return chain[::-1]


def get_output(layer, inputs=None, **kwargs):
    value = None
    for current in get_all_layers(layer):
        if isinstance(current, InputLayer):
            value = inputs if inputs is not None else current.input_var
        else:
            value = current.get_output_for(value, **kwargs)
    return value


def get_all_params(layer, **tags):
    params = []
    for current in get_all_layers(layer):
        for p in current.get_params(**tags):
            if p not in params:
                params.append(p)
    return params
