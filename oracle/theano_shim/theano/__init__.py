"""Minimal eager restatement of the Theano 0.8.2 API surface used by sert/models.py (test infrastructure;
see ../README.md).  Variables are lazy expression nodes evaluated with torch CPU float32 tensors."""
import collections

import numpy as np
import scipy.sparse
import torch

__version__ = '0.8.2-shim'


class _Config(object):
    device = 'cpu'
    floatX = 'float32'


config = _Config()

_TORCH_DTYPES = {
    'float32': torch.float32, 'float64': torch.float64, 'int8': torch.int8, 'int16': torch.int16,
    'int32': torch.int32, 'int64': torch.int64, 'uint8': torch.uint8, 'uint16': torch.int32,
    'uint32': torch.int64, 'uint64': torch.int64,
}


def _as_torch(value, dtype=None):
    if isinstance(value, torch.Tensor):
        return value
    arr = np.asarray(value)
    if arr.dtype == np.uint16:
        arr = arr.astype(np.int32)
    elif arr.dtype in (np.uint32, np.uint64):
        arr = arr.astype(np.int64)
    elif arr.dtype == np.float64 and dtype is None:
        arr = arr.astype(np.float32)          # python floats become floatX constants
    return torch.from_numpy(np.require(arr, requirements='C').copy() if arr.ndim else np.array(arr))


class TensorType(object):
    def __init__(self, dtype='float32', ndim=0):
        self.dtype, self.ndim = dtype, ndim


class Variable(object):
    """Lazy expression node: op(*evaluated inputs) -> torch tensor (or scipy matrix for sparse nodes)."""

    def __init__(self, op, inputs=(), ndim=None, dtype='float32', name=None):
        self.op, self.inputs, self.ndim, self.dtype, self.name = op, tuple(inputs), ndim, dtype, name
        self.type = TensorType(dtype, ndim)

    def __repr__(self):
        return self.name or '<%s ndim=%s>' % (getattr(self.op, '__name__', 'op'), self.ndim)

    __hash__ = object.__hash__

    # ---- arithmetic ----
    def _bin(self, other, fn, reverse=False):
        other = as_variable(other)
        a, b = (other, self) if reverse else (self, other)
        ndim = max(a.ndim or 0, b.ndim or 0)
        dtype = 'float32' if 'float' in (a.dtype + b.dtype) else a.dtype
        return Variable(fn, (a, b), ndim=ndim, dtype=dtype)

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: a + b, True)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: a - b, True)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: a * b, True)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: a / b, True)
    def __pow__(self, o): return self._bin(o, lambda a, b: a ** b)
    def __rpow__(self, o): return self._bin(o, lambda a, b: a ** b, True)
    def __neg__(self): return Variable(lambda a: -a, (self,), ndim=self.ndim, dtype=self.dtype)

    def __imul__(self, o): return self.__mul__(o)
    def __iadd__(self, o): return self.__add__(o)

    # ---- reductions / shape ----
    def sum(self, axis=None, keepdims=False, **kw):
        nd = 0 if axis is None else (self.ndim if keepdims else self.ndim - 1)
        if axis is None:
            return Variable(lambda a: a.sum(), (self,), ndim=0, dtype=self.dtype)
        return Variable(lambda a: a.sum(dim=axis, keepdim=keepdims), (self,), ndim=nd, dtype=self.dtype)

    def mean(self, axis=None, **kw):
        if axis is None:
            return Variable(lambda a: a.mean(), (self,), ndim=0, dtype=self.dtype)
        return Variable(lambda a: a.mean(dim=axis), (self,), ndim=self.ndim - 1, dtype=self.dtype)

    def clip(self, lo, hi):
        from theano import tensor
        return tensor.clip(self, lo, hi)

    def reshape(self, shape, ndim=None):
        shape = tuple(shape)
        return Variable(lambda a: a.reshape(shape), (self,), ndim=len(shape), dtype=self.dtype)

    def flatten(self, outdim=1):
        return Variable(lambda a: a.reshape(a.shape[:outdim - 1] + (-1,)), (self,), ndim=outdim, dtype=self.dtype)

    def dimshuffle(self, *pattern):
        if len(pattern) == 1 and isinstance(pattern[0], (tuple, list)):
            pattern = tuple(pattern[0])

        def op(a):
            keep = [p for p in pattern if p != 'x']
            dropped = [d for d in range(a.dim()) if d not in keep]
            for d in dropped:
                assert a.shape[d] == 1, 'dimshuffle can only drop broadcastable dimensions'
            out = a.permute(*(keep + dropped)).reshape([a.shape[d] for d in keep])
            for pos, p in enumerate(pattern):
                if p == 'x':
                    out = out.unsqueeze(pos)
            return out
        return Variable(op, (self,), ndim=len(pattern), dtype=self.dtype)

    def __getitem__(self, key):
        if isinstance(key, Variable):
            return Variable(lambda a, idx: a[idx.long()], (self, key), ndim=self.ndim - 1 + key.ndim, dtype=self.dtype)
        if isinstance(key, slice):
            start, stop = as_variable(key.start), as_variable(key.stop)

            def op(a, s, e):
                return a[int(s):int(e)]
            return Variable(op, (self, start, stop), ndim=self.ndim, dtype=self.dtype)
        raise NotImplementedError('indexing with %r' % (key,))


def as_variable(value):
    if isinstance(value, Variable):
        return value
    t = _as_torch(value)
    dtype = str(t.dtype).replace('torch.', '')
    return Variable(lambda: t, (), ndim=t.dim(), dtype=dtype, name='const')


class SharedVariable(Variable):
    def __init__(self, value, name=None, borrow=False):
        self.is_sparse = scipy.sparse.issparse(value)
        if self.is_sparse:
            self.value = value.tocsr()
            ndim, dtype = 2, str(value.dtype)
        else:
            self.value = _as_torch(np.asarray(value), dtype='keep') if not isinstance(value, torch.Tensor) else value
            if self.value.dtype == torch.float64:
                pass
            ndim, dtype = self.value.dim(), str(np.asarray(value).dtype)
        Variable.__init__(self, self._get, (), ndim=ndim, dtype=dtype, name=name)
        if self.is_sparse:
            from theano import sparse as S
            self.type = S.type.SparseType(dtype=dtype)

    def _get(self):
        return self.value

    def get_value(self, borrow=False, return_internal_type=False):
        if self.is_sparse:
            return self.value
        return self.value.detach().numpy()

    def set_value(self, value, borrow=False):
        self.value = _as_torch(np.asarray(value), dtype='keep')

    def __getitem__(self, key):
        if self.is_sparse and isinstance(key, slice):
            start, stop = as_variable(key.start), as_variable(key.stop)
            from theano import sparse as S
            out = Variable(lambda a, s, e: a[int(s):int(e)], (self, start, stop), ndim=2, dtype=self.dtype)
            out.type = S.type.SparseType(dtype=self.dtype)
            return out
        return Variable.__getitem__(self, key)


def shared(value, name=None, borrow=False, **kwargs):
    return SharedVariable(value, name=name, borrow=borrow)


# ---- evaluation ------------------------------------------------------------------------------------
def _evaluate(var, memo, givens):
    key = id(var)
    if key in memo:
        return memo[key]
    ctx = getattr(var, '_grad_ctx', None)
    if ctx is not None:
        out = ctx.grads(memo, givens)[var._grad_index]
    elif var in givens:
        out = _evaluate(givens[var], memo, givens)
    else:
        args = [_evaluate(i, memo, givens) for i in var.inputs]
        with torch.set_grad_enabled(bool(memo.get('__autograd__'))):
            out = var.op(*args)
    memo[key] = out
    return out


class _GradContext(object):
    """One torch.autograd.grad call per theano.grad(cost, wrt) group, shared by all update expressions.
    Must run before anything else of the function is evaluated (Function.__call__ guarantees it)."""

    def __init__(self, loss, wrt):
        self.loss, self.wrt = loss, list(wrt)

    def grads(self, memo, givens):
        key = ('grad', id(self))
        if key not in memo:
            leaves = []
            for p in self.wrt:
                assert id(p) not in memo, 'parameter evaluated before its gradient context'
                p.value = p.value.detach().requires_grad_(True)
                leaves.append(p.value)
            memo['__autograd__'] = True
            loss = _evaluate(self.loss, memo, givens)
            memo['__autograd__'] = False
            gs = torch.autograd.grad(loss, leaves, allow_unused=True)
            memo[key] = [torch.zeros_like(l) if g is None else g.detach() for g, l in zip(gs, leaves)]
            for p in self.wrt:
                p.value = p.value.detach()
        return memo[key]


def grad(cost, wrt, **kwargs):
    single = not isinstance(wrt, (list, tuple))
    wrt_list = [wrt] if single else list(wrt)
    ctx = _GradContext(cost, wrt_list)
    outs = []
    for i, p in enumerate(wrt_list):
        v = Variable(None, (), ndim=p.ndim, dtype=p.dtype, name='grad(%s)' % p.name)
        v._grad_ctx, v._grad_index = ctx, i
        outs.append(v)
    return outs[0] if single else outs


def _grad_contexts(exprs):
    seen, found, stack = set(), [], list(exprs)
    while stack:
        v = stack.pop()
        if id(v) in seen:
            continue
        seen.add(id(v))
        ctx = getattr(v, '_grad_ctx', None)
        if ctx is not None and ctx not in found:
            found.append(ctx)
        stack.extend(v.inputs)
    return found


class _Maker(object):
    class _FGraph(object):
        def toposort(self):
            return []
    fgraph = _FGraph()


class Function(object):
    def __init__(self, inputs, outputs, updates=None, givens=None, mode=None, on_unused_input=None, **kw):
        self.inputs, self.outputs = list(inputs), outputs
        self.updates = list((updates or {}).items())
        self.givens = dict(givens or {})
        self.single = not isinstance(outputs, (list, tuple))
        self.out_list = [outputs] if self.single else list(outputs)
        self.contexts = _grad_contexts([e for _, e in self.updates] + self.out_list)
        self.maker = _Maker()

    def __call__(self, *args):
        assert len(args) == len(self.inputs)
        memo = {'__autograd__': False}
        for var, val in zip(self.inputs, args):
            memo[id(var)] = _as_torch(np.asarray(val))
        for ctx in self.contexts:                 # differentiated forward passes first
            ctx.grads(memo, self.givens)
        new_values = [(sv, _evaluate(expr, memo, self.givens)) for sv, expr in self.updates]
        outs = [_evaluate(o, memo, self.givens) for o in self.out_list]
        for sv, val in new_values:
            sv.value = val.detach() if isinstance(val, torch.Tensor) else val
        outs = [np.asarray(o.detach().numpy()) if isinstance(o, torch.Tensor) else o for o in outs]
        return outs[0] if self.single else outs


def function(inputs, outputs=None, updates=None, givens=None, mode=None, on_unused_input=None, **kwargs):
    return Function(inputs, outputs, updates=updates, givens=givens, mode=mode, on_unused_input=on_unused_input)


class _Printing(object):
    @staticmethod
    def pprint(v):
        return repr(v)

    @staticmethod
    def debugprint(v, file=None):
        return repr(v)

    class Print(object):
        def __init__(self, message=''):
            self.message = message

        def __call__(self, v):
            return v


printing = _Printing()

from theano import tensor, sparse, compile  # noqa: E402,F401
