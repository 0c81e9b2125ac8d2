"""B200-native drop-in for the class surface of the reference's ``sert/models.py``.

Same class tree, constructor keywords and method contracts as the reference
(``ModelInterface -> ModelBase -> LanguageModelBase -> {LanguageModel,
VectorSpaceLanguageModelBase -> VectorSpaceLanguageModel}``, sert/models.py:295,414,686,
804,905,1024), so ``bin/train.py`` / ``bin/query.py`` run unchanged.  Everything the
reference delegates to a compiled Theano graph (train_fn / test_fn / validate_fn /
predict_fn, sert/models.py:554-608) is executed by hand-written sm_100a kernels reached
through the C-ABI in ``include/sert_b200.h``; PyTorch only allocates HBM and copies arrays.

Differences that are deliberate (and documented in DESIGN.md):
* ``train()`` enqueues the whole shuffled epoch with one C call and checks the per-batch losses
  for NaN/Inf afterwards (the reference synchronises after every batch, sert/models.py:370-379);
  the RuntimeError and its message are the same.
* negatives are sampled on the device with Philox (the reference uses an unseeded RandomStreams,
  sert/models.py:956-973); ``train_fn`` / ``test_fn`` / ``validate_fn`` accept explicit negatives so
  parity runs can feed both sides the same draw.
"""
import logging
import time

import numpy as np
import scipy.sparse as sparse

from sert_b200 import _native as N


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('sert_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.')
    return torch


def glorot_uniform(shape, rng=np.random):
    """lasagne.init.GlorotUniform().sample(shape) for 2-D shapes (bin/train.py:128,170)."""
    a = np.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-a, a, size=shape).astype(np.float32)


# The reference passes a function here (sert/models.py:92-120); on this path the L2 term is evaluated inside the dense
# optimiser kernel (csrc/opt_kernels.cu), so the attribute only names the regulariser.
l2_regularization = 'l2'


class _NativeModel(object):
    """Owns the HBM arena + the sert_model handle."""

    def __init__(self, kind, batch, window, vocab, entities, word_dim, entity_dim=0, num_negatives=0,
                 lam=0.0, loss_slots=1 << 16, seed=None, device=None, inference_only=False, dtype_mode=0):
        torch = _torch()
        self.lib = N.load()
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        if seed is None:
            seed = int(np.random.randint(low=0, high=(1 << 30)))      # sert/models.py:958-959
        self.cfg = N.SertConfig(kind=kind, batch=batch, window=window, num_negatives=num_negatives or 0,
                                vocab=vocab, entities=entities, word_dim=word_dim, entity_dim=entity_dim or 0,
                                lambda_=lam, loss_slots=loss_slots, seed=seed,
                                inference_only=int(bool(inference_only)), dtype_mode=int(dtype_mode), reserved1=0)
        nbytes = N.c_size_t(0)
        N.check(self.lib.sert_model_arena_bytes(N.ctypes.byref(self.cfg), N.ctypes.byref(nbytes)))
        with torch.cuda.device(self.device):
            self.arena = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
            self.stream = torch.cuda.current_stream(self.device)
            handle = N.c_void_p()
            N.check(self.lib.sert_model_create(N.ctypes.byref(self.cfg), N.dev_ptr(self.arena), nbytes.value,
                                               N.c_void_p(self.stream.cuda_stream), N.ctypes.byref(handle)))
        self.handle = handle
        self.loss_slots = loss_slots
        self._keep = []          # torch tensors borrowed by the library
        self.exchange = None     # sharding.Exchange of an entity-sharded log-linear model

    def _check(self, rc):
        """N.check, but an exception raised inside the exchange callback wins over the library's message."""
        if rc != 0 and self.exchange is not None:
            self.exchange.raise_pending()
        N.check(rc)

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.sert_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tensor(self, which, array, slot=N.STATE_PARAM):
        a = np.ascontiguousarray(array, dtype=np.float32)
        N.check(self.lib.sert_model_set_tensor(self.handle, which, slot, N.host_ptr(a), a.size))

    def get_tensor(self, which, shape, slot=N.STATE_PARAM):
        out = np.empty(shape, dtype=np.float32)
        N.check(self.lib.sert_model_get_tensor(self.handle, which, slot, N.host_ptr(out), out.size))
        return out

    def attach(self, split, x, y, w):
        torch = _torch()
        n = int(x.shape[0])
        x_dev = torch.from_numpy(np.ascontiguousarray(x, dtype=np.int32)).to(self.device)
        y_dev = indptr = indices = data = w_dev = None
        if sparse.issparse(y):
            y = y.tocsr()
            indptr = torch.from_numpy(np.ascontiguousarray(y.indptr, dtype=np.int64)).to(self.device)
            indices = torch.from_numpy(np.ascontiguousarray(y.indices, dtype=np.int32)).to(self.device)
            data = torch.from_numpy(np.ascontiguousarray(y.data, dtype=np.float32)).to(self.device)
        else:
            y_dev = torch.from_numpy(np.ascontiguousarray(y, dtype=np.int32)).to(self.device)
        if w is not None:
            w_dev = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32)).to(self.device)
        self._keep.append((x_dev, y_dev, indptr, indices, data, w_dev))
        N.check(self.lib.sert_model_attach_dataset(
            self.handle, split, n, N.dev_ptr(x_dev), N.dev_ptr(y_dev), N.dev_ptr(indptr), N.dev_ptr(indices),
            N.dev_ptr(data), N.dev_ptr(w_dev)))

    def _neg_dev(self, negatives, n):
        if negatives is None:
            return None, None
        torch = _torch()
        neg = np.ascontiguousarray(negatives, dtype=np.int32)
        assert neg.size == n * self.cfg.batch * self.cfg.num_negatives, 'negatives must be (n, B, k)'
        t = torch.from_numpy(neg).to(self.device)
        return t, N.dev_ptr(t)

    def run_batches(self, mode, order, negatives=None):
        """mode: 'train' | 'test' (eval on train rows) | 'validate'. Returns float32 losses (len(order),)."""
        order = np.ascontiguousarray(order, dtype=np.int64)
        out = np.empty(order.size, dtype=np.float32)
        done = 0
        while done < order.size:
            n = min(self.loss_slots, order.size - done)
            chunk = order[done:done + n]
            keep, negp = self._neg_dev(None if negatives is None else np.asarray(negatives)[done:done + n], n)
            if mode == 'train':
                self._check(self.lib.sert_train_batches(self.handle, N.host_ptr(chunk), n, negp, 0))
            else:
                split = N.SPLIT_TRAIN if mode == 'test' else N.SPLIT_VALIDATE
                self._check(self.lib.sert_eval_batches(self.handle, split, N.host_ptr(chunk), n, negp, 0))
            N.check(self.lib.sert_losses_fetch(self.handle, 0, n, N.host_ptr(out[done:done + n])))
            del keep
            done += n
        return out


HOT_WORD_MIN_PER_BATCH = 64     # expected occurrences per batch from which a word's gradient row counts as hot
MAX_HOT_WORDS = 32              # kMaxHotRows (csrc/kernels.cuh)


def hot_word_ids(x, vocabulary_size, batch_size, min_per_batch=HOT_WORD_MIN_PER_BATCH, max_words=MAX_HOT_WORDS):
    """The (at most `max_words`) word ids expected at least `min_per_batch` times in a batch of `batch_size` windows
    of x (N, W), most frequent first; ties keep the lower id."""
    x = np.asarray(x)
    if x.size == 0:
        return np.zeros(0, np.int32)
    per_batch = np.bincount(x.ravel(), minlength=vocabulary_size) * (float(batch_size) / x.shape[0])
    order = np.argsort(-per_batch, kind='stable')[:max_words]
    return order[per_batch[order] >= min_per_batch].astype(np.int32)


class ModelInterface(object):
    """sert/models.py:295-411."""

    TRAIN, VALIDATE, TEST = range(2, 5)

    def __init__(self, batch_size):
        assert batch_size > 0

        self.batch_size = batch_size

        logging.debug('Batch size: %d', self.batch_size)

    @classmethod
    def _get_batch_slice(cls, batch_index, batch_size):
        start = batch_index * batch_size
        end = (batch_index + 1) * batch_size

        return slice(start, end)

    def _number_of_batches(self, num_instances):
        return num_instances // self.batch_size

    def _iterate_batches(self, fn, num_instances, report_interval=10000, shuffle=False):
        """Per-batch host loop with the reference's contract (sert/models.py:351-399); the epoch-at-once
        device path used by train()/train_error()/validation_error() is _run_epoch()."""
        start = time.time()

        num_batches = self._number_of_batches(num_instances)
        incomplete_batch_size = num_instances % self.batch_size
        if incomplete_batch_size > 0:
            logging.warning('\tIgnoring incomplete batch of size %d.', incomplete_batch_size)

        results = []

        batch_indices = list(range(num_batches))
        if shuffle:
            logging.debug('Shuffling batches.')

            np.random.shuffle(batch_indices)

        for batch_idx in batch_indices:
            results.append(fn(batch_idx))

            if not np.all(np.isfinite(results[-1])):
                raise RuntimeError(
                    'Encountered NaN or infinity ({error}) '
                    'during batch iteration '
                    '(batch {batches_finished}/{num_batches}).'.format(
                        error=results[-1],
                        batches_finished=len(results),
                        num_batches=len(batch_indices)))

            if results and (len(results) % report_interval == 0 or len(results) == num_batches):
                self._report(len(results), num_batches, start)

        return num_batches, results

    @staticmethod
    def _report(finished, num_batches, start):
        elapsed = max(float(time.time() - start), 1e-9)
        batches_per_second = finished / elapsed
        remaining = (num_batches - finished) / batches_per_second
        logging.info('\tProcessed %d batches; %.2f batches per second; '
                     '%d minutes %d seconds remaining.',
                     finished, batches_per_second, remaining / 60, remaining % 60)

    def train(self):
        raise NotImplementedError()

    def train_error(self):
        raise NotImplementedError()

    def validation_error(self):
        raise NotImplementedError()

    def get_state(self):
        raise NotImplementedError()


class ModelBase(ModelInterface):
    """sert/models.py:414-683: owns the device-resident data set and the four callables."""

    def __init__(self, batch_size, training_set, validation_set, learning_method):
        super(ModelBase, self).__init__(batch_size)

        self.learning_method = learning_method

        self.training_num_instances = training_set[1].shape[0]
        self.validation_num_instances = validation_set[1].shape[0]

        # Determine number of instance features.
        self.num_instance_features = np.prod(training_set[0].shape[1:])

        assert self.num_instance_features == training_set[0].shape[1]

        if np.prod(validation_set[0].shape):
            assert self.num_instance_features == validation_set[0].shape[1]
        else:
            validation_set = (
                validation_set[0].reshape(0, self.num_instance_features),
                validation_set[1])

        logging.info('Data set contains %d training instances '
                     'and %d validation instances',
                     self.training_num_instances,
                     self.validation_num_instances)

        assert training_set[0].dtype == validation_set[0].dtype
        assert training_set[1].dtype == validation_set[1].dtype

        self.input_dtype = training_set[0].dtype
        self.output_dtype = training_set[1].dtype

        is_training_y_sparse = sparse.isspmatrix_csr(training_set[1])
        is_validation_y_sparse = sparse.isspmatrix_csr(validation_set[1])

        if is_training_y_sparse != is_validation_y_sparse:     # sert/models.py:566-571
            raise RuntimeError('Either training or validation truths are '
                               'sparse while the other is dense.')

        self.training_set = training_set
        self.validation_set = validation_set

    # -- native plumbing ------------------------------------------------------------------
    def _attach_datasets(self):
        x, y, w = self.training_set
        xv, yv = self.validation_set
        if self._native.cfg.kind == N.KIND_LOGLINEAR:
            if not sparse.issparse(y):
                y, yv = sparse.csr_matrix(np.asarray(y, dtype=np.float32)), \
                    sparse.csr_matrix(np.asarray(yv, dtype=np.float32))
        for name, idx in (('training', x), ('validation', xv)):
            if idx.size and (int(idx.max()) >= self.vocabulary_size or int(idx.min()) < 0):
                raise AssertionError('%s word indices out of range' % name)
        self._native.attach(N.SPLIT_TRAIN, x, y, w)
        self._native.attach(N.SPLIT_VALIDATE, xv, yv, None)

    def _create_functions(self):
        """train_fn / test_fn / validate_fn(batch_index) -> float32 loss (sert/models.py:581-608)."""
        native = self._native

        def make(mode):
            def fn(batch_index, negatives=None):
                neg = None if negatives is None else np.asarray(negatives)[np.newaxis]
                return native.run_batches(mode, [int(batch_index)], neg)[0]
            fn.__name__ = mode + '_fn'
            return fn

        self.train_fn = make('train')
        self.test_fn = make('test')
        self.validate_fn = make('validate')

    def _run_epoch(self, mode, num_instances, shuffle=False, report_interval=10000, negatives=None,
                   order=None):
        start = time.time()
        num_batches = self._number_of_batches(num_instances)
        incomplete_batch_size = num_instances % self.batch_size
        if incomplete_batch_size > 0:
            logging.warning('\tIgnoring incomplete batch of size %d.', incomplete_batch_size)
        if order is None:
            order = list(range(num_batches))
            if shuffle:
                logging.debug('Shuffling batches.')
                np.random.shuffle(order)
        results = self._native.run_batches(mode, order, negatives)
        if num_batches:
            self._report(len(results), num_batches, start)
        return num_batches, results

    def train(self, order=None, negatives=None):
        logging.info('Training on %d training instances (%d batches).',
                     self.training_num_instances,
                     self._number_of_batches(self.training_num_instances))

        num_batches, errors = self._run_epoch(
            'train', self.training_num_instances, report_interval=1000, shuffle=True,
            order=order, negatives=negatives)

        return num_batches, np.mean(errors)

    def train_error(self, negatives=None):
        logging.info('Measuring error on %d training instances (%d batches).',
                     self.training_num_instances,
                     self._number_of_batches(self.training_num_instances))

        num_batches, errors = self._run_epoch('test', self.training_num_instances, negatives=negatives)

        return np.mean(errors), np.std(errors)

    def validation_error(self, negatives=None):
        logging.info('Measuring error on %d validation instances '
                     '(%d batches).',
                     self.validation_num_instances,
                     self._number_of_batches(self.validation_num_instances))

        num_batches, errors = self._run_epoch('validate', self.validation_num_instances, negatives=negatives)

        return np.mean(errors), np.std(errors)

    # -- checkpoint / resume (the reference saves no optimiser state and cannot resume) --
    def _tensor_shapes(self):
        raise NotImplementedError()

    def get_checkpoint(self):
        """Every parameter tensor with both optimiser-state arrays plus the optimiser step, as host arrays."""
        ckpt = {'step': np.int64(self._native_step())}
        seed, draws = N.ctypes.c_uint64(0), N.ctypes.c_uint64(0)
        N.check(self._native.lib.sert_model_get_sampler(self._native.handle, N.ctypes.byref(seed), N.ctypes.byref(draws)))
        ckpt['sampler_seed'], ckpt['sampler_draws'] = np.uint64(seed.value), np.uint64(draws.value)
        for name, (which, shape) in self._tensor_shapes().items():
            for slot, suffix in ((N.STATE_PARAM, ''), (N.STATE_S1, '/state1'), (N.STATE_S2, '/state2')):
                ckpt[name + suffix] = self._gather_columns(name, self._native.get_tensor(which, shape, slot))
        return ckpt

    def set_checkpoint(self, ckpt):
        for name, (which, shape) in self._tensor_shapes().items():
            for slot, suffix in ((N.STATE_PARAM, ''), (N.STATE_S1, '/state1'), (N.STATE_S2, '/state2')):
                array = self._slice_columns(name, np.asarray(ckpt[name + suffix], dtype=np.float32))
                assert array.shape == tuple(shape), (name + suffix, array.shape, shape)
                self._native.set_tensor(which, array, slot)
        N.check(self._native.lib.sert_model_set_step(self._native.handle, int(ckpt['step'])))
        if 'sampler_seed' in ckpt:       # the device negative sampler continues where the checkpointed run stood
            N.check(self._native.lib.sert_model_set_sampler(self._native.handle, int(ckpt['sampler_seed']),
                                                            int(ckpt['sampler_draws'])))

    # entity-sharded models (LanguageModel(entity_shard=...)) hold column slices of some tensors; checkpoints and
    # get_state() always carry the full tensors
    _column_sharded = ()
    _shard_span = None

    def _gather_columns(self, name, array):
        if self._shard_span is None or name not in self._column_sharded:
            return array
        return self._native.exchange.gather_columns(array, self._shard_span[2])

    def _slice_columns(self, name, array):
        if self._shard_span is None or name not in self._column_sharded:
            return array
        return np.ascontiguousarray(array[..., self._shard_span[0]:self._shard_span[1]])

    def _native_step(self):
        t = N.c_int64(0)
        N.check(self._native.lib.sert_model_get_step(self._native.handle, N.ctypes.byref(t)))
        return t.value

    def get_state(self):
        state = [self.predict_fn]

        all_representations = self.get_representations()
        if not isinstance(all_representations, (tuple, list)):
            all_representations = (all_representations, )

        for representations in all_representations:
            state.append(representations)

        return state

    def get_representations(self):
        raise RuntimeError()


class LanguageModelBase(ModelBase):
    """sert/models.py:686-801."""

    def __init__(self, window_size, representations_init, regularization_lambda, regularization_fn, **kwargs):
        super(LanguageModelBase, self).__init__(**kwargs)

        assert window_size >= 1
        self.window_size = window_size

        self.initial_representations = representations_init

        self.vocabulary_size = representations_init.shape[0]
        self.representation_size = representations_init.shape[1]

        self.regularization_lambda = regularization_lambda

        self.regularization_fn = regularization_fn

    def get_representations(self):
        return self._native.get_tensor(N.PARAM_WORD_REPR, (self.vocabulary_size, self.representation_size))


class LogLinearPredictFn(object):
    """Picklable predict_fn of the log-linear model: (batch uintK (B,W), mask int8 (B,W)) -> f32 (B,W,E)
    per-word softmax (sert/models.py:880-890; the mask is accepted and unused there).  Carries host copies of
    R, W, b (the reference pickles the compiled Theano function, which embeds them) and builds its device
    model on first use in the process that unpickled it."""

    def __init__(self, representations, dense_w, dense_b, batch_size, window_size):
        self.representations = np.ascontiguousarray(representations, dtype=np.float32)
        self.dense_w = np.ascontiguousarray(dense_w, dtype=np.float32)
        self.dense_b = np.ascontiguousarray(dense_b, dtype=np.float32)
        self.batch_size, self.window_size = int(batch_size), int(window_size)
        self._native = None

    def __getstate__(self):
        d = dict(self.__dict__)
        d['_native'] = None
        return d

    def _ensure(self):
        if self._native is None:
            V, dw = self.representations.shape
            E = self.dense_w.shape[1]
            nat = _NativeModel(N.KIND_LOGLINEAR, self.batch_size, self.window_size, V, E, dw, loss_slots=4,
                               inference_only=True)
            nat.set_tensor(N.PARAM_WORD_REPR, self.representations)
            nat.set_tensor(N.PARAM_DENSE_W, self.dense_w)
            nat.set_tensor(N.PARAM_DENSE_B, self.dense_b)
            self._native = nat
        return self._native

    def __call__(self, batch, mask=None):
        nat = self._ensure()
        batch = np.ascontiguousarray(batch, dtype=np.int32)
        assert batch.ndim == 2 and batch.shape[1] == self.window_size
        rows, E = batch.shape[0], self.dense_w.shape[1]
        out = np.empty((rows, self.window_size, E), dtype=np.float32)
        done = 0
        while done < rows:
            n = min(self.batch_size, rows - done)
            N.check(nat.lib.sert_predict_loglinear(nat.handle, N.host_ptr(batch[done:done + n]), n,
                                                   N.host_ptr(out[done:done + n])))
            done += n
        return out


    def rank(self, batch, requests, top=None):
        """LogLinearCallback.process for every query of a WordBatcher batch, on the device (bin/query.py:204-233):
        ``requests`` = [(first_row, num_tokens), ...] in ``batch`` (rows, W).  Returns per query
        (entity ids (top,), relevances (top,), per-term normalised entropies (T,), normalised entropy of the final
        distribution, its mass); top=None ranks all E entities like the reference.  Only (queries, top) ids and
        relevances cross PCIe -- not the (rows, W, E) tensor ``__call__`` returns."""
        nat = self._ensure()
        batch = np.ascontiguousarray(batch, dtype=np.int32)
        assert batch.ndim == 2 and batch.shape[1] == self.window_size
        E = self.dense_w.shape[1]
        top = E if top is None else min(int(top), E)
        results = []
        for lo in range(0, len(requests), self.batch_size):
            chunk = requests[lo:lo + self.batch_size]
            row_lo = chunk[0][0]
            row_hi = max(first + -(-n // self.window_size) for first, n in chunk)
            assert row_hi - row_lo <= self.batch_size
            first = np.array([f - row_lo for f, _ in chunk], np.int32)
            terms = np.array([n for _, n in chunk], np.int32)
            nq = len(chunk)
            idx = np.empty((nq, top), np.int32)
            rel = np.empty((nq, top), np.float32)
            term_entropy = np.empty(int(terms.sum()), np.float32)
            entropy = np.empty(nq, np.float32)
            mass = np.empty(nq, np.float32)
            N.check(nat.lib.sert_ll_rank_queries(
                nat.handle, N.host_ptr(batch[row_lo:row_hi]), row_hi - row_lo, N.host_ptr(first), N.host_ptr(terms), nq,
                top, N.host_ptr(idx), N.host_ptr(rel), N.host_ptr(term_entropy), N.host_ptr(entropy), N.host_ptr(mass)))
            ends = np.cumsum(terms)
            for j in range(nq):
                results.append((idx[j], rel[j], term_entropy[ends[j] - terms[j]:ends[j]], float(entropy[j]),
                                float(mass[j])))
        return results


class VectorSpacePredictFn(object):
    """Picklable predict_fn of the vector-space model: avg word embedding (dw,) -> tanh(avg.W+b) (de,),
    no clip (sert/models.py:1107-1118).  ``project`` is the batched form used by the GPU ranker."""

    def __init__(self, dense_w, dense_b):
        self.dense_w = np.ascontiguousarray(dense_w, dtype=np.float32)
        self.dense_b = np.ascontiguousarray(dense_b, dtype=np.float32)
        self._native = None

    def __getstate__(self):
        d = dict(self.__dict__)
        d['_native'] = None
        return d

    def _ensure(self):
        if self._native is None:
            dw, de = self.dense_w.shape
            nat = _NativeModel(N.KIND_VECTORSPACE, 1024, 1, 4, 4, dw, entity_dim=de, num_negatives=1, loss_slots=4,
                               inference_only=True)
            nat.set_tensor(N.PARAM_DENSE_W, self.dense_w)
            nat.set_tensor(N.PARAM_DENSE_B, self.dense_b)
            self._native = nat
        return self._native

    def project(self, avg_matrix):
        nat = self._ensure()
        avg = np.ascontiguousarray(avg_matrix, dtype=np.float32)
        assert avg.ndim == 2 and avg.shape[1] == self.dense_w.shape[0]
        out = np.empty((avg.shape[0], self.dense_w.shape[1]), dtype=np.float32)
        N.check(nat.lib.sert_project_queries(nat.handle, N.host_ptr(avg), avg.shape[0], N.host_ptr(out)))
        return out

    def __call__(self, avg_word_embedding):
        # shape (1, de): the reference's DenseLayer adds b.dimshuffle('x', 0) to the (de,) product
        return self.project(np.asarray(avg_word_embedding, dtype=np.float32).reshape(1, -1))


class LanguageModel(LanguageModelBase):
    """Log-linear expert-finding model (sert/models.py:804-890): params [R, W, b], Adadelta, dense L2."""

    def __init__(self,
                 batch_size, window_size,
                 representations_init,
                 output_layer_size,
                 regularization_lambda,
                 training_set,
                 validation_set,
                 dense_init=None, device=None, loss_slots=1 << 16, entity_shard=None):
        """``entity_shard``: a ``sert_b200.sharding.Exchange`` (one process per GPU, e.g. ``DistExchange()``);
        this rank then owns a contiguous block of the E output columns, the word table is
        replicated, and train()/train_error()/validation_error()/get_state() return the same values on every
        rank.  All ranks must pass identical initial values and the same batch order."""
        super(LanguageModel, self).__init__(
            batch_size=batch_size,
            window_size=window_size,
            representations_init=representations_init,
            regularization_lambda=regularization_lambda,
            regularization_fn=l2_regularization,
            training_set=training_set, validation_set=validation_set,
            learning_method='adadelta')

        self.output_layer_size = int(output_layer_size)
        logging.debug('Input layer has shape %s.', (self.batch_size, self.window_size))

        if dense_init is None:          # lasagne DenseLayer defaults: GlorotUniform W, zero b
            dense_init = (glorot_uniform((self.representation_size, self.output_layer_size)),
                          np.zeros(self.output_layer_size, dtype=np.float32))

        self._local_entities = self.output_layer_size
        if entity_shard is not None:
            from sert_b200 import sharding
            begin, end = sharding.shard_bounds(self.output_layer_size, entity_shard.world, entity_shard.rank)
            assert end - begin >= 2, 'every shard needs at least two entity columns'
            self._shard_span = (begin, end, self.output_layer_size)
            self._column_sharded = ('dense_w', 'dense_b')
            self._local_entities = end - begin

        self._native = _NativeModel(
            N.KIND_LOGLINEAR, self.batch_size, self.window_size, self.vocabulary_size,
            self._local_entities, self.representation_size, lam=float(regularization_lambda),
            loss_slots=loss_slots, device=device)
        if entity_shard is not None:
            sharding.attach(self._native, entity_shard, self._shard_span[0], self.output_layer_size)
        self._native.set_tensor(N.PARAM_WORD_REPR, representations_init)
        self._native.set_tensor(N.PARAM_DENSE_W, self._slice_columns('dense_w', np.asarray(dense_init[0])))
        self._native.set_tensor(N.PARAM_DENSE_B, self._slice_columns('dense_b', np.asarray(dense_init[1])))
        self._attach_datasets()
        self._create_functions()

    def get_dense(self):
        """Full (dw,E) W and (E,) b; a collective over the shards when the model is entity-sharded."""
        shapes = self._tensor_shapes()
        return tuple(self._gather_columns(name, self._native.get_tensor(*shapes[name]))
                     for name in ('dense_w', 'dense_b'))

    def _tensor_shapes(self):
        return {'representations': (N.PARAM_WORD_REPR, (self.vocabulary_size, self.representation_size)),
                'dense_w': (N.PARAM_DENSE_W, (self.representation_size, self._local_entities)),
                'dense_b': (N.PARAM_DENSE_B, (self._local_entities,))}

    @property
    def predict_fn(self):
        w, b = self.get_dense()
        return LogLinearPredictFn(self.get_representations(), w, b, self.batch_size, self.window_size)


def inproduct_sigmoid_distance(target_embeddings, output):
    """sert/models.py:893-902, numpy form (host-side helper; the training path uses csrc/vs_kernels.cu)."""
    assert target_embeddings.ndim == output.ndim

    activation = 1.0 / (1.0 + np.exp(-np.sum(target_embeddings * output, axis=target_embeddings.ndim - 1)))

    return np.clip(activation, 1e-7, 1.0 - 1e-7)


class VectorSpaceLanguageModelBase(LanguageModelBase):
    """sert/models.py:905-1021."""

    def __init__(self,
                 batch_size, window_size,
                 num_negative_samples,
                 representations_init,
                 entity_representations_init,
                 regularization_lambda,
                 training_set,
                 validation_set):
        super(VectorSpaceLanguageModelBase, self).__init__(
            batch_size=batch_size,
            window_size=window_size,
            representations_init=representations_init,
            regularization_lambda=regularization_lambda,
            regularization_fn=l2_regularization,
            training_set=training_set, validation_set=validation_set,
            learning_method='adam')

        self.num_entities = entity_representations_init.shape[0]
        self.entity_representation_size = entity_representations_init.shape[1]

        assert self.training_set[1].ndim == 1, \
            'Only one-hot vectors supported.'

        assert num_negative_samples is None or num_negative_samples >= 0, \
            'Number of negative samples should be None, zero or positive ' \
            '(currently: {0}).'.format(num_negative_samples)

        self.num_negative_samples = num_negative_samples

    def get_representations(self):
        return (self._native.get_tensor(N.PARAM_WORD_REPR, (self.vocabulary_size, self.representation_size)),
                self._native.get_tensor(N.PARAM_ENTITY_REPR, (self.num_entities, self.entity_representation_size)))


class VectorSpaceLanguageModel(VectorSpaceLanguageModelBase):
    """Latent vector space (LSE) model (sert/models.py:1024-1118): params [Eemb, R, W, b], Adam, dense L2."""

    def __init__(self,
                 batch_size, window_size,
                 num_negative_samples,
                 representations_init,
                 entity_representations_init,
                 regularization_lambda,
                 training_set,
                 validation_set,
                 dense_init=None, device=None, seed=None, loss_slots=1 << 16, optimizer_state_dtype='float32',
                 table_shard=None, table_shard_peer_stores=True):
        """``table_shard``: a ``sert_b200.comm.Communicator`` (one process per GPU of one NVLink domain): ONE model at
        the global batch whose Adam stream over the two tables is split over the ranks (include/sert_b200.h:
        sert_model_set_table_shard_comm).  Every rank must be fed the same batches; parameters, optimiser state and
        the negative sampler are taken from rank 0.  ``table_shard_peer_stores``: True = the update kernels write the
        new parameters into every rank's copy over NVLink (CUDA IPC); False = grouped NCCL broadcasts;
        ``'instances'`` (or 2) = peer stores AND each rank runs the forward / backward of its own slice of the batch's
        instances only, adding the gradient rows into their owners' arenas over NVLink (no NCCL call on the step).

        ``optimizer_state_dtype``: 'float32' (the reference's arithmetic; parity mode) or 'bfloat16' (perf mode of
        BASELINE.json configs[1]: Adam's m and v stored as bfloat16 with stochastic rounding, 16 instead of 24 bytes
        per parameter and step; parameters, gradients and every forward/backward value stay float32)."""
        assert optimizer_state_dtype in ('float32', 'bfloat16')
        super(VectorSpaceLanguageModel, self).__init__(
            batch_size=batch_size,
            window_size=window_size,
            num_negative_samples=num_negative_samples,
            representations_init=representations_init,
            entity_representations_init=entity_representations_init,
            regularization_lambda=regularization_lambda,
            training_set=training_set,
            validation_set=validation_set)

        # sert/models.py:948 (_negative_sampling asserts num_negative_samples > 0 when the loss is built)
        assert num_negative_samples is not None and num_negative_samples > 0

        if dense_init is None:
            dense_init = (glorot_uniform((self.representation_size, self.entity_representation_size)),
                          np.zeros(self.entity_representation_size, dtype=np.float32))

        y = self.training_set[1]
        if y.size and (int(y.max()) >= self.num_entities or int(y.min()) < 0):
            raise AssertionError('entity labels out of range')

        self._native = _NativeModel(
            N.KIND_VECTORSPACE, self.batch_size, self.window_size, self.vocabulary_size, self.num_entities,
            self.representation_size, entity_dim=self.entity_representation_size,
            num_negatives=int(num_negative_samples), lam=float(regularization_lambda),
            loss_slots=loss_slots, seed=seed, device=device,
            dtype_mode=1 if optimizer_state_dtype == 'bfloat16' else 0)
        self._native.set_tensor(N.PARAM_WORD_REPR, representations_init)
        self._native.set_tensor(N.PARAM_ENTITY_REPR, entity_representations_init)
        self._native.set_tensor(N.PARAM_DENSE_W, dense_init[0])
        self._native.set_tensor(N.PARAM_DENSE_B, dense_init[1])
        self._attach_datasets()
        self.set_hot_words(self.pick_hot_words())
        self.table_shard = table_shard
        if table_shard is not None:
            N.check(self._native.lib.sert_model_set_table_shard_comm(
                self._native.handle, table_shard.handle,
                2 if table_shard_peer_stores in ('instances', 2) else int(bool(table_shard_peer_stores))))
        self._create_functions()

    def table_shard_info(self):
        """(mode, own_begin, own_end, table_floats): mode 0 none / 1 broadcast / 2 peer stores / 3 instance shards; this rank updates the
        floats [own_begin, own_end) of the two tables' table_floats."""
        mode, b, e, n = N.ctypes.c_int32(0), N.c_int64(0), N.c_int64(0), N.c_int64(0)
        N.check(self._native.lib.sert_model_table_shard_info(self._native.handle, N.ctypes.byref(mode), N.ctypes.byref(b),
                                                             N.ctypes.byref(e), N.ctypes.byref(n)))
        return mode.value, b.value, e.value, n.value

    def get_checkpoint(self):
        if self.table_shard is not None:      # each rank's Adam state is current only inside its own piece
            N.check(self._native.lib.sert_model_gather_table_state(self._native.handle))
        return super(VectorSpaceLanguageModel, self).get_checkpoint()

    def pick_hot_words(self):
        """Word ids whose gradient row receives so many additions per batch that they serialise in L2
        (include/sert_b200.h: sert_model_set_hot_words), from the training set's word counts."""
        return hot_word_ids(self.training_set[0], self.vocabulary_size, self.batch_size)

    def set_hot_words(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        N.check(self._native.lib.sert_model_set_hot_words(self._native.handle, N.host_ptr(ids), int(ids.size)))
        self.hot_words = ids

    def train_stream(self, batches, depth=2):
        """Trains on host batches that are NOT part of the device-resident data set, pipelined: the host->device copy and
        the enqueueing of batch n+1 overlap the device work of batch n (include/sert_b200.h:
        sert_train_batch_host_async).  `batches` yields (x (B,W), y (B,), w (B,) or None, negatives (B,k) or None);
        returns the float32 train losses, one per batch; NaN/Inf raises like train() (sert/models.py:372-379).
        Arrays are used in place when they already are C-contiguous int32 / float32 (pin them for real overlap)."""
        nat = self._native
        B, W, k = self.batch_size, self.window_size, nat.cfg.num_negatives
        losses, inflight = [], []
        ticket, value = N.c_int64(0), N.ctypes.c_float(0)

        def wait_oldest():
            t, keep = inflight.pop(0)
            N.check(nat.lib.sert_train_host_wait(nat.handle, t, N.ctypes.byref(value)))
            losses.append(value.value)
            del keep

        for x, y, w, neg in batches:
            x = np.ascontiguousarray(x, dtype=np.int32)
            y = np.ascontiguousarray(y, dtype=np.int32)
            w = None if w is None else np.ascontiguousarray(w, dtype=np.float32)
            neg = None if neg is None else np.ascontiguousarray(neg, dtype=np.int32)
            assert x.shape == (B, W) and y.shape == (B,), 'host batches are (batch_size, window_size) / (batch_size,)'
            assert w is None or w.shape == (B,)
            assert neg is None or neg.shape == (B, k)
            N.check(nat.lib.sert_train_batch_host_async(nat.handle, N.host_ptr(x), N.host_ptr(y), N.host_ptr(w),
                                                        N.host_ptr(neg), N.ctypes.byref(ticket)))
            inflight.append((ticket.value, (x, y, w, neg)))
            while len(inflight) > max(1, depth) - 1 and len(inflight) > 1:
                wait_oldest()
        while inflight:
            wait_oldest()
        return np.asarray(losses, dtype=np.float32)

    def get_dense(self):
        return (self._native.get_tensor(N.PARAM_DENSE_W, (self.representation_size, self.entity_representation_size)),
                self._native.get_tensor(N.PARAM_DENSE_B, (self.entity_representation_size,)))

    def _tensor_shapes(self):
        return {'entity_representations': (N.PARAM_ENTITY_REPR, (self.num_entities, self.entity_representation_size)),
                'representations': (N.PARAM_WORD_REPR, (self.vocabulary_size, self.representation_size)),
                'dense_w': (N.PARAM_DENSE_W, (self.representation_size, self.entity_representation_size)),
                'dense_b': (N.PARAM_DENSE_B, (self.entity_representation_size,))}

    @property
    def predict_fn(self):
        w, b = self.get_dense()
        return VectorSpacePredictFn(w, b)
