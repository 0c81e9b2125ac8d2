#!/usr/bin/env python
"""Benchmark of the SERT hot path on B200 (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

Workload at every N: BASELINE.json configs[1] -- VectorSpaceLanguageModel |V|=100k |E|=50k d=128
window=10, B=4096, k=10, lambda=0.01, float32 (the reference computes in float32; a float64 op is a
hard error there).  A "step" is one training batch (forward, backward, dense Adam+L2 update of all
tables).  N>1: one process per GPU, independent replicas on disjoint data shards (training has no
exchange step in the reference; DESIGN.md "Multi-GPU"), value = total pairs / max-over-ranks time.
The same JSON line carries the entity-scoring metric (row-sharded over the N ranks, one NCCL
all-gather of per-shard top-k) under "scoring".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG2 = dict(V=100000, E=50000, dw=128, de=128, W=10, B=4096, k=10, lam=0.01)
CFG3 = dict(Q=10000, E=50000, d=128, k=100)
CFG4 = dict(Q=10000, E=1000000, d=256, k=100)
METRIC = '(word,entity) pairs/sec train'
WORKLOAD = ('BASELINE.json configs[1]: VectorSpaceLanguageModel V=100k E=50k d=128 window=10 B=4096 k=10 lambda=0.01 '
            '(Adam + dense L2, float32)')
MIN_TIMED_S = 0.4           # the K timed steps are repeated as whole blocks until the timed region is at least this long
UNIT = 'pairs/s'


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def measured_tensor_peak():
    """Sustained dense bf16 TFLOP/s of this pool's B200s (a scoring sweep is a long tensor-core job)."""
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        if 'bf16_tflops_sustained' in d:
            return float(d['bf16_tflops_sustained']), 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)'
    return 1374.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(object):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run_nvml(self):
        """NVML polling (about 1 kHz possible; 2 ms period here) -- the nvidia-smi CLI needs ~100 ms per query,
        longer than a short timed region."""
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        bits = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40}
        get_reasons = getattr(pynvml, 'nvmlDeviceGetCurrentClocksEventReasons',
                              getattr(pynvml, 'nvmlDeviceGetCurrentClocksThrottleReasons', None))
        while not self.stop_flag.is_set():
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            mask = get_reasons(h) if get_reasons else 0
            row = ['', str(sm), str(mx), '', '']
            row += ['Active' if mask & bits[n] else 'Not Active'
                    for n in ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')]
            self.samples.append(row)
            self.stop_flag.wait(0.002)

    def _run(self):
        try:
            return self._run_nvml()
        except Exception:
            pass
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], stdout=subprocess.PIPE,
                                     stderr=subprocess.DEVNULL, timeout=5).stdout.decode().strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop_flag.set()
        self.thread.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                mx.append(float(s[2]))
            except Exception:
                continue
            for name, v in zip(names, s[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


def make_problem(rank, n_batches, cfg=CFG2):
    from sert_b200 import synth
    seed = 20160816 + 2 + 1000 * rank
    rng = np.random.default_rng(seed)
    B = cfg['B']
    train, val = synth.vectorspace_corpus(seed, cfg['V'], cfg['E'], cfg['W'], B * n_batches, B)
    if os.environ.get('SERT_BENCH_UNIFORM'):      # diagnostic only: uniform instead of Zipf word / label ids
        u = np.random.default_rng(seed + 1)
        train = (u.integers(0, cfg['V'], train[0].shape).astype(train[0].dtype),
                 u.integers(0, cfg['E'], train[1].shape).astype(np.int32), train[2])
    R = synth.glorot(rng, (cfg['V'], cfg['dw']))
    Eemb = synth.glorot(rng, (cfg['E'], cfg['de']))
    Wp = synth.glorot(rng, (cfg['dw'], cfg['de']))
    bp = np.zeros(cfg['de'], np.float32)
    neg = rng.integers(0, cfg['E'], size=(n_batches, B, cfg['k'])).astype(np.int32)
    return dict(train=train, val=val, R=R, Eemb=Eemb, Wp=Wp, bp=bp, neg=neg)


def algorithmic_bytes_per_step(cfg=CFG2):
    """dense update 24 B/param + word gather/scatter + entity rows (+ indices)."""
    P = cfg['V'] * cfg['dw'] + cfg['E'] * cfg['de'] + cfg['dw'] * cfg['de'] + cfg['de']
    dense = 24 * P
    words = 2 * cfg['B'] * cfg['W'] * cfg['dw'] * 4 + cfg['B'] * cfg['W'] * 4
    ents = 2 * cfg['B'] * (1 + cfg['k']) * cfg['de'] * 4 + cfg['B'] * (1 + cfg['k']) * 4
    return dense, dense + words + ents


# --------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path: Theano/Lasagne cannot be installed here (no
    network; Python 3.12), so this arm times the numpy float32 restatement under oracle/ (kind: port)."""
    if rank != 0:
        return
    from oracle import sert_oracle as O
    cfg = CFG2
    sample_steps, warm = args.steps, args.warmup          # exactly what the driver asked for; one step = one full batch
    p = make_problem(0, sample_steps + warm)
    orc = O.VectorSpaceOracle(cfg['B'], p['R'], p['Wp'], p['bp'], p['Eemb'], cfg['lam'], p['train'], p['val'])
    for j in range(warm):
        orc.train_batch(j, p['neg'][j])
    t0 = time.perf_counter()
    for j in range(sample_steps):
        orc.train_batch(warm + j, p['neg'][warm + j])
    dt = time.perf_counter() - t0
    value = sample_steps * cfg['B'] / dt
    cores = len(os.sched_getaffinity(0))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': sample_steps, 'warmup': warm, 'ms_per_step': 1e3 * dt / sample_steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'parallelism': 'cpu'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d training batches of 4096 pairs (numpy f32 oracle port of sert/models.py; '
                                   'BLAS uses %d threads, elementwise passes are single-threaded like '
                                   "Theano's CPU ops)" % (sample_steps, cores)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from sert_b200 import _native as N, models
    from sert_b200.scoring import ShardedScorer

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    cfg = CFG2
    steps, warm = args.steps, max(args.warmup, 3)
    n_batches = steps + warm
    p = make_problem(rank, n_batches)
    model = models.VectorSpaceLanguageModel(
        batch_size=cfg['B'], window_size=cfg['W'], num_negative_samples=cfg['k'],
        representations_init=p['R'], entity_representations_init=p['Eemb'], regularization_lambda=cfg['lam'],
        training_set=p['train'], validation_set=p['val'], dense_init=(p['Wp'], p['bp']),
        loss_slots=max(1024, n_batches))
    nat = model._native
    lib = nat.lib
    neg_dev = torch.from_numpy(p['neg']).cuda()
    order = np.arange(n_batches, dtype=np.int64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def train(lo, hi):
        N.check(lib.sert_train_batches(nat.handle, N.host_ptr(order[lo:hi]), hi - lo,
                                       N.c_void_p(neg_dev.data_ptr() + lo * cfg['B'] * cfg['k'] * 4), lo))

    # ---- device-resident throughput ("value") ----
    # One block = the K timed steps (batches warm..warm+K-1).  A block of K=20 steps lasts 2.4 ms, too short for the
    # clock sampler and for the driver's wall clock to vouch for, so the block is repeated R times back to back
    # (further passes over the same K batches: the same kernels on the same shapes) until the timed region reaches
    # MIN_TIMED_S; ms_per_step = region / (R * K).
    train(0, warm)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    train(warm, n_batches)
    ev1.record()
    barrier()
    first_losses = np.empty(n_batches, np.float32)
    N.check(lib.sert_losses_fetch(nat.handle, 0, n_batches, N.host_ptr(first_losses)))
    est = torch.tensor([max(ev0.elapsed_time(ev1), 1e-3)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(est, op=dist.ReduceOp.MAX)
    repeats = max(1, int(np.ceil(MIN_TIMED_S * 1e3 / float(est.item()))))
    launches0 = N.launch_count()
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record()
        for _ in range(repeats):
            train(warm, n_batches)
        ev1.record()
        barrier()
    launches = N.launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    losses = np.empty(n_batches, np.float32)
    N.check(lib.sert_losses_fetch(nat.handle, 0, n_batches, N.host_ptr(losses)))
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    timed_steps = repeats * steps
    value = world * timed_steps * cfg['B'] / (ms_max * 1e-3)

    # ---- roofline of the dominant kernel (dense Adam+L2 update), CUDA events around each launch ----
    N.check(lib.sert_model_profile(nat.handle, 1))
    prof_steps = min(steps, 50)
    N.check(lib.sert_train_batches(nat.handle, N.host_ptr(order[:prof_steps]), prof_steps,
                                   N.c_void_p(neg_dev.data_ptr()), 0))
    tot, cnt, bpl = N.ctypes.c_double(0), N.c_int64(0), N.ctypes.c_double(0)
    N.check(lib.sert_model_profile_read(nat.handle, N.ctypes.byref(tot), N.ctypes.byref(cnt), N.ctypes.byref(bpl)))
    N.check(lib.sert_model_profile(nat.handle, 0))
    upd_ms = tot.value / max(cnt.value, 1)
    peak, peak_src = measured_peaks()
    achieved = bpl.value / (upd_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'dense_update_traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get('dram_bytes_per_launch')
    dense_bytes, step_bytes = algorithmic_bytes_per_step()

    # ---- end-to-end: host batches through the C-ABI, H2D + step + D2H loss every step ----
    x, y, w = p['train']
    B = cfg['B']
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    xs = pin(x.astype(np.int32))
    ys = pin(y.astype(np.int32))
    ws = pin(w.astype(np.float32))
    ns = pin(p['neg'])
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()

    def e2e_step(b):
        N.check(lib.sert_train_batch_host(
            nat.handle, N.c_void_p(xs.data_ptr() + b * B * cfg['W'] * 4), N.c_void_p(ys.data_ptr() + b * B * 4),
            None, None, None, N.c_void_p(ws.data_ptr() + b * B * 4), N.c_void_p(ns.data_ptr() + b * B * cfg['k'] * 4),
            N.c_void_p(loss_host.data_ptr())))

    for b in range(warm):
        e2e_step(b)
    barrier()
    sync_repeats = max(1, repeats // 4)
    t0 = time.perf_counter()
    for _ in range(sync_repeats):
        for b in range(warm, n_batches):
            e2e_step(b)
    torch.cuda.synchronize()
    e2e_sync_s = time.perf_counter() - t0
    te = torch.tensor([e2e_sync_s], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_sync_value = world * sync_repeats * steps * B / float(te.item())

    # the same, pipelined: batch b is copied and enqueued before the loss of batch b-1 is waited for and checked
    ticket, prev, lossv = N.c_int64(0), None, N.ctypes.c_float(0)

    def e2e_async_step(b):
        N.check(lib.sert_train_batch_host_async(
            nat.handle, N.c_void_p(xs.data_ptr() + b * B * cfg['W'] * 4), N.c_void_p(ys.data_ptr() + b * B * 4),
            N.c_void_p(ws.data_ptr() + b * B * 4), N.c_void_p(ns.data_ptr() + b * B * cfg['k'] * 4),
            N.ctypes.byref(ticket)))
        return ticket.value

    def e2e_wait(t):
        N.check(lib.sert_train_host_wait(nat.handle, t, N.ctypes.byref(lossv)))

    for b in range(warm):
        t = e2e_async_step(b)
        if prev is not None:
            e2e_wait(prev)
        prev = t
    e2e_wait(prev)
    prev = None
    barrier()
    t0 = time.perf_counter()
    for _ in range(repeats):
        for b in range(warm, n_batches):
            t = e2e_async_step(b)
            if prev is not None:
                e2e_wait(prev)
            prev = t
    e2e_wait(prev)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * timed_steps * B / float(te.item())
    h2d = B * cfg['W'] * 4 + B * 4 + B * 4 + B * cfg['k'] * 4
    model._native.close()
    del model

    # ---- entity scoring (row-sharded; one all-gather of per-shard top-k) ----
    def leg(fn, *a, **kw):
        """A secondary leg must not take the headline line down with it: its failure is reported in its place."""
        try:
            return fn(*a, **kw)
        except Exception as exc:                     # noqa: BLE001
            import traceback
            traceback.print_exc(file=sys.stderr)
            return {'error': '%s: %s' % (type(exc).__name__, exc)}

    bf16_mode = leg(run_bf16_state_mode, p, cfg, steps, warm, repeats, first_losses, losses, rank, world, barrier)
    if isinstance(bf16_mode, dict) and 'error' not in bf16_mode:
        bf16_mode['float32_rerun_control'] = leg(run_bf16_state_mode, p, cfg, steps, warm, repeats, first_losses, losses,
                                                 rank, world, barrier, state_dtype='float32')
    table_shards = None
    if world > 1:
        table_shards = leg(run_table_shards, cfg, steps, warm, repeats, ms_max / timed_steps, first_losses, rank, world, barrier)
    scoring = leg(run_scoring, CFG4, 'BASELINE.json configs[3]', rank, world, barrier, cpu=False)
    scoring_small = leg(run_scoring, CFG3, 'BASELINE.json configs[2]', rank, world, barrier, cpu=(rank == 0))

    loglinear = leg(run_loglinear_cfg1, rank) if rank == 0 else None
    product_search = leg(run_product_search_shape, rank) if rank == 0 else None
    loglinear_stress = None if os.environ.get('SERT_BENCH_SKIP_CFG5') else leg(run_loglinear_cfg5, rank, world, barrier)

    if rank == 0:
        cpu = leg(cpu_baseline_sample, p, first_losses)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps, 'warmup': warm,
            'ms_per_step': ms_max / timed_steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD,
                       'global_batch': world * B,
                       'parallelism': 'single' if world == 1 else 'replicas x%d (no training collective)' % world,
                       'l2_policy': 'working set (params+Adam state+grads = 307 MB) larger than L2; no flush',
                       'final_loss': float(losses[-1])},
            'clocks': clocks.summary(),
            'gpu_launches': int(launches),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                    'api': 'sert_train_batch_host_async + sert_train_host_wait: every step copies its pinned host batch '
                           'to the device and its loss back to the host; the loss of step b-1 is waited for after '
                           'step b is enqueued',
                    'synchronous_value': e2e_sync_value,
                    'synchronous_api': 'sert_train_batch_host (one host/device sync per step)'},
            'roofline': {'bound': 'hbm', 'kernel': 'dense_update_kernel<Adam, tables> (csrc/opt_kernels.cu), CUDA events '
                                                   'around every launch inside running steps',
                         'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': bpl.value, 'kernel_ms': upd_ms,
                         'kernel_share_of_step': upd_ms / (ms_max / timed_steps),
                         'step_algorithmic_bytes': step_bytes,
                         'step_frac_of_hbm_peak': step_bytes / (ms_max / timed_steps * 1e-3) / 1e9 / peak},
            'cpu_baseline': cpu,
            'parity_rel_err': (cpu or {}).get('parity_rel_err'),
            'timed_steps': timed_steps, 'timed_region_s': ms_max * 1e-3,
            'scoring_ms': (scoring or {}).get('ms'),
            'scoring_frac': ((scoring or {}).get('roofline') or {}).get('frac'),
            'scoring_lists_identical': (scoring or {}).get('lists_identical'),
            'scoring_small_ms': (scoring_small or {}).get('ms'),
            'scoring_small_frac': ((scoring_small or {}).get('roofline') or {}).get('frac'),
            'scoring_small_lists_identical': (scoring_small or {}).get('lists_identical'),
            'loglinear_stress_ms': (loglinear_stress or {}).get('ms_per_step'),
            'table_shards_ms': min([v['ms_per_step'] for v in (table_shards or {}).values()
                                    if isinstance(v, dict) and 'ms_per_step' in v] or [None]),
            'table_shards': table_shards,
            'bf16_state_mode': bf16_mode,
            'product_search_shape': product_search,
            'loglinear': loglinear,
            'loglinear_stress': loglinear_stress,
            'scoring': scoring,
            'scoring_small': scoring_small,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_bf16_state_mode(p, cfg, steps, warm, repeats, f32_first_losses, f32_last_losses, rank, world, barrier,
                        state_dtype='bfloat16'):
    """BASELINE.json configs[1] says "bf16": the perf mode next to the float32 parity headline.  Same model, same
    batches, same step sequence, with Adam's m / v stored as bfloat16 (stochastic rounding) -- sert_config.dtype_mode 1:
    the dense update streams 16 instead of 24 bytes per parameter.  Parameters, gradients, forward and backward stay
    float32.  Reports the throughput, the roofline of its dense update and the measured deviation from the float32
    run over the identical step sequence."""
    import torch
    import torch.distributed as dist
    from sert_b200 import _native as N, models
    n_batches = steps + warm
    model = models.VectorSpaceLanguageModel(
        batch_size=cfg['B'], window_size=cfg['W'], num_negative_samples=cfg['k'],
        representations_init=p['R'], entity_representations_init=p['Eemb'], regularization_lambda=cfg['lam'],
        training_set=p['train'], validation_set=p['val'], dense_init=(p['Wp'], p['bp']),
        loss_slots=max(1024, n_batches), optimizer_state_dtype=state_dtype)
    nat, lib = model._native, model._native.lib
    neg_dev = torch.from_numpy(p['neg']).cuda()
    order = np.arange(n_batches, dtype=np.int64)

    def train(lo, hi):
        N.check(lib.sert_train_batches(nat.handle, N.host_ptr(order[lo:hi]), hi - lo,
                                       N.c_void_p(neg_dev.data_ptr() + lo * cfg['B'] * cfg['k'] * 4), lo))

    train(0, n_batches)
    first = np.empty(n_batches, np.float32)
    N.check(lib.sert_losses_fetch(nat.handle, 0, n_batches, N.host_ptr(first)))
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(repeats):
        train(warm, n_batches)
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / (repeats * steps)
    last = np.empty(n_batches, np.float32)
    N.check(lib.sert_losses_fetch(nat.handle, 0, n_batches, N.host_ptr(last)))
    N.check(lib.sert_model_profile(nat.handle, 1))
    prof_steps = min(steps, 50)
    N.check(lib.sert_train_batches(nat.handle, N.host_ptr(order[:prof_steps]), prof_steps,
                                   N.c_void_p(neg_dev.data_ptr()), 0))
    tot, cnt, bpl = N.ctypes.c_double(0), N.c_int64(0), N.ctypes.c_double(0)
    N.check(lib.sert_model_profile_read(nat.handle, N.ctypes.byref(tot), N.ctypes.byref(cnt), N.ctypes.byref(bpl)))
    upd_ms = tot.value / max(cnt.value, 1)
    peak, peak_src = measured_peaks()
    achieved = bpl.value / (upd_ms * 1e-3) / 1e9
    arena_mb = nat.arena.numel() / 1e6
    nat.close()
    del model
    torch.cuda.empty_cache()
    rel = lambda a, b: float(np.max(np.abs(np.asarray(a, np.float64) - b) / np.abs(b)))
    if state_dtype == 'float32':
        # control: a second float32 run of the same sequence.  Its distance from the first one (float atomics order the
        # gradient sums differently from run to run, and training amplifies that) is the yardstick for the bf16 figure.
        return {'deviation_rel': rel(first, f32_first_losses),
                'deviation_rel_last_pass': rel(last[warm:], f32_last_losses[warm:]), 'ms_per_step': ms}
    return {'workload': WORKLOAD.replace('float32', 'float32 parameters and gradients, bfloat16 Adam state (dtype_mode 1)'),
            'value': world * cfg['B'] / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'dtype': 'f32 + bf16 optimiser state',
            'roofline': {'bound': 'hbm', 'kernel': 'dense_update_kernel<Adam, tables, bf16 state>', 'achieved': achieved,
                         'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': bpl.value, 'kernel_ms': upd_ms, 'traffic': bf16_traffic()},
            'deviation_rel': rel(first, f32_first_losses),
            'deviation_rel_after_%d_steps' % (n_batches + repeats * steps): rel(last[warm:], f32_last_losses[warm:]),
            'deviation_what': 'max relative difference of the per-batch training losses against the float32 run of '
                              'the identical step sequence (first pass / last timed pass)',
            'arena_mb': arena_mb}


def run_table_shards(cfg, steps, warm, repeats, single_ms, single_first_losses, rank, world, barrier):
    """SURVEY.md 8(e), vector-space training as ONE model over the N GPUs (strong scaling; the headline line stays N
    independent replicas, weak).  Every rank is fed rank 0's batches and computes the step's gradient; the 24 B/param
    Adam stream over the two tables is split into N pieces, and the new parameters reach the other ranks either through
    the update kernels' own stores into every rank's next parameter buffer (CUDA IPC over NVLink: the update IS the
    exchange, one 512-byte ncclAllReduce per step is the barrier) or through grouped ncclBroadcast calls."""
    import torch
    import torch.distributed as dist
    from sert_b200 import _native as N, models
    from sert_b200.comm import Communicator
    n_batches = steps + warm
    p = make_problem(0, n_batches)                       # the same batches, negatives and initial values on every rank
    neg_dev = torch.from_numpy(p['neg']).cuda()
    order = np.arange(n_batches, dtype=np.int64)
    comm = Communicator.from_torch_distributed()
    out = {'workload': WORKLOAD + '; ONE model, global batch %d, table pieces over %d ranks' % (cfg['B'], world),
           'scaling': 'strong', 'single_gpu_ms_per_step': single_ms}
    for label, peer in (('instance_shards', 'instances'), ('peer_stores', True), ('nccl_broadcast', False)):
        model = models.VectorSpaceLanguageModel(
            batch_size=cfg['B'], window_size=cfg['W'], num_negative_samples=cfg['k'],
            representations_init=p['R'], entity_representations_init=p['Eemb'], regularization_lambda=cfg['lam'],
            training_set=p['train'], validation_set=p['val'], dense_init=(p['Wp'], p['bp']),
            loss_slots=max(1024, n_batches), table_shard=comm, table_shard_peer_stores=peer)
        nat, lib = model._native, model._native.lib
        mode, own_b, own_e, table_floats = model.table_shard_info()

        def train(lo, hi):
            N.check(lib.sert_train_batches(nat.handle, N.host_ptr(order[lo:hi]), hi - lo,
                                           N.c_void_p(neg_dev.data_ptr() + lo * cfg['B'] * cfg['k'] * 4), lo))

        train(0, n_batches)
        first = np.empty(n_batches, np.float32)
        N.check(lib.sert_losses_fetch(nat.handle, 0, n_batches, N.host_ptr(first)))
        info0 = comm.info()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(repeats):
            train(warm, n_batches)
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / (repeats * steps)
        info1 = comm.info()
        # every rank must hold the same model: compare a checksum of the parameters with rank 0's
        R, Eemb = model.get_representations()
        digest = torch.tensor([float(np.sum(R, dtype=np.float64)), float(np.sum(Eemb, dtype=np.float64))],
                              dtype=torch.float64, device='cuda')
        ref = digest.clone()
        dist.broadcast(ref, 0)
        same = torch.tensor([1.0 if torch.equal(ref, digest) else 0.0], device='cuda')
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        own_bytes = (own_e - own_b) * 4
        dense_tail = (cfg['dw'] * cfg['de'] + cfg['de']) * 4
        leg = {'ms_per_step': ms, 'value': cfg['B'] / (ms * 1e-3), 'unit': UNIT,
               'speedup_vs_single_gpu': single_ms / ms,
               'replicas_identical': bool(same.item() == 1.0),
               'nccl_collectives_per_step': (info1['collectives'] - info0['collectives']) / float(repeats * steps)}
        if rank == 0:
            rel = np.abs(first.astype(np.float64) - single_first_losses) / np.abs(single_first_losses)
            leg['loss_rel_err_vs_single_gpu'] = float(rel.max())
        if peer:
            leg['exchange'] = ('dense_update_kernel / hot_update_kernel store each new 16-byte chunk into the next parameter '
                               'buffer of all %d ranks (CUDA IPC mappings, NVLink); the partial sums of the loss cross in a '
                               'barrier kernel over the same mappings (csrc/peer_sync.cu), no NCCL call on the step' % world)
            if peer == 'instances':
                leg['exchange'] = ('each rank runs the tile kernel over its %d instances and adds the gradient rows into their '
                                   'owners\' arenas (red.global.add.v4.f32 over NVLink); ' % (cfg['B'] // world)) + leg['exchange']
            # sent by this rank per step: with look-ahead only the rows of its piece that the next batch reads (counted
            # here from the batches themselves, mean over the timed ones), everything on the last step of a call
            E_f = cfg['E'] * cfg['de']
            e_lo, e_hi = min(own_b, E_f) // cfg['de'], min(own_e, E_f) // cfg['de']
            r_lo, r_hi = max(own_b - E_f, 0) // cfg['dw'], max(own_e - E_f, 0) // cfg['dw']
            x_all, y_all = p['train'][0], p['train'][1]
            sent = []
            for b in range(warm, min(n_batches, warm + 8)):
                sl = slice(b * cfg['B'], (b + 1) * cfg['B'])
                words = np.unique(x_all[sl])
                ents = np.unique(np.concatenate([y_all[sl].ravel(), p['neg'][b].ravel()]))
                rows_r = int(np.count_nonzero((words >= r_lo) & (words < r_hi)))
                rows_e = int(np.count_nonzero((ents >= e_lo) & (ents < e_hi)))
                sent.append((rows_r * cfg['dw'] + rows_e * cfg['de']) * 4 * (world - 1))
            if peer != 'instances':      # instance shards send a row only to the ranks whose instances read it (fewer bytes)
                leg['nvlink_bytes_sent_per_step_this_rank'] = float(np.mean(sent))
            leg['nvlink_bytes_sent_full_push_this_rank'] = (own_bytes + (dense_tail if rank == world - 1 else 0)) * (world - 1)
            leg['nccl_bytes_per_step'] = 0
        else:
            leg['exchange'] = ('%d grouped ncclBroadcast (one table piece per owner) + ncclBroadcast of the projection from '
                               'its owner + ncclAllReduce of 64 doubles, behind the update kernels' % world)
            leg['nccl_bytes_per_step'] = table_floats * 4 + dense_tail + 512
        leg['adam_stream_bytes_per_rank'] = 24 * (own_e - own_b)
        out[label] = leg
        nat.close()
        del model
        torch.cuda.empty_cache()
        barrier()
    comm.close()
    return out


def run_product_search_shape(rank, steps=200):
    """The reference's canonical LSE recipe (product-search.sh:102-147): window 4, B = 4096, d_w = 300, d_e = 128, k = 10,
    on the synthetic vocabulary / entity counts of configs[1].  One GPU, device-resident batches."""
    import torch
    from sert_b200 import _native as N, models, synth
    cfg = dict(V=100000, E=50000, dw=300, de=128, W=4, B=4096, k=10, lam=0.01)
    nb = 30
    seed = 20160816 + 7
    rng = np.random.default_rng(seed)
    train, val = synth.vectorspace_corpus(seed, cfg['V'], cfg['E'], cfg['W'], cfg['B'] * nb, cfg['B'])
    model = models.VectorSpaceLanguageModel(
        batch_size=cfg['B'], window_size=cfg['W'], num_negative_samples=cfg['k'],
        representations_init=synth.glorot(rng, (cfg['V'], cfg['dw'])),
        entity_representations_init=synth.glorot(rng, (cfg['E'], cfg['de'])), regularization_lambda=cfg['lam'],
        training_set=train, validation_set=val, loss_slots=1024)
    nat, lib = model._native, model._native.lib
    neg_dev = torch.from_numpy(rng.integers(0, cfg['E'], size=(nb, cfg['B'], cfg['k'])).astype(np.int32)).cuda()
    order = np.arange(nb, dtype=np.int64)
    run = lambda: N.check(lib.sert_train_batches(nat.handle, N.host_ptr(order), nb, N.c_void_p(neg_dev.data_ptr()), 0))
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(1, steps // nb)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * nb)
    losses = np.empty(nb, np.float32)
    N.check(lib.sert_losses_fetch(nat.handle, 0, nb, N.host_ptr(losses)))
    P = cfg['V'] * cfg['dw'] + cfg['E'] * cfg['de'] + cfg['dw'] * cfg['de'] + cfg['de']
    step_bytes = 24 * P + 2 * cfg['B'] * cfg['W'] * cfg['dw'] * 4 + 2 * cfg['B'] * (1 + cfg['k']) * cfg['de'] * 4
    peak, _ = measured_peaks()
    nat.close()
    del model
    torch.cuda.empty_cache()
    return {'workload': 'product-search.sh recipe: VectorSpaceLanguageModel V=100k E=50k d_w=300 d_e=128 window=4 B=4096 '
                        'k=10 lambda=0.01 (Adam + dense L2, float32)',
            'value': cfg['B'] / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms,
            'step_algorithmic_bytes': step_bytes, 'step_frac_of_hbm_peak': step_bytes / (ms * 1e-3) / 1e9 / peak,
            'final_loss': float(losses[-1])}


def bf16_traffic():
    tpath = os.path.join(ROOT, 'profiles', 'dense_update_traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            return (json.load(f).get('bf16_state') or {}).get('dram_bytes_per_launch')
    return None


def scoring_shard(sc, world, rank):
    """Rank `rank`'s rows of the synthetic entity matrix (unit-normalised N(0,1) rows)."""
    from sert_b200.scoring import shard_bounds
    begin, end = shard_bounds(sc['E'], world, rank)
    rng = np.random.default_rng(20160816 + 3 + 7919 * rank)
    ent = rng.standard_normal((end - begin, sc['d']), dtype=np.float32)
    ent /= np.linalg.norm(ent, axis=1)[:, None]
    return ent


def run_scoring(sc, label, rank, world, barrier, cpu):
    """Scored entities/s for Q queries x E entities, top-k: rows sharded over the ranks (each rank builds only
    its shard), local sweep + ONE ncclAllGather of the per-shard lists + merge, all inside libsert_b200."""
    import torch
    import torch.distributed as dist
    from sert_b200.scoring import EntityScorer, ShardedScorer
    ent = scoring_shard(sc, world, rank)
    qs = np.random.default_rng(20160816 + 4).standard_normal((sc['Q'], sc['d']), dtype=np.float32)
    qs /= np.linalg.norm(qs, axis=1)[:, None]
    scorer = ShardedScorer(ent, sc['E'], is_shard=True, max_queries=sc['Q'], max_k=128)
    q_dev = torch.from_numpy(qs).cuda()
    for _ in range(3):
        idx_dev, score_dev = scorer.topk_dev(q_dev, sc['k'])
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    launches0 = N_launches()
    s0.record()
    for _ in range(reps):
        scorer.topk_dev(q_dev, sc['k'])
    s1.record()
    barrier()
    launches = (N_launches() - launches0) / reps
    sms = torch.tensor([s0.elapsed_time(s1) / reps], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(sms, op=dist.ReduceOp.MAX)
    ms = float(sms.item())
    # e2e: host queries in, host lists out (H2D of the queries, D2H of the lists inside the timed region)
    # (page-locked host buffers, as the training e2e: the copies run at PCIe speed)
    qs_pin = torch.from_numpy(qs).pin_memory()
    out_pin = (torch.empty((sc['Q'], sc['k']), dtype=torch.int32).pin_memory(),
               torch.empty((sc['Q'], sc['k']), dtype=torch.float32).pin_memory())
    out_np = (out_pin[0].numpy(), out_pin[1].numpy())
    scorer.topk(qs_pin.numpy(), sc['k'], out=out_np)
    barrier()
    e2e_calls = 5                     # back to back, like the device-timed loop above: the mean of a warm sequence
    t0 = time.perf_counter()
    for _ in range(e2e_calls):
        idx_host, score_host = scorer.topk(qs_pin.numpy(), sc['k'], out=out_np)
    e2e_s = (time.perf_counter() - t0) / e2e_calls
    te = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    seeded, fell_back = scorer.local.stats()
    out = {'metric': 'scored entities/sec', 'value': sc['Q'] * sc['E'] / (ms * 1e-3), 'unit': 'entities/s',
           'workload': '%s: Q=%d x E=%d d=%d top-%d, float32 vectors, rows sharded over %d GPU(s), one ncclAllGather '
                       'issued by the library' % (label, sc['Q'], sc['E'], sc['d'], sc['k'], world),
           'ms': ms, 'e2e_value': sc['Q'] * sc['E'] / e2e_s, 'e2e_ms': e2e_s * 1e3,
           'e2e_h2d_bytes': int(qs.nbytes), 'e2e_d2h_bytes': int(idx_host.nbytes + score_host.nbytes),
           'e2e_calls_timed': e2e_calls,
           'launches_per_call': launches, 'seeded_sweeps': seeded, 'fallback_sweeps': fell_back,
           'algorithmic_tflops': 2.0 * sc['Q'] * sc['E'] * sc['d'] / (ms * 1e-3) / 1e12,
           'arithmetic': 'coarse-then-exact: one bf16 tcgen05 GEMM launch (2*Q*E*d flops) behind thresholds seeded from '
                         'a strided row sample, rigorous rounding-error margin, fp32 re-scoring of the survivors; '
                         'returned lists are the exact fp32 top-k'}
    tpeak, tsrc = measured_tensor_peak()
    # per-GPU fraction: the aggregate rate of the N shards against N times one GPU's measured peak
    out['roofline'] = {'bound': 'tensor', 'kernel': 'gemm_tc_kernel + top-k epilogue (csrc/gemm_tc.cu), whole sweep '
                                                    'incl. sample, finalize, all-gather and merge',
                       'achieved': out['algorithmic_tflops'], 'peak': tpeak * world, 'unit': 'TFLOP/s',
                       'frac': out['algorithmic_tflops'] / (tpeak * world), 'peak_source': tsrc + ' x %d GPUs' % world,
                       'traffic': None}
    # ---- verification at the benchmarked shape (outside the timed regions) ----
    if world > 1:
        out['comm'] = scorer.comm.info()
    if rank == 0:
        nv = min(512, sc['Q'])
        got_idx, got_score = idx_host[:nv], score_host[:nv]
        full = ent if world == 1 else np.concatenate([ent] + [scoring_shard(sc, world, r) for r in range(1, world)])
        if world > 1:
            single = EntityScorer(full, max_queries=nv, max_k=128)
            ref_idx, ref_score = single.topk(qs[:nv], sc['k'])
            single.close()
            del single
            out['lists_identical'] = bool((ref_idx == got_idx).all() and (ref_score == got_score).all())
            out['lists_identical_what'] = ('merged (row id, score) lists of %d queries, bit for bit, against ONE single-GPU '
                                           'scorer over all %d rows (rank 0)' % (nv, sc['E']))
        ne = 16
        exact = qs[:ne].astype(np.float64) @ full.astype(np.float64).T
        order = np.argsort(-exact, axis=1, kind='stable')[:, :sc['k']]
        ex_score = np.take_along_axis(exact, order, axis=1)
        sep = np.abs(np.diff(ex_score, axis=1)).min(axis=1) > 1e-6
        out['exact_check'] = {'queries': ne, 'max_abs_score_err': float(np.abs(got_score[:ne] - ex_score).max()),
                              'lists_equal_float64_ranking': bool((got_idx[:ne][sep] == order[sep]).all()),
                              'queries_with_separated_scores': int(sep.sum())}
        del full
    barrier()
    if cpu and world == 1:
        out['cpu_baseline'] = scoring_cpu_baseline(ent, qs, sc['k'])
    scorer.close()
    del scorer
    torch.cuda.empty_cache()
    return out


def N_launches():
    from sert_b200 import _native
    return _native.launch_count()


def scoring_cpu_baseline(ent, qs, k):
    """(i) the reference's literal path (bin/query.py:304-359): per-query kd-tree k-NN + per-candidate scoring,
    on a bounded sample of queries; (ii) a strong CPU baseline: batched sgemm + argpartition."""
    import sklearn.neighbors
    n_lit = 100
    nn = sklearn.neighbors.NearestNeighbors(n_neighbors=k, algorithm='kd_tree', metric='euclidean')
    nn.fit(ent)                                    # index build is not timed (done once in the callback's __init__)
    t0 = time.perf_counter()
    for i in range(n_lit):
        _, indices = nn.kneighbors(qs[i:i + 1])
        cands = {}
        for candidate in indices[0, :]:            # the per-candidate Python loop of bin/query.py:348-359
            cands[candidate] = (np.sum(ent[candidate, :] * qs[i, :]) + 1.0) / 2.0
        sorted(cands.items(), reverse=True, key=lambda kv: kv[1])
    lit = (time.perf_counter() - t0) / n_lit
    n_b = 512
    t0 = time.perf_counter()
    s = qs[:n_b] @ ent.T
    part = np.argpartition(-s, k, axis=1)[:, :k]
    np.take_along_axis(s, part, axis=1).argsort(axis=1)
    bat = (time.perf_counter() - t0) / n_b
    E = ent.shape[0]
    return {'literal_reference_path': {'value': E / lit, 'unit': 'entities/s', 'sample': '%d queries, sklearn kd_tree '
                                       'k-NN (index build untimed) + per-candidate scoring loop' % n_lit},
            'batched_sgemm_argpartition': {'value': E / bat, 'unit': 'entities/s', 'sample': '%d queries' % n_b},
            'cores': len(os.sched_getaffinity(0)), 'kind': 'port'}


def run_loglinear_cfg1(rank):
    """BASELINE.json configs[0]: log-linear V=5k E=200 d=64 window=10 B=1024 (the reference's CPU-runnable case)."""
    import torch
    from oracle import sert_oracle as O
    from sert_b200 import _native as N, models, synth
    V, E, dw, W, B, nb = 5000, 200, 64, 10, 1024, 60
    rng = np.random.default_rng(20160816 + 1)
    train, val = synth.loglinear_corpus(20160817, V, E, W, B * nb, B * 2)
    R, Wd, bd = synth.glorot(rng, (V, dw)), synth.glorot(rng, (dw, E)), np.zeros(E, np.float32)
    model = models.LanguageModel(batch_size=B, window_size=W, representations_init=R, output_layer_size=E,
                                 regularization_lambda=0.01, training_set=train, validation_set=val,
                                 dense_init=(Wd, bd))
    nat = model._native
    order = np.arange(nb, dtype=np.int64)
    N.check(nat.lib.sert_train_batches(nat.handle, N.host_ptr(order[:10]), 10, None, 0))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    N.check(nat.lib.sert_train_batches(nat.handle, N.host_ptr(order[10:]), nb - 10, None, 10))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (nb - 10)
    orc = O.LogLinearOracle(B, R, Wd, bd, 0.01, train, val)
    orc.train_batch(0)
    t0 = time.perf_counter()
    for b in range(1, 4):
        orc.train_batch(b)
    cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    model._native.close()
    return {'workload': 'BASELINE.json configs[0]: LanguageModel (log-linear) V=5k E=200 d=64 window=10 B=1024, '
                        'Adadelta + dense L2, exact per-word clipped path',
            'value': B / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms,
            'cpu_port_value': B / (cpu_ms * 1e-3), 'cpu_port_ms_per_step': cpu_ms}


def run_loglinear_cfg5(rank, world, barrier, steps=4):
    """BASELINE.json configs[4]: log-linear V=500k E=200k d=300 window=10 B=1024 full-softmax stress.  N>1: the E
    axis (dense layer columns, logits, softmax) is sharded over the ranks, the word table is replicated; five small
    NCCL exchanges per step (sert_b200/sharding.py) -- strong scaling of ONE model, unlike the replica headline."""
    import torch
    import torch.distributed as dist
    from sert_b200 import _native as N, models, sharding, synth
    V, E, dw, W, B = 500000, 200000, 300, 10, 1024
    nb = steps + 2
    rng = np.random.default_rng(20160816 + 5)                  # same seed on every rank: identical initial values
    train, val = synth.loglinear_corpus(20160821, V, E, W, B * nb, B)
    R, Wd, bd = synth.glorot(rng, (V, dw)), synth.glorot(rng, (dw, E)), np.zeros(E, np.float32)
    exchange = sharding.CommExchange() if world > 1 else None      # the five exchanges are NCCL calls of the library
    model = models.LanguageModel(batch_size=B, window_size=W, representations_init=R, output_layer_size=E,
                                 regularization_lambda=0.01, training_set=train, validation_set=val,
                                 dense_init=(Wd, bd), loss_slots=64, entity_shard=exchange)
    nat = model._native
    order = np.arange(nb, dtype=np.int64)
    nat._check(nat.lib.sert_train_batches(nat.handle, N.host_ptr(order[:2]), 2, None, 0))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nat._check(nat.lib.sert_train_batches(nat.handle, N.host_ptr(order[2:]), steps, None, 2))
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    losses = np.empty(nb, np.float32)
    N.check(nat.lib.sert_losses_fetch(nat.handle, 0, nb, N.host_ptr(losses)))
    arena_gb = nat.arena.numel() / 1e9
    nat.close()
    del model
    torch.cuda.empty_cache()
    tpeak, tsrc = measured_tensor_peak()
    hbm_peak, hsrc = measured_peaks()
    alg_tflops = 6.0 * B * W * dw * E / (ms * 1e-3) / 1e12
    # HBM floor of the step as built: the float32 logit matrix is written once and read three times (joint, acc_r, dZ),
    # dZ's two-block bf16 operand is written once and read by both gradient GEMMs (dX reads it per n-tile: twice), plus
    # the Adadelta stream over the parameters (24 B each)
    z_bytes = 4.0 * B * W * E
    step_bytes = (4 * z_bytes + 4 * z_bytes + 24.0 * (V * dw + dw * E + E)) / world
    return {'workload': 'BASELINE.json configs[4]: LanguageModel (log-linear) V=500k E=200k d=300 window=10 B=1024, '
                        'Adadelta + dense L2, exact per-word clipped path, E sharded over %d GPU(s)' % world,
            'value': B / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'scaling': 'strong',
            'algorithmic_tflops': alg_tflops,
            'roofline': {'bound': 'tensor', 'kernel': 'the three word x entity GEMMs of the step (gemm_tc_kernel, store '
                         'epilogue, pair operands) inside the whole step', 'achieved': alg_tflops, 'peak': tpeak * world,
                         'unit': 'TFLOP/s', 'frac': alg_tflops / (tpeak * world),
                         'frac_counting_the_three_bf16_products': 3.0 * alg_tflops / (tpeak * world),
                         'peak_source': tsrc + ' x %d GPUs' % world,
                         'step_hbm_bytes_as_built_per_gpu': step_bytes,
                         'step_hbm_floor_ms': step_bytes / (hbm_peak * 1e9) * 1e3, 'hbm_peak_source': hsrc},
            'exchanges_per_step': 0 if world == 1 else 5, 'exchange': 'ncclAllReduce / ncclAllGather issued by libsert_b200', 'arena_gb_per_gpu': arena_gb,
            'losses': [float(v) for v in losses[:3]]}


def cpu_baseline_sample(p, gpu_losses):
    """Oracle port timed on the host cores (rank 0, N=1 semantics): a bounded sample of the same workload -- the
    first 7 batches of the very problem the GPU trained on (same initial values, batch order and negatives), so the
    oracle's losses double as the parity check of the benchmarked shape."""
    from oracle import sert_oracle as O
    cfg = CFG2
    n = max(1, min(6, len(gpu_losses) - 1))
    orc = O.VectorSpaceOracle(cfg['B'], p['R'], p['Wp'], p['bp'], p['Eemb'], cfg['lam'], p['train'], p['val'])
    ref = [orc.train_batch(0, p['neg'][0])]
    t0 = time.perf_counter()
    for j in range(1, n + 1):
        ref.append(orc.train_batch(j, p['neg'][j]))
    dt = time.perf_counter() - t0
    ref = np.asarray(ref, np.float64)
    got = np.asarray(gpu_losses[:n + 1], np.float64)
    rel = float(np.max(np.abs(got - ref) / np.abs(ref)))
    return {'value': n * cfg['B'] / dt, 'unit': UNIT, 'cores': len(os.sched_getaffinity(0)), 'kind': 'port',
            'sample': '%d training batches of 4096 pairs, numpy f32 oracle (oracle/sert_oracle.py)' % n,
            'parity_rel_err': rel, 'parity_what': 'max relative difference of the first %d per-batch training losses, '
            'CUDA path vs oracle, at the benchmarked shape (tolerance 1e-4)' % (n + 1),
            'parity_ok': bool(rel <= 1e-4)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', choices=['ours', 'reference'], default='ours')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return 0
    run_ours(args, rank, world, local_rank)
    return 0


if __name__ == '__main__':
    sys.exit(main())
