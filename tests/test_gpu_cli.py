"""End to end through the reference's command lines: bin/train.py -> model_<epoch>.bin -> bin/query.py -> TREC runs,
on synthetic data.npz / meta / topics files in the reference's formats; the runs are checked against the oracle."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(script, *args):
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bin', script)] + list(args), env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert out.returncode == 0, out.stdout.decode()[-4000:]
    return out.stdout.decode()


def parse_run(path):
    per_topic = {}
    for line in open(path):
        topic, q0, obj, rank, relevance, model = line.split()
        assert q0 == 'Q0'
        per_topic.setdefault(topic, []).append((obj, int(rank), float(relevance)))
    return per_topic


def load_model(path):
    with open(path, 'rb') as f:
        objs = []
        while True:
            try:
                objs.append(pickle.load(f))
            except EOFError:
                break
    return objs


@pytest.mark.parametrize('kind', ['loglinear', 'vectorspace'])
def test_train_and_query_cli(tmp_path, kind):
    from cvangysel import trec_utils
    from oracle import sert_oracle as O
    from sert_b200 import synth
    V, E, W = 300, 40, 4
    data, meta, topics = synth.write_corpus_files(str(tmp_path), kind, 77, V, E, W, n_train=64 * 6 + 5, n_val=64 * 2)
    model_out = str(tmp_path / 'model')
    args = ['--data', data, '--meta', meta, '--type', kind, '--iterations', '2', '--batch_size', '64',
            '--word_representation_size', '32', '--model_output', model_out]
    if kind == 'vectorspace':
        args += ['--one_hot_classes', '--entity_representation_size', '16', '--num_negative_samples', '5']
    log = run('train.py', *args)
    assert 'Epoch 2' in log and 'Saved model' in log
    for epoch in (0, 1, 2):
        assert os.path.exists('%s_%d.bin' % (model_out, epoch))
    objs = load_model(model_out + '_2.bin')
    train_args, predict_fn, R = objs[0], objs[1], objs[2]
    assert R.shape == (V, 32) and train_args.batch_size == 64
    run_out = str(tmp_path / 'run')
    qargs = ['--meta', meta, '--model', model_out + '_2.bin', '--topics', topics, '--run_out', run_out]
    if kind == 'vectorspace':
        qargs += ['--top', '10']
    qlog = run('query.py', *qargs)
    assert 'Skipping query' in qlog                       # the all-OOV topic
    ef = parse_run(run_out + '_ef')
    ep = parse_run(run_out + '_ep')
    assert os.path.exists(run_out + '_debug')
    assert 'T999' not in ef and len(ef) == 12
    # ---- check every ranked list against the oracle on the trained parameters ----
    with open(meta, 'rb') as f:
        data_args, words, tokens, entity_indices_inv, _ = (pickle.load(f) for _ in range(5))
    topic_terms = trec_utils.parse_topics([open(topics)])
    for topic_id, terms in topic_terms.items():
        toks = [words[t].id for t in trec_utils.parse_query(terms) if t in words]
        if not toks:
            continue
        if kind == 'loglinear':
            dist = O.loglinear_predict(predict_fn.representations, predict_fn.dense_w, predict_fn.dense_b,
                                       np.array(toks, dtype=np.int64).reshape(1, -1))[0]
            idx, val = O.loglinear_rank(dist)
            assert len(ef[topic_id]) == E                 # --top is ignored: all entities are ranked
        else:
            Eemb = objs[3]
            proj = O.vectorspace_predict(predict_fn.dense_w, predict_fn.dense_b, R[toks].mean(axis=0))
            idx, val = O.vectorspace_rank(O.normalise_rows(Eemb), proj, top=10)
            assert len(ef[topic_id]) == 10
        expected = O.write_run_order([(v, entity_indices_inv[int(i)]) for i, v in zip(idx, val)])
        got = ef[topic_id]
        assert [g[1] for g in got] == list(range(1, len(got) + 1))
        # identical ranked lists wherever the oracle's relevances are separated by more than fp32 noise
        ref_ids = [e[1] for e in expected]
        ref_vals = np.array([float(e[0]) for e in expected])
        got_ids = [g[0] for g in got]
        np.testing.assert_allclose([g[2] for g in got], ref_vals, rtol=2e-4, atol=1e-7)
        stable = np.abs(np.diff(ref_vals)) > 1e-5 * np.abs(ref_vals[:-1]).max()
        firm = np.concatenate([[True], stable]) & np.concatenate([stable, [True]])
        assert all(g == r for g, r, ok in zip(got_ids, ref_ids, firm) if ok)
    assert sum(len(v) for v in ep.values()) == sum(len(v) for v in ef.values())
