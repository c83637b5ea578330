"""Device-side evaluation (blp_eval_*) against the restatement of eval.py."""
import numpy as np
import pytest

from conftest import pkg
from test_gpu_parity import mods  # noqa: F401

pytestmark = pytest.mark.gpu


def _case(seed, n_users, k_lo, k_hi, kind):
    rng = np.random.default_rng(seed)
    examples, preds = {}, {}
    for u in range(n_users):
        m = int(rng.integers(k_lo, k_hi + 1))
        bs = rng.choice(10_000, size=m, replace=False)
        examples[str(u)] = {str(b): int(rng.random() < 0.3) for b in bs}
        if kind == 'int':                       # common-neighbour style: many ties, ints and int 0
            preds[str(u)] = {str(b): int(rng.integers(0, 6)) for b in bs}
        elif kind == 'mixed':                   # adamic style: int 0 mixed with floats
            preds[str(u)] = {str(b): (0 if rng.random() < 0.4 else float(rng.random() * 5)) for b in bs}
        else:
            preds[str(u)] = {str(b): float(rng.normal()) for b in bs}
    return examples, preds


@pytest.mark.parametrize('seed,n_users,k_lo,k_hi,kind', [
    (0, 300, 1, 40, 'int'), (1, 200, 5, 90, 'mixed'), (2, 50, 200, 700, 'float'), (3, 1, 3, 3, 'int'),
    # long lists take the top-k selection path of k_precision_at_k: heavy ties, int 0 mixed with floats
    (4, 30, 150, 400, 'int'), (5, 20, 97, 300, 'mixed'),
])
def test_metrics_match_eval_py(mods, seed, n_users, k_lo, k_hi, kind):
    from oracle import eval_oracle
    evaluate = pkg('evaluate')
    examples, preds = _case(seed, n_users, k_lo, k_hi, kind)
    labels = [y for u in preds for y in (examples[u][b] for b in preds[u])]
    if len(set(labels)) < 2:
        with pytest.raises(ValueError):
            evaluate.metrics(*evaluate.flatten(examples, preds), len(examples))
        return
    for k in (1, 20):
        want = eval_oracle.run_evaluation(examples, {'m': preds}, precision_at=k)['m']
        off, lab, sc = evaluate.flatten(examples, preds)
        p, auc = evaluate.metrics(off, lab, sc, len(examples), precision_at=k)
        assert p == pytest.approx(want['precision_at_%d' % k], rel=1e-12, abs=1e-15)
        assert auc == pytest.approx(want['roc_auc'], rel=1e-12)


def test_run_evaluation_on_score_files(mods, tmp_path):
    """Scores written by the drop-in similarity.main, evaluated like eval.py's __main__."""
    from oracle import eval_oracle
    graph, synth = mods
    sim, util, evaluate = pkg('similarity'), pkg('util'), pkg('evaluate')
    eu, eb = synth.make_graph(500, 80, 2500, seed=31, shift_u=1.0, shift_b=2.0)
    pu, pv = synth.make_pairs(500, 80, eu, eb, 3000, k=8, seed=32, invalid_frac=0.02)
    ids_eu, ids_eb = synth.shared_ids(500, eu, eb)
    ids_pu, ids_pv = synth.shared_ids(500, pu, pv)
    ex = synth.examples_dict(ids_pu, ids_pv)
    rng = np.random.default_rng(5)
    for u in ex:
        for b in ex[u]:
            ex[u][b] = int(rng.random() < 0.25)
    d = str(tmp_path) + '/'
    util.write_edge_list(d + 'graph.txt', ids_eu, ids_eb)
    util.write_json(ex, d + 'examples.json')
    M = ['common_neighbors', 'jaccard', 'adamic_adar']
    sim.main(d + 'examples.json', d + 'graph.txt', M, [d + 'u_cn.json', d + 'u_jaccard.json', d + 'u_adamic.json'],
             M, [d + 'b_cn.json', d + 'b_jaccard.json', d + 'b_adamic.json'])
    methods = ['u_adamic', 'u_cn', 'u_jaccard', 'b_adamic', 'b_cn', 'b_jaccard']   # eval.py:52-57
    got = evaluate.run_evaluation(util.load_json(d + 'examples.json'), methods, data_dir=d, quiet=True)
    want = eval_oracle.run_evaluation(util.load_json(d + 'examples.json'),
                                      {m: util.load_json(d + m + '.json') for m in methods})
    for m in methods:
        for key in want[m]:
            assert got[m][key] == pytest.approx(want[m][key], rel=1e-12), (m, key)
