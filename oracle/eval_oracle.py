"""Oracle for the evaluation step -- restatement of the reference's eval.py:10-32
(TEST INFRASTRUCTURE ONLY; parity unpinned, the reference records no evaluation numbers).

Same loops, Python 3 spelling (zip objects made lists), sklearn's roc_auc_score as in the
reference (eval.py:3,26).  The ROC plot (eval.py:34-46) is not drawn.
"""
from operator import itemgetter

from sklearn.metrics import roc_auc_score


def run_evaluation(examples, predictions_by_method, precision_at=20):
    """predictions_by_method: {method: {u: {b: score}}} (what util.load_json returns per file)."""
    out = {}
    for method, predictions in predictions_by_method.items():
        total_precision = 0
        all_ys, all_ps = [], []
        for u in predictions:                                            # eval.py:17
            ys, ps = zip(*[(examples[u][b], predictions[u][b]) for b in predictions[u]])
            all_ys += ys
            all_ps += ps
            n = min(precision_at, len(ys))                               # eval.py:22
            top_ys = list(zip(*sorted(zip(ys, ps), key=itemgetter(1), reverse=True)))[0][:n]
            total_precision += sum(top_ys) / float(n)                    # eval.py:24
        roc_auc = roc_auc_score(all_ys, all_ps)                          # eval.py:26
        out[method] = {'precision_at_%d' % precision_at: total_precision / len(examples),
                       'roc_auc': roc_auc}
    return out
