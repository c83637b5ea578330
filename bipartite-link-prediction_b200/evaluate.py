"""Drop-in for the reference's ``eval.py`` metrics.

    run_evaluation(examples, methods, precision_at=20)                             eval.py:10

For every method the score file ``./data/test/<method>.json`` is paired with ``examples`` exactly
as eval.py:17-18 does (``for u in predictions: for b in predictions[u]``), precision@k is
averaged over ``len(examples)`` (:22-24,31) and one global ROC-AUC is taken over all pairs (:26).
The numbers come from the CUDA library (``blp_eval_precision_at_k`` / ``blp_eval_roc_auc``).  The
ROC plot of eval.py:34-46 is not drawn (no matplotlib here); the curve is a by-product of the
same sorted order and is left to the caller.
"""
import ctypes

import numpy as np
import torch

from . import _lib, util


def flatten(examples, predictions):
    """Arrays in eval.py's iteration order: offsets per user, labels, scores."""
    offsets, labels, scores = [0], [], []
    for u in predictions:
        row, ex = predictions[u], examples[u]
        for b in row:
            labels.append(ex[b])
            scores.append(row[b])
        offsets.append(len(labels))
    return (np.asarray(offsets, dtype=np.int64), np.asarray(labels, dtype=np.int32),
            np.asarray(scores, dtype=np.float64))


def metrics(offsets, labels, scores, n_example_users, precision_at=20, device=None):
    """(precision@k, roc_auc) of one method from flat arrays (numpy or CUDA tensors)."""
    lib = _lib.load()
    dev = torch.device('cuda', torch.cuda.current_device() if device is None else int(device))

    def to_dev(x, dt):
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        return t.to(device=dev, dtype=dt).contiguous()

    with torch.cuda.device(dev):
        off = to_dev(offsets, torch.int64)
        lab = to_dev(labels, torch.int32)
        sc = to_dev(scores, torch.float64)
        n_groups, n = off.numel() - 1, lab.numel()
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        prec = torch.zeros(max(n_groups, 1), dtype=torch.float64, device=dev)
        _lib.check(lib.blp_eval_precision_at_k(ctypes.c_void_p(off.data_ptr()),
                                               ctypes.c_void_p(lab.data_ptr()),
                                               ctypes.c_void_p(sc.data_ptr()), n_groups,
                                               int(precision_at), ctypes.c_void_p(prec.data_ptr()),
                                               st), 'blp_eval_precision_at_k')
        counts = (ctypes.c_uint64 * 4)()
        _lib.check(lib.blp_eval_roc_auc(ctypes.c_void_p(lab.data_ptr()),
                                        ctypes.c_void_p(sc.data_ptr()), n, counts, st),
                   'blp_eval_roc_auc')
        total_precision = float(prec[:n_groups].cpu().numpy().sum()) if n_groups else 0.0
    n_pos, n_neg, greater, equal = (int(c) for c in counts)
    if n_pos == 0 or n_neg == 0:
        raise ValueError('Only one class present in y_true. ROC AUC score is not defined in '
                         'that case.')          # what sklearn raises at eval.py:26
    auc = (greater + 0.5 * equal) / (float(n_pos) * float(n_neg))
    return total_precision / n_example_users, auc


def run_evaluation(examples, methods, precision_at=20, data_dir='./data/test/', quiet=False):
    results = {}
    for method in methods:
        predictions = util.load_json(data_dir + method + '.json')
        off, lab, sc = flatten(examples, predictions)
        p, auc = metrics(off, lab, sc, len(examples), precision_at)
        results[method] = {'precision_at_%d' % precision_at: p, 'roc_auc': auc}
        if not quiet:
            print('Method:', method)
            print('  Precision @{:} = {:.4f}'.format(precision_at, p))
            print('  ROC Auc = {:.4f}'.format(auc))
    return results
