"""JSON helpers with the reference's on-disk formats (util.py:12-21 of the reference).

``examples.json``: ``{"<user>": {"<business>": 0|1}}``; score files: ``{"<user>": {"<business>": n}}``
-- whole-file JSON, string keys of the shared int id space, ints as bare ints and floats in
shortest-repr form (what ``json.dumps`` emits).
"""
import json


def load_json(fname):
    with open(fname) as fh:
        return json.load(fh)


def write_json(d, fname):
    with open(fname, 'w') as fh:
        fh.write(json.dumps(d))


def write_edge_list(fname, ids_u, ids_b):
    """``graph.txt``: one "<user_id> <business_id>" line per review (dataset_maker.py:197)."""
    import numpy as np
    arr = np.stack([np.asarray(ids_u, dtype=np.int64), np.asarray(ids_b, dtype=np.int64)], axis=1)
    np.savetxt(fname, arr, fmt='%d %d')
