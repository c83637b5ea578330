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


# ---------------------------------------------------------------------------------------------
# Columnar sidecar.  Nested JSON is tens of GB of text at 100M
# pairs; the same information as three flat arrays in an .npz is 16-20 B per pair and loads
# without parsing.  The JSON format stays the interchange format of the reference's readers
# (eval.py:14, supervised_models.py:81-86); these two functions convert losslessly both ways.
# ---------------------------------------------------------------------------------------------
def write_pairs_npz(fname, ids_u, ids_b, values):
    """One score (or label) file as columns: user id, business id, value.  Integer-valued files
    (cn, labels, literal zeros) are stored as int64, everything else as float64."""
    import numpy as np
    v = np.asarray(values)   # floats that happen to be whole stay float: jaccard 0.0 is not int 0
    np.savez(fname, u=np.asarray(ids_u, dtype=np.int64), b=np.asarray(ids_b, dtype=np.int64), v=v)


def read_pairs_npz(fname):
    import numpy as np
    z = np.load(fname)
    return z['u'], z['b'], z['v']


def dict_to_columns(d):
    """{"<u>": {"<b>": x}} -> (ids_u, ids_b, values, int_mask): int_mask remembers which values
    were Python ints (the reference mixes int 0 and floats in one file, similarity.py:59-60,118)."""
    import numpy as np
    us, bs, vs, im = [], [], [], []
    for u, row in d.items():
        for b, x in row.items():
            us.append(int(u))
            bs.append(int(b))
            vs.append(float(x))
            im.append(isinstance(x, int))
    return (np.asarray(us, dtype=np.int64), np.asarray(bs, dtype=np.int64),
            np.asarray(vs, dtype=np.float64), np.asarray(im, dtype=bool))


def columns_to_dict(ids_u, ids_b, values, int_mask=None):
    """Inverse of dict_to_columns: the nested dict json.dumps turns into the reference's file."""
    import numpy as np
    v = np.asarray(values)
    ints = v.dtype.kind in 'iu'
    out = {}
    im = None if int_mask is None else np.asarray(int_mask, dtype=bool).tolist()
    for i, (u, b, x) in enumerate(zip(np.asarray(ids_u).tolist(), np.asarray(ids_b).tolist(),
                                      v.tolist())):
        if ints or (im is not None and im[i]):
            x = int(x)
        out.setdefault(str(u), {})[str(b)] = x
    return out


def json_to_npz(json_file, npz_file):
    import numpy as np
    u, b, v, im = dict_to_columns(load_json(json_file))
    np.savez(npz_file, u=u, b=b, v=v, int_mask=im)


def npz_to_json(npz_file, json_file):
    import numpy as np
    z = np.load(npz_file)
    write_json(columns_to_dict(z['u'], z['b'], z['v'], z['int_mask'] if 'int_mask' in z else None),
               json_file)
