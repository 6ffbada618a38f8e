"""Compare two logits dumps (oracle/harness/logits_dump.cpp): relative error on logits and greedy (argmax) agreement."""
import sys
import numpy as np


def load(p):
    raw = np.fromfile(p, np.uint8)
    rows, V = raw[:8].view(np.int32)
    return raw[8:].view(np.float32).reshape(rows, V)


def compare(a, b):
    assert a.shape == b.shape
    scale = np.abs(a).max(axis=1)
    rel = np.abs(a - b).max(axis=1) / scale
    am_a, am_b = a.argmax(1), b.argmax(1)
    srt = np.sort(a, axis=1)
    margin = (srt[:, -1] - srt[:, -2]) / scale
    dis = np.nonzero(am_a != am_b)[0]
    return {"rows": int(a.shape[0]), "vocab": int(a.shape[1]), "max_rel_err": float(rel.max()), "mean_rel_err": float(rel.mean()),
            "argmax_agree": int((am_a == am_b).sum()), "disagree_rows": [int(d) for d in dis[:10]],
            "min_rel_margin": float(margin.min()), "margins_at_disagreement": [float(margin[d]) for d in dis[:10]],
            "rel_err_at_disagreement": [float(rel[d]) for d in dis[:10]]}


if __name__ == "__main__":
    import json
    print(json.dumps(compare(load(sys.argv[1]), load(sys.argv[2]))))
