"""Per-node comparison of two LOGITS_DUMP_NODES files: prints relative L2 error per graph node in execution order."""
import sys
import numpy as np


def load(p):
    raw = open(p, "rb").read()
    off, out = 0, []
    while off < len(raw):
        name = raw[off:off + 96].split(b"\0")[0].decode(); off += 96
        n = int(np.frombuffer(raw[off:off + 8], np.int64)[0]); off += 8
        out.append((name, np.frombuffer(raw[off:off + 4 * n], np.float32))); off += 4 * n
    return out


a, b = load(sys.argv[1]), load(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
bi = 0
shown = 0
for name, va in a:
    # align by name (the two runs may split graphs differently)
    j = next((k for k in range(bi, min(bi + 50, len(b))) if b[k][0] == name and b[k][1].size == va.size), None)
    if j is None:
        continue
    vb = b[j][1]; bi = j + 1
    den = np.sqrt((va.astype(np.float64) ** 2).sum()) + 1e-30
    rel = np.sqrt(((va.astype(np.float64) - vb) ** 2).sum()) / den
    if rel >= thr:
        print("%-48s n=%-8d rel_l2=%.3e  max|a|=%.3g" % (name, va.size, rel, np.abs(va).max()))
        shown += 1
        if shown > 120:
            break
