"""Fixture design for the greedy-parity tests: run the REFERENCE CPU backend (oracle/_ref/logits_dump,
greedy mode) over a range of prompt seeds and report, per seed, the smallest top-1/top-2 logit margin (relative to the largest
|logit| of the row) over the rows that choose a token.  With random-init weights the logits are near-gaussian, so a 128-step
stream always contains a few near-ties; the parity tests use the seed whose smallest margin is largest, and log it.
usage: python tools/scan_margins.py GGUF KV N_PROMPT N_GEN SEED0 SEED1"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def margins(path, n_prompt):
    raw = np.fromfile(path, np.uint8)
    rows, V = raw[:8].view(np.int32)
    a = raw[8:].view(np.float32).reshape(rows, V)[n_prompt - 1:]
    part = np.partition(a, V - 2, axis=1)
    return (part[:, -1] - part[:, -2]) / np.abs(a).max(1)


if __name__ == "__main__":
    gguf, kv, n_prompt, n_gen, s0, s1 = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "oracle", "_ref"), LOGITS_DUMP_GREEDY="1")
    for seed in range(s0, s1):
        env["LOGITS_DUMP_SEED"] = str(seed)
        out = "/tmp/scan_%d.bin" % os.getpid()
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "logits_dump"), gguf, out, "0", str(n_prompt), str(n_gen), kv, "1", str(os.cpu_count())],
                       env=env, check=True, capture_output=True)
        m = margins(out, n_prompt)
        print("seed %d min_margin %.5f 2nd %.5f median %.4f" % (seed, m.min(), np.sort(m)[1], np.median(m)), flush=True)
