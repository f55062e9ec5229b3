"""Generates tests/golden/dropout_mask_golden.json: packed bits and counts of the LoRA-dropout keep mask for a few
(shape, seed, p).  The mask is this repo's own definition (oracle/dropout.py; torch's Philox stream is not reproducible
bit for bit), so the fixture pins that definition against drift -- product (csrc/b2q_internal.h), oracle and fixture must
change together.  Run from the repo root:  python oracle/make_dropout_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dropout  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dropout_mask_golden.json")
CASES = [((64, 256), 1234, 0.05), ((333, 768), 0x3FFFFFFFFFFFFFF1, 0.05), ((16, 64), 7, 0.5), ((8, 4096), 2**40 + 3, 0.1)]


def main():
    out = []
    for shape, seed, p in CASES:
        m = dropout.keep_mask(shape, seed, p)
        out.append({"shape": list(shape), "seed": seed, "p": p, "kept": int(m.sum()),
                    "first_row_hex": np.packbits(m.reshape(shape)[0][:64]).tobytes().hex(),
                    "sha256": hashlib.sha256(np.packbits(m).tobytes()).hexdigest()})
    json.dump({"definition": "oracle/dropout.py keep_mask; bits packed MSB first (numpy.packbits)", "cases": out},
              open(OUT, "w"), indent=1)
    print(OUT)


if __name__ == "__main__":
    main()
