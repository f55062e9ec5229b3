"""Generates tests/golden/nf4_golden.{npz,json} from the oracle (run from the repo root).

The reference holds no golden vectors for this path and bitsandbytes / peft cannot be imported
here, so these fixtures pin the ORACLE against drift; they are not
outputs of the reference.  PARITY UNPINNED.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nf4  # noqa: E402
from oracle.qlora import make_case, qlora_linear_fwd_bwd  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    arrays, meta = {}, {"cases": {}}
    rng = np.random.default_rng(20261018)
    specs = {
        "plain_64x256": ((64, 256), False),
        "nested_128x512": ((128, 512), True),
        "nested_zero_block": ((16, 256), True),
        "plain_outliers": ((32, 128), False),
    }
    for name, (shape, dq) in specs.items():
        W = rng.normal(0, 0.02, shape).astype(np.float32)
        if name == "nested_zero_block":
            W[2, 64:128] = 0.0
        if name == "plain_outliers":
            W[::5, ::7] *= 40.0
        st = nf4.quantize_nf4(W, 64, dq)
        arrays[f"{name}.w"] = W
        arrays[f"{name}.packed"] = st["packed"]
        arrays[f"{name}.decoded_bits"] = nf4.dequantize_nf4(st)
        if dq:
            arrays[f"{name}.absmax_q"] = st["absmax_q"]
            arrays[f"{name}.absmax2"] = st["absmax2"]
            arrays[f"{name}.offset"] = np.float32(st["offset"])
        else:
            arrays[f"{name}.absmax"] = st["absmax"]
        meta["cases"][name] = {"shape": list(shape), "double_quant": dq}
    lin = {"M": 48, "N": 128, "K": 256, "r": 16, "seed": 11, "scale": 0.25}
    case = make_case(lin["M"], lin["N"], lin["K"], lin["r"], seed=lin["seed"], double_quant=True)
    o = qlora_linear_fwd_bwd(case["x"], case["state"], case["A"], case["B"], lin["scale"], case["dy"], mode="bf16")
    for k in ("y", "dx", "dA", "dB"):
        arrays[f"linear.{k}"] = o[k].bfloat16().view(torch.int16).numpy()
    meta["linear"] = lin
    np.savez_compressed(os.path.join(OUT, "nf4_golden.npz"), **arrays)
    json.dump(meta, open(os.path.join(OUT, "nf4_golden.json"), "w"), indent=1)
    print("wrote", OUT, sum(a.nbytes for a in arrays.values()), "bytes")


if __name__ == "__main__":
    main()
