"""Known-answer tests that pin the CPU oracle (SURVEY.md section 8c list; none exist upstream)."""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import nf4
from oracle.qlora import make_case, qlora_linear_fwd_bwd, rel_err

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_nf4_table_is_the_qlora_construction():
    # (1) table identity: quantiles of N(0,1), 8 positive + 7 negative + 0, normalised (bnb create_normal_map)
    from scipy.stats import norm

    off = 0.9677083
    v1 = norm.ppf(torch.linspace(off, 0.5, 9)[:-1]).tolist()
    v3 = (-norm.ppf(torch.linspace(off, 0.5, 8)[:-1])).tolist()
    v = torch.Tensor(v1 + [0] + v3).sort().values
    v /= v.max()
    assert np.array_equal(v.numpy(), nf4.NF4_CODE)


def test_thresholds_are_midpoints():
    mid = (nf4.NF4_CODE[1:].astype(np.float64) + nf4.NF4_CODE[:-1].astype(np.float64)) / 2
    assert np.abs(mid - nf4.NF4_THRESHOLDS.astype(np.float64)).max() < 6e-8


def test_dynamic_map():
    c = nf4.create_dynamic_map()
    assert c.shape == (256,) and c.dtype == np.float32
    assert (np.diff(c) > 0).all()
    assert c[-1] == 1.0 and abs(c[0] + 0.99296875) < 1e-6 and (c == 0).sum() == 1


@pytest.mark.parametrize("a", [1.0, 0.02, 3.7e-3, 123.456])
def test_code_values_round_trip(a):
    # (2) decode(quantise(NF4[j] * a)) == bf16_rn(NF4[j] * a) for all 16 codes
    w = np.zeros(64, np.float32)
    w[:16] = nf4.NF4_CODE * np.float32(a)
    st = nf4.quantize_nf4(w, 64, False)
    assert st["absmax"][0] == np.float32(a)
    codes = nf4.unpack_codes(st["packed"], 64)
    assert list(codes[:16]) == list(range(16))
    bits = nf4.dequantize_nf4(st)
    assert np.array_equal(bits[:16], nf4.bf16_round(nf4.NF4_CODE * np.float32(a)))


def test_nibble_order_even_element_high():
    # (3) weights [NF4[15], NF4[0], ...] -> first byte 0xF0
    w = np.zeros(64, np.float32)
    w[0], w[1], w[2], w[3] = 1.0, -1.0, nf4.NF4_CODE[3], nf4.NF4_CODE[12]
    st = nf4.quantize_nf4(w, 64, False)
    assert st["packed"][0] == 0xF0 and st["packed"][1] == 0x3C


def test_threshold_edges_go_down():
    # (4) a value exactly on a mid-point maps to the lower code (strict >)
    w = np.zeros(64, np.float32)
    w[0] = 1.0  # absmax = 1 so normalised value == value
    w[1:16] = nf4.NF4_THRESHOLDS
    w[16:31] = np.nextafter(nf4.NF4_THRESHOLDS, np.float32(2.0))
    codes = nf4.unpack_codes(nf4.quantize_nf4(w, 64, False)["packed"], 64)
    assert list(codes[1:16]) == list(range(15))
    assert list(codes[16:31]) == list(range(1, 16))


def test_all_zero_block():
    # (5) all-zero block -> codes 0, absmax 0, decode compares equal to 0
    w = np.zeros(128, np.float32)
    w[64:] = 0.5
    st = nf4.quantize_nf4(w, 64, False)
    assert st["absmax"][0] == 0 and (nf4.unpack_codes(st["packed"], 128)[:64] == 0).all()
    assert (nf4.dequantize_nf4(st, as_bits=False)[:64] == 0).all()


def test_double_quant_op_order_and_bound():
    # (6) nested absmax = fl32(fl32(code256[q]*absmax2) + offset); error bound of the 8-bit code
    rng = np.random.default_rng(0)
    W = rng.normal(0, 0.02, (256, 1024)).astype(np.float32)
    st = nf4.quantize_nf4(W, 64, True)
    plain = nf4.quantize_nf4(W, 64, False)
    assert np.array_equal(st["packed"], plain["packed"])
    am = nf4.dequantize_absmax(st)
    q, j = st["absmax_q"], np.arange(st["absmax_q"].size) // 256
    manual = (st["code256"][q] * st["absmax2"][j]).astype(np.float32) + st["offset"]
    assert np.array_equal(am, manual.astype(np.float32))
    fused = (st["code256"][q].astype(np.float64) * st["absmax2"][j].astype(np.float64) + float(st["offset"]))
    assert not np.array_equal(am, fused.astype(np.float32)) or True  # documents: no FMA contraction assumed
    assert np.abs(am - plain["absmax"]).max() <= 0.02 * np.abs(plain["absmax"] - st["offset"]).max() + 1e-7


def test_bf16_round_matches_torch():
    x = np.random.default_rng(1).normal(0, 1, 4096).astype(np.float32)
    ours = nf4.bf16_round(x).astype(np.int16)
    theirs = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy()
    assert np.array_equal(ours, theirs)


def test_lora_b_zero_is_base_only():
    # (7) B = 0 -> output equals base output bit for bit and dA = 0
    case = make_case(32, 64, 128, 16, seed=3, lora_b_zero=True)
    o = qlora_linear_fwd_bwd(case["x"], case["state"], case["A"], case["B"], 0.25, case["dy"], mode="bf16")
    W = torch.from_numpy(nf4.dequantize_nf4(case["state"], as_bits=False).copy())
    base = (case["x"].float() @ W.t()).bfloat16().float()
    assert torch.equal(o["y"], base)
    assert float(o["dA"].abs().max()) == 0.0


def test_fp64_autograd_agrees_with_explicit_backward():
    # (8) gradcheck-style: the explicit backward formulas equal autograd of the forward in fp64
    case = make_case(8, 64, 64, 8, seed=4, p=0.25, dtype=torch.float32)
    W = torch.from_numpy(nf4.dequantize_nf4(case["state"], as_bits=False).copy()).double()
    x = case["x"].double().requires_grad_(True)
    A = case["A"].double().requires_grad_(True)
    B = case["B"].double().requires_grad_(True)
    mask = case["mask"].double()
    s, p = 0.25, 0.25
    y = x @ W.t() + s * (((x * mask / (1 - p)) @ A.t()) @ B.t())
    dy = case["dy"].double()
    y.backward(dy)
    o = qlora_linear_fwd_bwd(case["x"], case["state"], case["A"], case["B"], s, case["dy"], case["mask"], p, "fp32")
    assert rel_err(o["y"], y.detach()) < 1e-5
    assert rel_err(o["dx"], x.grad) < 1e-5
    assert rel_err(o["dA"], A.grad) < 1e-5
    assert rel_err(o["dB"], B.grad) < 1e-5


def test_bf16_mode_within_tolerance_of_fp32():
    case = make_case(64, 128, 256, 16, seed=5)
    a = qlora_linear_fwd_bwd(case["x"], case["state"], case["A"], case["B"], 0.25, case["dy"], mode="bf16")
    b = qlora_linear_fwd_bwd(case["x"], case["state"], case["A"], case["B"], 0.25, case["dy"], mode="fp32")
    for k in ("y", "dx", "dA", "dB"):
        assert rel_err(a[k], b[k]) < 2e-2, k


def test_dp_mean_of_rank_grads_equals_full_batch():
    # (10) mean over ranks of per-rank grads (each with dy scaled as a mean loss would) == single-process grad
    case = make_case(64, 64, 128, 8, seed=6, dtype=torch.float32)
    full = qlora_linear_fwd_bwd(case["x"], case["state"], case["A"], case["B"], 0.25, case["dy"], mode="fp32")
    parts = []
    for r in range(2):
        sl = slice(r * 32, (r + 1) * 32)
        parts.append(qlora_linear_fwd_bwd(case["x"][sl], case["state"], case["A"], case["B"], 0.25,
                                          case["dy"][sl] * 2.0, mode="fp32"))
    mean_dA = (parts[0]["dA"] + parts[1]["dA"]) / 2
    assert rel_err(mean_dA, full["dA"]) < 1e-5


def test_golden_vectors():
    """Committed fixtures (tests/golden/, made by oracle/make_golden.py) pin the oracle against drift."""
    path = os.path.join(GOLDEN, "nf4_golden.npz")
    meta = json.load(open(os.path.join(GOLDEN, "nf4_golden.json")))
    g = np.load(path)
    for name, m in meta["cases"].items():
        W = g[f"{name}.w"]
        st = nf4.quantize_nf4(W, 64, m["double_quant"])
        assert np.array_equal(st["packed"], g[f"{name}.packed"]), name
        assert np.array_equal(nf4.dequantize_nf4(st), g[f"{name}.decoded_bits"]), name
        if m["double_quant"]:
            assert np.array_equal(st["absmax_q"], g[f"{name}.absmax_q"]), name
            assert np.array_equal(st["absmax2"], g[f"{name}.absmax2"]), name
            assert np.float32(st["offset"]) == g[f"{name}.offset"], name
        else:
            assert np.array_equal(st["absmax"], g[f"{name}.absmax"]), name
    lin = meta["linear"]
    case = make_case(lin["M"], lin["N"], lin["K"], lin["r"], seed=lin["seed"], double_quant=True)
    o = qlora_linear_fwd_bwd(case["x"], case["state"], case["A"], case["B"], lin["scale"], case["dy"], mode="bf16")
    for k in ("y", "dx", "dA", "dB"):
        gold = torch.from_numpy(g[f"linear.{k}"]).view(torch.bfloat16).float()
        # fp32 BLAS summation order may differ between hosts: allow one bf16 ulp of the largest value
        assert rel_err(o[k], gold) <= 2.0 ** -7, k


def test_dropout_mask_restatement_matches_the_product_source_and_is_well_behaved():
    """oracle/dropout.py against the host compilation of csrc/b2q_internal.h's dropout_keep (the source the device
    runs), plus the statistics a dropout mask needs (rate, no row / column / neighbour structure)."""
    import subprocess

    from oracle import build_c, dropout

    exe = build_c.build_dropout_ref()
    for seed, p, n in ((0, 0.05, 4096), (1234, 0.05, 5000), (0x3FFFFFFFFFFFFFFF, 0.5, 3000), (0x123456789ABCDEF, 0.01, 2048)):
        out = subprocess.run([exe, str(seed), str(p), str(n)], capture_output=True, text=True, check=True).stdout.strip()
        ref = np.frombuffer(out.encode(), dtype=np.uint8) - ord("0")
        assert np.array_equal(ref, dropout.keep_mask((n,), seed, p)), (seed, p)
    m = dropout.keep_mask((1024, 4096), 0x1234ABCD5678, 0.05).astype(np.float64)
    d = 1.0 - m
    assert abs(d.mean() - 0.05) < 5e-4
    assert d.mean(1).std() < 1.5 * np.sqrt(0.05 * 0.95 / 4096) and d.mean(0).std() < 1.5 * np.sqrt(0.05 * 0.95 / 1024)
    dc = d - d.mean()
    for dr, dcol in ((0, 1), (0, 2), (0, 4), (0, 8), (1, 0)):
        a = dc[: d.shape[0] - dr, : d.shape[1] - dcol]
        b = dc[dr:, dcol:]
        assert abs(float((a * b).mean() / dc.var())) < 5e-3, (dr, dcol)


def test_dropout_mask_golden_fixture():
    """tests/golden/dropout_mask_golden.json (oracle/make_dropout_golden.py) pins the mask definition against drift."""
    import hashlib

    from oracle import dropout

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dropout_mask_golden.json")
    for c in json.load(open(path))["cases"]:
        m = dropout.keep_mask(tuple(c["shape"]), c["seed"], c["p"])
        assert int(m.sum()) == c["kept"]
        assert np.packbits(m.reshape(c["shape"])[0][:64]).tobytes().hex() == c["first_row_hex"]
        assert hashlib.sha256(np.packbits(m).tobytes()).hexdigest() == c["sha256"]


def test_quantise_properties_over_random_scales():
    """Property tests (hypothesis): for weights of any scale, (i) the decode error is bounded by the largest half-gap of
    the code book times the block's absmax, (ii) re-quantising the fp32 decode reproduces the same codes and absmax
    (the code book contains +-1, so absmax survives), (iii) the entry that set the absmax decodes to +-absmax exactly."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(0, 2**31 - 1), st.floats(-18.0, 6.0), st.integers(1, 6))
    def check(seed, log10_scale, nblocks):
        rng = np.random.default_rng(seed)
        w = (rng.standard_normal(64 * nblocks) * 10.0 ** log10_scale).astype(np.float32)
        stq = nf4.quantize_nf4(w, 64, False)
        dec = nf4.dequantize_nf4_f32(stq).reshape(-1)
        am = np.repeat(stq["absmax"], 64)
        half_gap = 0.5 * float(np.max(np.diff(nf4.NF4_CODE)))
        # fl32 rounding of w * (1/absmax) can move a value across a threshold by one ulp: allow a few ulps of slack
        assert np.all(np.abs(dec - w) <= am * (half_gap + 1e-6))
        st2 = nf4.quantize_nf4(dec, 64, False)
        assert np.array_equal(st2["packed"], stq["packed"]) and np.array_equal(st2["absmax"], stq["absmax"])
        blocks = w.reshape(nblocks, 64)
        idx = np.argmax(np.abs(blocks), axis=1)
        peak = dec.reshape(nblocks, 64)[np.arange(nblocks), idx]
        assert np.array_equal(np.abs(peak), stq["absmax"])

    check()


def test_absmax_mean_summation_order_does_not_move_the_decoded_weight():
    """bitsandbytes takes the double-quant offset with torch's CUDA `absmax.mean()` (fp32 partial sums in a device-dependent
    order); the oracle and the product return the correctly rounded mean.  Whatever the order, the offset moves by a few
    ulp at most, the 8-bit absmax codes change for at most 0.2 % of the blocks (19 of 16384 for a strictly sequential
    fp32 sum, the worst order tried; a handful for tree orders like torch's), and the decoded weight -- which uses the STORED
    offset -- stays within one step of the 8-bit absmax code of the value either way (oracle/nf4.py
    double_quantize_absmax).  So: decode parity with bitsandbytes does not depend on the order; quantiser BYTES may differ
    from bitsandbytes' own in a fraction of a percent of the absmax codes."""
    from oracle import nf4

    rng = np.random.default_rng(5)
    W = (rng.standard_normal((512, 2048)) * 0.02).astype(np.float32)
    st = nf4.quantize_nf4(W, 64, True)
    absmax = np.abs(W.reshape(-1, 64)).max(axis=1).astype(np.float32)
    means = {"fp64": np.float32(absmax.astype(np.float64).mean())}
    acc = np.float32(0.0)
    for v in absmax:                                   # fp32, sequential
        acc = np.float32(acc + v)
    means["fp32 sequential"] = np.float32(acc / np.float32(absmax.size))
    means["fp32 pairwise"] = np.float32(absmax.sum(dtype=np.float32) / np.float32(absmax.size))   # numpy's pairwise tree
    chunks = absmax.reshape(-1, 32).sum(axis=1, dtype=np.float32)                                   # warp-sized partials
    means["fp32 two-level"] = np.float32(chunks.sum(dtype=np.float32) / np.float32(absmax.size))
    assert st["offset"] == means["fp64"]
    ref = nf4.dequantize_nf4(st, as_bits=False)
    ulp = np.spacing(np.float32(means["fp64"]))
    for name, off in means.items():
        assert abs(float(off) - float(means["fp64"])) <= 64 * float(ulp), name
        # re-run the second-level quantisation with this offset
        centred = (absmax - off).astype(np.float32)
        blk2 = centred.reshape(-1, 256)
        a2 = np.abs(blk2).max(axis=1).astype(np.float32)
        q = nf4._nearest_code(st["code256"], (blk2 * (np.float32(1.0) / a2)[:, None]).astype(np.float32)).reshape(-1)
        changed = int((q != st["absmax_q"]).sum())
        assert changed <= absmax.size // 500, (name, changed)
        if "sequential" not in name:
            assert changed <= 8, (name, changed)
        st2 = dict(st, offset=off, absmax_q=q, absmax2=a2)
        got = nf4.dequantize_nf4(st2, as_bits=False)
        # decoded weights agree to within one quantisation step of the 8-bit absmax code (relative 1/127 of the local scale)
        assert float(np.abs(got - ref).max()) <= float(np.abs(ref).max()) * 2e-2, name
