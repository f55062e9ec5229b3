"""The plain-C restatement (oracle/nf4_ref.c) and the numpy oracle agree bit for bit."""
import ctypes as ct

import numpy as np
import pytest

from oracle import build_c, nf4


@pytest.fixture(scope="module")
def cref():
    lib = ct.CDLL(build_c.build())
    lib.nf4ref_quantize.argtypes = [ct.c_void_p, ct.c_int64, ct.c_void_p, ct.c_void_p]
    lib.nf4ref_dequantize_bf16.argtypes = [ct.c_void_p] * 5 + [ct.c_float, ct.c_void_p, ct.c_int64, ct.c_void_p]
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ct.c_void_p)


@pytest.mark.parametrize("dq", [False, True])
def test_c_matches_numpy(cref, dq):
    rng = np.random.default_rng(5)
    W = rng.normal(0, 0.02, (96, 512)).astype(np.float32)
    W[7, 128:192] = 0.0
    W[::9, ::11] *= 30
    st = nf4.quantize_nf4(W, 64, dq)
    n = W.size
    packed = np.zeros(n // 2, np.uint8)
    absmax = np.zeros(n // 64, np.float32)
    cref.nf4ref_quantize(_p(W), n, _p(packed), _p(absmax))
    assert np.array_equal(packed, st["packed"])
    if not dq:
        assert np.array_equal(absmax, st["absmax"])
    out = np.zeros(n, np.uint16)
    cref.nf4ref_dequantize_bf16(_p(st["packed"]), _p(st.get("absmax")), _p(st.get("absmax_q")), _p(st.get("absmax2")),
                                _p(st.get("code256")), float(st.get("offset", 0.0)), _p(st["code"]), n, _p(out))
    assert np.array_equal(out.reshape(W.shape), nf4.dequantize_nf4(st))
