"""The k-permutation of the tensor-core GEMV experiment (csrc/nf4_kernels.cu, nf4_gemv_mma_kernel), emulated on the CPU.

The kernel feeds decoded weights straight from registers into mma.sync.m16n8k16: lane (g = lane / 4, t = lane % 4) holds
the t-th 32-weight chunk of rows g and g + 8 of a 128-weight k-step, and MMA j of the step contracts elements 4j..4j+3
of every lane's chunk, the B fragment being loaded from x with the same permutation.  This test restates the PTX
fragment layouts (A row-major 16x16, B col-major 16x8, C 16x8) in numpy, runs the kernel's loop structure lane by lane,
and checks the result against x @ W^T -- i.e. that the sum really runs over the true k.  It checks the index mapping,
not the CUDA code itself (that needs a GPU: B2Q_GEMV_CFG=3 with the GPU GEMV parity test).
"""
import numpy as np
import pytest


def mma_m16n8k16(a_frag, b_frag, c_frag):
    """PTX ISA layouts: a_frag[lane][reg 0..3][2], b_frag[lane][reg 0..1][2], c_frag[lane][0..3] (accumulated in place)."""
    A = np.zeros((16, 16))
    B = np.zeros((16, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for reg in range(4):
            for e in range(2):
                A[g + 8 * (reg & 1), 2 * t + e + 8 * (reg >> 1)] = a_frag[lane][reg][e]
        for reg in range(2):
            for e in range(2):
                B[2 * t + e + 8 * reg, g] = b_frag[lane][reg][e]
    D = A @ B
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for i in range(4):
            c_frag[lane][i] += D[g + 8 * (i >> 1), 2 * t + (i & 1)]


def gemv_like_the_kernel(x, W, M):
    N, K = W.shape
    chunks = K // 32
    steps = (chunks + 3) // 4
    y = np.zeros((M, N))
    for n0 in range(0, N, 16):                       # one block = 16 weight rows
        acc = np.zeros((32, 4))                      # the 8 warps' partial sums, already added
        for s in range(steps):                       # k-steps (the kernel deals them to its 8 warps round-robin)
            a_fr = np.zeros((8, 32, 4, 2))
            b_fr = np.zeros((8, 32, 2, 2))
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                c = 4 * s + t
                da, db, xr = np.zeros(32), np.zeros(32), np.zeros(32)
                if c < chunks:
                    if n0 + g < N:
                        da = W[n0 + g, c * 32:(c + 1) * 32]
                    if n0 + g + 8 < N:
                        db = W[n0 + g + 8, c * 32:(c + 1) * 32]
                    if g < M:
                        xr = x[g, c * 32:(c + 1) * 32]
                for j in range(8):                   # registers da[2j] = elements (4j, 4j+1), da[2j+1] = (4j+2, 4j+3)
                    a_fr[j, lane] = [[da[4 * j], da[4 * j + 1]], [db[4 * j], db[4 * j + 1]],
                                     [da[4 * j + 2], da[4 * j + 3]], [db[4 * j + 2], db[4 * j + 3]]]
                    b_fr[j, lane] = [[xr[4 * j], xr[4 * j + 1]], [xr[4 * j + 2], xr[4 * j + 3]]]
            for j in range(8):
                mma_m16n8k16(a_fr[j], b_fr[j], acc)
        for idx in range(128):                       # the epilogue's (lane, element) -> (row, token) map
            ln, i = idx >> 2, idx & 3
            row, token = n0 + (ln >> 2) + 8 * (i >> 1), 2 * (ln & 3) + (i & 1)
            if token < M and row < N:
                y[token, row] = acc[ln][i]
    return y


@pytest.mark.parametrize("M,N,K", [(1, 32, 128), (5, 40, 192), (8, 24, 320)])
def test_fragment_mapping_sums_over_the_true_k(M, N, K):
    rng = np.random.default_rng(M + N + K)
    x = rng.standard_normal((M, K))
    W = rng.standard_normal((N, K))
    assert np.abs(gemv_like_the_kernel(x, W, M) - x @ W.T).max() < 1e-12
