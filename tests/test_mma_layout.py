"""Host-side check of the tensor-core GEMV weight layout (zb_mma_repack_host, csrc/gemv_mma.cu): a numpy walk of the
block-tiles with the kernel's own lane addressing must reproduce the oracle's Q4_K GEMV (gemv_q4k.cu:38-56 semantics)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from zerfoo_b200 import gguf as G, lib


def _walk_q4k(tiles: np.ndarray, rows: int, K: int, x: np.ndarray) -> np.ndarray:
    nb = K // 256
    n_tiles = (rows + 15) // 16
    y = np.zeros(n_tiles * 16, np.float64)
    f16 = lambda b: float(np.frombuffer(bytes(b), np.float16)[0])
    for tau in range(n_tiles):
        for b in range(nb):
            bt = tiles[(tau * nb + b) * 2304:(tau * nb + b + 1) * 2304]
            xb = x[b * 256:(b + 1) * 256].astype(np.float64)
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                for h in range(2):
                    row = tau * 16 + g + 8 * h
                    hdr = bt[2048 + (h * 8 + g) * 16: 2048 + (h * 8 + g) * 16 + 16]
                    d, dmin = f16(hdr[0:2]), f16(hdr[2:4])
                    sc = hdr[4:16]
                    acc = 0.0
                    for Gp in range(4):
                        chunk = bt[((h * 2 + (Gp >> 1)) * 32 + lane) * 16:][:16]
                        by = chunk[8:16] if (Gp & 1) else chunk[0:8]
                        for nib in range(2):
                            j = 2 * Gp + nib
                            if j < 4:
                                s, m = sc[j] & 63, sc[4 + j] & 63
                            else:
                                s = (sc[4 + j] & 0xF) | ((sc[j - 4] >> 6) << 4)
                                m = (sc[4 + j] >> 4) | ((sc[j] >> 6) << 4)
                            for k in range(8):
                                q = (int(by[k]) >> (4 * nib)) & 15
                                acc += (d * s * q - dmin * m) * xb[64 * Gp + 32 * nib + 8 * t + k]
                    y[row] += acc
    return y[:rows]


@pytest.mark.parametrize("rows,K", [(32, 512), (23, 256)])
def test_mma_repack_q4k_walk(rows, K):
    L = lib.load()
    rng = np.random.default_rng(5)
    raw = G.quantize(rng.standard_normal((rows, K), dtype=np.float32) * np.float32(0.05), G.Q4_K)
    x = rng.standard_normal(K, dtype=np.float32)
    wb, sb = C.c_int64(), C.c_int64()
    assert L.zb_mma_layout(G.Q4_K, rows, K, C.byref(wb), C.byref(sb)) == 0
    assert wb.value == ((rows + 15) // 16) * (K // 256) * 2304 and sb.value > 0
    out = np.zeros(wb.value, np.uint8)
    rawb = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
    assert L.zb_mma_repack_host(G.Q4_K, rawb.ctypes.data, rows, K, out.ctypes.data) == 0
    # a pure byte permutation (plus zero rows): same multiset of bytes
    assert np.array_equal(np.sort(out[out != 0]), np.sort(rawb[rawb != 0]))
    got = _walk_q4k(out, rows, K, x)
    ref = O.gemv_f64(G.Q4_K, raw, rows, K, x)
    assert np.allclose(got, ref, rtol=1e-9, atol=1e-9)


def test_mma_check_rejects_unsupported():
    L = lib.load()
    assert L.zb_mma_check(G.Q4_K, 64, 1152) != 0      # K not a multiple of 256
    assert L.zb_mma_check(G.Q4_0, 64, 1024) != 0      # other formats stay on the CUDA-core kernel
    assert L.zb_mma_check(G.Q4_K, 128256, 3072) == 0 and L.zb_mma_check(G.Q4_K, 3072, 8192) == 0
