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
    assert L.zb_mma_check(G.Q8_0, 64, 1024) != 0      # other formats stay on the CUDA-core kernel
    assert L.zb_mma_check(G.Q4_0, 64, 1152) == 0 and L.zb_mma_check(G.Q4_0, 64, 1120) != 0   # Q4_0 units are 128 weights
    assert L.zb_mma_check(G.Q4_K, 128256, 3072) == 0 and L.zb_mma_check(G.Q4_K, 3072, 8192) == 0


def _walk_q6k(tiles: np.ndarray, rows: int, K: int, x: np.ndarray) -> np.ndarray:
    """gemv_q6k.cu:11-25 semantics, addressed the way block_tile_q6k addresses a block-tile."""
    nb = K // 256
    n_tiles = (rows + 15) // 16
    y = np.zeros(n_tiles * 16, np.float64)
    for tau in range(n_tiles):
        for b in range(nb):
            bt = tiles[(tau * nb + b) * 3360:(tau * nb + b + 1) * 3360]
            xb = x[b * 256:(b + 1) * 256].astype(np.float64)
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                for h in range(2):
                    row = tau * 16 + g + 8 * h
                    A = bt[((h * 3 + 0) * 32 + lane) * 16:][:16]
                    B = bt[((h * 3 + 1) * 32 + lane) * 16:][:16]
                    H = bt[((h * 3 + 2) * 32 + lane) * 16:][:16]
                    sc = bt[3072 + (h * 8 + g) * 16:][:16].view(np.int8)
                    d = float(bt[3328 + (h * 8 + g) * 2:][:2].view(np.float16)[0])
                    acc = 0.0
                    for hf in range(2):
                        for i_s in range(2):
                            c = hf * 2 + i_s
                            for qi in range(4):
                                sg = 8 * hf + 2 * qi + i_s
                                for k in range(4):
                                    w = int((B if (qi & 1) else A)[4 * c + k])
                                    hb = int(H[4 * c + k])
                                    q = ((w >> 4) if qi >= 2 else (w & 15)) | (((hb >> (2 * qi)) & 3) << 4)
                                    acc += d * int(sc[sg]) * (q - 32) * xb[16 * sg + 4 * t + k]
                    y[row] += acc
    return y[:rows]


@pytest.mark.parametrize("rows,K", [(32, 512), (19, 256)])
def test_mma_repack_q6k_walk(rows, K):
    L = lib.load()
    rng = np.random.default_rng(6)
    raw = G.quantize(rng.standard_normal((rows, K), dtype=np.float32) * np.float32(0.05), G.Q6_K)
    x = rng.standard_normal(K, dtype=np.float32)
    wb, sb = C.c_int64(), C.c_int64()
    assert L.zb_mma_layout(G.Q6_K, rows, K, C.byref(wb), C.byref(sb)) == 0
    assert wb.value == ((rows + 15) // 16) * (K // 256) * 3360
    out = np.zeros(wb.value, np.uint8)
    rawb = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
    assert L.zb_mma_repack_host(G.Q6_K, rawb.ctypes.data, rows, K, out.ctypes.data) == 0
    assert np.array_equal(np.sort(out[out != 0]), np.sort(rawb[rawb != 0]))
    got = _walk_q6k(out, rows, K, x)
    ref = O.gemv_f64(G.Q6_K, raw, rows, K, x)
    assert np.allclose(got, ref, rtol=1e-9, atol=1e-9)


def _walk_q40(tiles: np.ndarray, rows: int, K: int, x: np.ndarray) -> np.ndarray:
    """Q4_0 (q4dot.go:10-29: low nibble = element j, high nibble = element j + 16, w = (q - 8) * d), addressed like block_tile_q40."""
    nb = K // 128
    n_tiles = (rows + 15) // 16
    y = np.zeros(n_tiles * 16, np.float64)
    for tau in range(n_tiles):
        for b in range(nb):
            bt = tiles[(tau * nb + b) * 1152:(tau * nb + b + 1) * 1152]
            xb = x[b * 128:(b + 1) * 128].astype(np.float64)
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                for h in range(2):
                    row = tau * 16 + g + 8 * h
                    words = bt[(h * 32 + lane) * 16:][:16]
                    acc = 0.0
                    for bi in range(4):
                        d = float(bt[1024 + g * 16 + h * 8 + bi * 2:][:2].view(np.float16)[0])
                        for k in range(4):
                            byte = int(words[4 * bi + k])
                            acc += d * ((byte & 15) - 8) * xb[32 * bi + 4 * t + k] + d * ((byte >> 4) - 8) * xb[32 * bi + 16 + 4 * t + k]
                    y[row] += acc
    return y[:rows]


@pytest.mark.parametrize("rows,K", [(32, 256), (21, 1152)])
def test_mma_repack_q40_walk(rows, K):
    L = lib.load()
    rng = np.random.default_rng(7)
    raw = G.quantize(rng.standard_normal((rows, K), dtype=np.float32) * np.float32(0.05), G.Q4_0)
    x = rng.standard_normal(K, dtype=np.float32)
    wb, sb = C.c_int64(), C.c_int64()
    assert L.zb_mma_layout(G.Q4_0, rows, K, C.byref(wb), C.byref(sb)) == 0
    assert wb.value == ((rows + 15) // 16) * (K // 128) * 1152
    out = np.zeros(wb.value, np.uint8)
    rawb = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
    assert L.zb_mma_repack_host(G.Q4_0, rawb.ctypes.data, rows, K, out.ctypes.data) == 0
    assert np.array_equal(np.sort(out[out != 0]), np.sort(rawb[rawb != 0]))
    got = _walk_q40(out, rows, K, x)
    ref = O.gemv_f64(G.Q4_0, raw, rows, K, x)
    assert np.allclose(got, ref, rtol=1e-9, atol=1e-9)


def test_mma_repack_q5k_is_q4k_tile_plus_high_bits():
    """Q5_K block-tile = the Q4_K tile of the same header / nibble bytes followed by [h][lane][8 B] of qh (gemv_q5k.cu:15-23)."""
    L = lib.load()
    rows, K = 24, 512
    rng = np.random.default_rng(8)
    raw = np.ascontiguousarray(G.quantize(rng.standard_normal((rows, K), dtype=np.float32) * np.float32(0.05), G.Q5_K)).view(np.uint8).reshape(rows, K // 256, 176)
    wb, sb = C.c_int64(), C.c_int64()
    assert L.zb_mma_layout(G.Q5_K, rows, K, C.byref(wb), C.byref(sb)) == 0 and wb.value == 2 * 2 * 2816
    out = np.zeros(wb.value, np.uint8)
    flat = raw.reshape(-1)
    assert L.zb_mma_repack_host(G.Q5_K, flat.ctypes.data, rows, K, out.ctypes.data) == 0
    assert np.array_equal(np.sort(out[out != 0]), np.sort(flat[flat != 0]))
    q4 = np.ascontiguousarray(raw[:, :, :144]).reshape(-1)
    out4 = np.zeros(2 * 2 * 2304, np.uint8)
    assert L.zb_mma_repack_host(G.Q4_K, q4.ctypes.data, rows, K, out4.ctypes.data) == 0
    tiles, tiles4 = out.reshape(4, 2816), out4.reshape(4, 2304)
    assert np.array_equal(tiles[:, :2304], tiles4)
    for tau in range(2):
        for b in range(2):
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                for h in range(2):
                    row = tau * 16 + g + 8 * h
                    want = raw[row, b, 144 + 8 * t:144 + 8 * t + 8] if row < rows else np.zeros(8, np.uint8)
                    assert np.array_equal(tiles[tau * 2 + b, 2304 + (h * 32 + lane) * 8:][:8], want)


GEOM_SHAPES = [(G.Q4_K, 5120, 3072), (G.Q4_K, 3072, 3072), (G.Q4_K, 16384, 3072), (G.Q4_K, 3072, 8192), (G.Q6_K, 128256, 3072),
               (G.Q6_K, 1024, 3072), (G.Q5_K, 28672, 4096), (G.Q5_K, 4096, 14336), (G.Q4_0, 262144, 1152), (G.Q4_0, 1152, 6912),
               (G.Q4_K, 16, 256), (G.Q4_K, 40, 512), (G.Q4_K, 57344, 8192), (G.Q4_K, 1031, 3072), (G.Q6_K, 257, 2048)]


@pytest.mark.parametrize("qt,rows,K", GEOM_SHAPES, ids=[f"{G.TYPE_NAMES[q]}-{m}x{k}" for q, m, k in GEOM_SHAPES])
@pytest.mark.parametrize("max_ctas", [148, 74])
def test_mma_work_split_invariants(qt, rows, K, max_ctas):
    """The launch geometry of the tensor-core GEMV, checked on the host: every block-tile is owned by exactly one (CTA, warp),
    a row tile is shared by at most 8 CTAs, the per-tile partial-sum slots and the per-CTA tile table are large enough, the ring fits
    shared memory (csrc/gemv_mma.cu: make_mgeom and the index arithmetic of gemv_mma_kernel)."""
    L = lib.load()
    out = (C.c_int * 12)()
    assert L.zb_mma_geometry(qt, rows, K, max_ctas, out) == 0
    nb, n_tiles, total, per_cta, ctas, per_warp, chunk, stages, slots, max_local, smem, bt = list(out)
    W = 16
    assert nb == K // (128 if qt == G.Q4_0 else 256) and n_tiles == (rows + 15) // 16 and total == nb * n_tiles
    assert 1 <= ctas <= max_ctas and (ctas - 1) * per_cta < total <= ctas * per_cta
    assert per_warp * W >= per_cta and (per_warp - 1) * W < per_cta
    assert stages >= 2 and 1 <= chunk <= 4 and smem <= 225 * 1024 - 2048 and W * stages * chunk * bt <= smem
    owner = np.full(total, -1, np.int64)
    for c in range(ctas):
        i0, i1 = c * per_cta, min(total, (c + 1) * per_cta)
        tiles = set()
        for w in range(W):
            r0, r1 = i0 + w * per_warp, min(i1, i0 + (w + 1) * per_warp)
            if r1 > r0:
                assert np.all(owner[r0:r1] == -1)
                owner[r0:r1] = c * W + w
                tiles.update(range(r0 // nb, (r1 - 1) // nb + 1))
        assert len(tiles) <= max_local
        for t in tiles:   # warps of this CTA that touch row tile t are consecutive and fit the slot table
            lo, hi = max(i0, t * nb), min(i1, (t + 1) * nb)
            assert (hi - 1 - i0) // per_warp - (lo - i0) // per_warp + 1 <= slots
    assert np.all(owner >= 0)
    first = (np.arange(n_tiles) * nb) // per_cta
    last = ((np.arange(n_tiles) + 1) * nb - 1) // per_cta
    assert np.all(last - first + 1 <= 8)
    wb, sb = C.c_int64(), C.c_int64()
    assert L.zb_mma_layout(qt, rows, K, C.byref(wb), C.byref(sb)) == 0
    assert wb.value == total * bt and sb.value >= n_tiles * 8 * 16 * 8


def test_repack_overwrites_every_tile_byte_and_threads_do_not_change_it(monkeypatch):
    """zb_mma_repack_host no longer clears the whole output first: every byte of every tile (ragged last row tile included)
    must come from the repack itself, single- or multi-threaded (zb_parallel_for)."""
    import ctypes as C
    from zerfoo_b200 import lib
    L = lib.load()
    rng = np.random.default_rng(1)
    tile = {G.Q4_K: 2304, G.Q5_K: 2816, G.Q6_K: 3360, G.Q4_0: 1152}
    for qt in tile:
        for rows, cols in ((1000, 1024), (2048, 2048), (37, 512)):
            rb = cols // G.BLOCK_ELEMS[qt] * G.BLOCK_BYTES[qt]
            raw = rng.integers(0, 256, rows * rb, dtype=np.uint8)
            wb, sb = C.c_int64(), C.c_int64()
            assert L.zb_mma_layout(qt, rows, cols, C.byref(wb), C.byref(sb)) == 0
            clean, dirty = np.zeros(wb.value, np.uint8), np.full(wb.value, 0xAB, np.uint8)
            monkeypatch.setenv("ZB_HOST_THREADS", "1")
            assert L.zb_mma_repack_host(qt, raw.ctypes.data, rows, cols, clean.ctypes.data) == 0
            monkeypatch.setenv("ZB_HOST_THREADS", "8")
            assert L.zb_mma_repack_host(qt, raw.ctypes.data, rows, cols, dirty.ctypes.data) == 0
            used = ((rows + 15) // 16) * (cols // (128 if qt == G.Q4_0 else 256)) * tile[qt]
            assert np.array_equal(clean[:used], dirty[:used]), (G.TYPE_NAMES[qt], rows, cols)
