"""Device residency and several devices for one call (§8 b: mxg_set_devices), through the C ABI.

* explicit handles with HOST operands (mxg_csr_spmm_host / _t_host / mxg_csr_spmv_host, what the glue's as_gpu_csr
  objects call): same results as the level-1 exports and the oracle, and only the dense operand + the result move;
* the level-1 operand cache (option cache_mb): a repeated product on the same host arrays moves no CSR bytes and
  returns the same bits; in-place modification at a sampled position is noticed; LRU budget is respected;
* mxg_set_devices(n): one level-1 call spread over n GPUs by n host threads is bit-identical to n = 1
  (skipped on a 1-GPU box, where mxg_set_devices(2) must fail cleanly).
"""
import ctypes as C

import numpy as np
import pytest

from helpers import FP32_TOL, FP64_TOL, powerlaw_csr, rel_err, rsparsematrix

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rx():
    from matrixextra_b200 import rcpp_exports
    return rcpp_exports


@pytest.fixture()
def lib():
    from matrixextra_b200 import _lib
    yield _lib
    _lib.set_option("cache_mb", 0)
    _lib.call("mxg_cache_clear")
    _lib.call("mxg_set_devices", 1)
    _lib.set_option("multi_min_nnz", 4 << 20)
    _lib.set_option("multi_dense_share", 1)
    _lib.set_option("multi_pageable", 0)
    _lib.set_option("host_result_pool_mb", 4096)
    _lib.set_option("piece", 1024)
    _lib.set_option("pipe_chunk_nnz", 0)
    _lib.set_option("host_colsplit", 1)
    _lib.set_option("pipe_slots", 4)


def _bytes(lib):
    up, down = C.c_size_t(0), C.c_size_t(0)
    lib.call("mxg_last_call_bytes", C.byref(up), C.byref(down))
    return int(up.value), int(down.value)


def _tol(dt):
    return FP64_TOL if dt == np.float64 else FP32_TOL


def _sfx(dt):
    return "numeric" if dt == np.float64 else "float32"


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n", [64, 24, 7])
def test_handle_products_with_host_operands(rx, lib, port, dt, n):
    lib.set_option("piece", 64)  # rows longer than a piece: the long-row launch that precedes the row chunks
    m, K = 6000, 2500
    p, j, x = powerlaw_csr(m, K, 30, seed=5, cap=2000)
    h = rx.as_gpu_csr(p, j, x, K, keep_float64=True, keep_float32=True)
    try:
        rng = np.random.default_rng(n)
        Y = np.asfortranarray(rng.standard_normal((n, K)).astype(dt))  # (n x K): A %*% t(Y)
        got = getattr(rx, "gpu_csr_tcrossprod_dense_" + _sfx(dt))(h, Y)
        up, down = _bytes(lib)
        assert (up, down) == (Y.nbytes, m * n * Y.itemsize)  # the CSR did not move
        want = getattr(port, "tcrossprod_csr_dense_" + _sfx(dt))(p, j, x, Y)
        assert got.shape == (m, n) and got.flags.f_contiguous and rel_err(got, want) <= _tol(dt)
        # bit-identical to the level-1 export on the same operands (same kernels, other chunking)
        assert np.array_equal(got, getattr(rx, "tcrossprod_csr_dense_" + _sfx(dt))(p, j, x, Y))
        got2 = getattr(rx, "gpu_csr_dense_tcrossprod_" + _sfx(dt))(Y, h)  # X %*% t(A): rows-contiguous result
        assert got2.shape == (n, m) and np.array_equal(got2, np.asfortranarray(got.T))
        Z = np.asfortranarray(rng.standard_normal((m, n)).astype(dt))  # t(A) %*% Z through the device-built CSC
        got3 = getattr(rx, "gpu_csr_crossprod_dense_" + _sfx(dt))(h, Z)
        assert np.array_equal(got3, rx.crossprod_csr_dense(p, j, x, K, Z, lib.MXG_F64 if dt == np.float64 else lib.MXG_F32))
        got3b = getattr(rx, "gpu_csr_crossprod_dense_" + _sfx(dt))(h, Z)  # the cached CSC
        up, down = _bytes(lib)
        assert np.array_equal(got3, got3b) and (up, down) == (Z.nbytes, K * n * Z.itemsize)
        if dt == np.float64:
            y = rng.standard_normal(K)
            want_v = rx.matmul_csr_dvec_numeric(p, j, x, y)
            assert np.array_equal(rx.gpu_csr_dvec_numeric(h, y), want_v)
            assert _bytes(lib) == (8 * K, 8 * m)
        with pytest.raises(ValueError, match="Matrix dimensions do not match"):
            rx.gpu_csr_tcrossprod_dense_numeric(h, np.zeros((3, K + 1)))
    finally:
        rx.gpu_csr_free(h)
    with pytest.raises(RuntimeError, match="has been freed"):
        rx.gpu_csr_tcrossprod_dense_numeric(h, np.zeros((3, K)))


def test_handle_edge_cases_and_s4_dispatch(rx, lib, port):
    from matrixextra_b200 import as_gpu, crossprod, dgRMatrix, float32, matmul, tcrossprod
    # a matrix without stored entries, one with empty rows, page-locked operands, many small result chunks
    import torch
    p0 = np.zeros(41, dtype=np.int32)
    h = rx.as_gpu_csr(p0, np.zeros(0, np.int32), np.zeros(0), 9)
    out = rx.gpu_csr_tcrossprod_dense_numeric(h, np.ones((5, 9), order="F"))
    assert out.shape == (40, 5) and not out.any()
    rx.gpu_csr_free(h)
    S = rsparsematrix(3000, 700, 0.02, 3)
    A = dgRMatrix.from_scipy(S)
    G = as_gpu(A, float64=True, float32=True)
    rng = np.random.default_rng(8)
    B = np.asfortranarray(rng.standard_normal((700, 33)))
    want = matmul(A, B)
    assert np.array_equal(matmul(G, B), want)
    pin = torch.from_numpy(np.ascontiguousarray(B)).pin_memory().numpy()  # page-locked operand: DMA'd in place
    assert np.array_equal(matmul(G, pin), want)
    assert np.array_equal(tcrossprod(G, np.asfortranarray(B.T)), tcrossprod(A, np.asfortranarray(B.T)))
    X = np.asfortranarray(rng.standard_normal((17, 700)))
    assert np.array_equal(tcrossprod(X, G), tcrossprod(X, A))
    Y = np.asfortranarray(rng.standard_normal((3000, 6)))
    assert np.array_equal(crossprod(G, Y), crossprod(A, Y))
    Xm = np.asfortranarray(rng.standard_normal((6, 3000)))
    assert np.array_equal(matmul(Xm, G), matmul(Xm, A))
    Bf = float32(B.astype(np.float32))
    assert np.array_equal(matmul(G, Bf).Data, matmul(A, Bf).Data)
    v = rng.standard_normal(700)
    assert np.array_equal(matmul(G, v), matmul(A, v))
    with pytest.raises(ValueError, match="Matrix dimensions do not match"):
        matmul(G, np.zeros((701, 2)))
    G.free()
    # a handle made with float64 values only refuses float32 products (like the glue)
    G64 = as_gpu(A)
    with pytest.raises(RuntimeError, match="without values of this type"):
        matmul(G64, Bf)
    G64.free()


def test_level1_cache_reuses_the_device_matrix(rx, lib, port):
    m, K, n = 20000, 6000, 32
    p, j, x = powerlaw_csr(m, K, 25, seed=21, cap=3000)
    rng = np.random.default_rng(21)
    Y = np.asfortranarray(rng.standard_normal((n, K)))
    cold = rx.tcrossprod_csr_dense_numeric(p, j, x, Y)
    up_cold, down_cold = _bytes(lib)
    assert up_cold >= p.nbytes + j.nbytes + x.nbytes + Y.nbytes
    lib.set_option("cache_mb", 256)
    first = rx.tcrossprod_csr_dense_numeric(p, j, x, Y)  # streams the matrix in and keeps it
    assert _bytes(lib)[0] >= p.nbytes + j.nbytes + x.nbytes
    warm = rx.tcrossprod_csr_dense_numeric(p, j, x, Y)  # moves the dense operand (kept from now on) and the result
    assert _bytes(lib) == (Y.nbytes, m * n * 8)
    hot = rx.tcrossprod_csr_dense_numeric(p, j, x, Y)  # dense operand resident too
    assert _bytes(lib) == (0, m * n * 8)
    assert np.array_equal(cold, first) and np.array_equal(cold, warm) and np.array_equal(cold, hot)
    # another product on the same matrix: SpMV and the transposed product share the entry
    y = rng.standard_normal(K)
    assert rel_err(rx.matmul_csr_dvec_numeric(p, j, x, y), port.matmul_csr_dvec_numeric(p, j, x, y)) <= FP64_TOL
    assert _bytes(lib) == (8 * K, 8 * m)
    Z = np.asfortranarray(rng.standard_normal((m, 8)))
    t1 = rx.crossprod_csr_dense(p, j, x, K, Z)
    t2 = rx.crossprod_csr_dense(p, j, x, K, Z)
    lib.set_option("cache_mb", 0)
    assert np.array_equal(t1, t2) and np.array_equal(t1, rx.crossprod_csr_dense(p, j, x, K, Z))
    lib.set_option("cache_mb", 256)
    # float32 product on the same arrays: the entry holds float64 values only, so it is replaced, never misused
    Yf = Y.astype(np.float32)
    f1 = rx.tcrossprod_csr_dense_float32(p, j, x, Yf)
    assert rel_err(f1, port.tcrossprod_csr_dense_float32(p, j, x, Yf)) <= FP32_TOL
    assert np.array_equal(f1, rx.tcrossprod_csr_dense_float32(p, j, x, Yf))
    # modified in place at positions the fingerprint samples (both ends): noticed, the matrix is streamed again
    x[0] += 1.0
    x[-1] -= 1.0
    again = rx.tcrossprod_csr_dense_numeric(p, j, x, Y)
    assert _bytes(lib)[0] >= j.nbytes + x.nbytes
    assert rel_err(again, port.tcrossprod_csr_dense_numeric(p, j, x, Y)) <= FP64_TOL and not np.array_equal(again, cold)
    Y[0, 0] += 1.0  # the dense operand too
    again2 = rx.tcrossprod_csr_dense_numeric(p, j, x, Y)
    assert rel_err(again2, port.tcrossprod_csr_dense_numeric(p, j, x, Y)) <= FP64_TOL
    hits, misses, nbytes, entries = C.c_ulonglong(), C.c_ulonglong(), C.c_size_t(), C.c_int()
    lib.call("mxg_cache_stats", C.byref(hits), C.byref(misses), C.byref(nbytes), C.byref(entries))
    assert hits.value >= 4 and misses.value >= 2 and 0 < nbytes.value <= 256 << 20 and entries.value >= 1
    # a budget too small for the matrix: nothing is kept, results unchanged
    lib.call("mxg_cache_clear")
    lib.set_option("cache_mb", 1)
    small = rx.tcrossprod_csr_dense_numeric(p, j, x, Y)
    lib.call("mxg_cache_stats", C.byref(hits), C.byref(misses), C.byref(nbytes), C.byref(entries))
    assert np.array_equal(small, again2) and nbytes.value <= 1 << 20
    # an invalid matrix never enters the cache
    jb = j.copy()
    jb[5] = K
    lib.set_option("cache_mb", 256)
    with pytest.raises(lib.MxgError):
        rx.tcrossprod_csr_dense_numeric(p, jb, x, Y)
    with pytest.raises(lib.MxgError):
        rx.tcrossprod_csr_dense_numeric(p, jb, x, Y)


def test_set_devices_argument_checks(lib):
    n = lib.device_count()
    with pytest.raises(lib.MxgError):
        lib.call("mxg_set_devices", 0)
    with pytest.raises(lib.MxgError):
        lib.call("mxg_set_devices", n + 1)
    lib.call("mxg_set_devices", 1)
    got = C.c_int(0)
    lib.call("mxg_get_devices", C.byref(got))
    assert got.value == 1


@pytest.mark.parametrize("share", [1, 0])
def test_one_call_over_several_devices_is_bit_identical(rx, lib, port, share):
    """mxg_set_devices(n): n host threads, n streamed pipelines on nnz-balanced row blocks, each device uploading its
    block (and one slice of the dense operand) over its own link.  Bit-identical to one device."""
    ndev = lib.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    G = min(ndev, 4)
    m, K = 60000, 40000
    p, j, x = powerlaw_csr(m, K, 40, seed=31, cap=5000)
    rng = np.random.default_rng(31)
    lib.set_option("multi_min_nnz", 1000)
    lib.set_option("multi_dense_share", share)
    lib.set_option("multi_pageable", 1)  # numpy arrays are pageable: by default such calls stay on one device
    cases = []
    for dt, n in ((np.float64, 32), (np.float32, 64), (np.float64, 5)):
        Y = np.asfortranarray(rng.standard_normal((n, K)).astype(dt))
        cases.append((dt, Y))
    y = rng.standard_normal(K)
    lib.call("mxg_set_devices", 1)
    ref_cm = [getattr(rx, "tcrossprod_csr_dense_" + _sfx(dt))(p, j, x, Y) for dt, Y in cases]
    ref_rm = [getattr(rx, "tcrossprod_dense_csr_" + _sfx(dt))(Y, p, j, x, 0, K) for dt, Y in cases]
    ref_v = rx.matmul_csr_dvec_numeric(p, j, x, y)
    lib.call("mxg_set_devices", G)
    got = C.c_int(0)
    lib.call("mxg_get_devices", C.byref(got))
    assert got.value == G
    for (dt, Y), want_cm, want_rm in zip(cases, ref_cm, ref_rm):
        assert np.array_equal(getattr(rx, "tcrossprod_csr_dense_" + _sfx(dt))(p, j, x, Y), want_cm)
        up, down = _bytes(lib)
        assert down == want_cm.nbytes + 4 * G  # + every device's 4-byte validation flag
        if share and Y.nbytes >= 4 << 20:
            assert up < p.nbytes + j.nbytes + x.nbytes + Y.nbytes + (1 << 20)  # the dense operand crossed PCIe once
        assert np.array_equal(getattr(rx, "tcrossprod_dense_csr_" + _sfx(dt))(Y, p, j, x, 0, K), want_rm)
        assert rel_err(want_cm, getattr(port, "tcrossprod_csr_dense_" + _sfx(dt))(p, j, x, Y)) <= _tol(dt)
    assert np.array_equal(rx.matmul_csr_dvec_numeric(p, j, x, y), ref_v)
    # an invalid column id in ONE block fails the whole call with the library's message
    jb = j.copy()
    jb[-3] = K + 7
    with pytest.raises(lib.MxgError, match="column index"):
        rx.tcrossprod_csr_dense_numeric(p, jb, x, cases[0][1])
    # and the next call works
    assert np.array_equal(rx.tcrossprod_csr_dense_numeric(p, j, x, cases[0][1]), ref_cm[0])


def test_results_are_allocated_from_the_page_locked_pool(rx, lib, port):
    """What the glue does with Rf_allocVector3 (rglue/mxgpu_result_alloc.h), mirrored by rcpp_exports._result: a large
    result is a block of the library's page-locked pool — written by the device directly, recycled when collected."""
    import gc

    def stats():
        live, free, blocks = C.c_size_t(), C.c_size_t(), C.c_int()
        lib.call("mxg_host_pool_stats", C.byref(live), C.byref(free), C.byref(blocks))
        return live.value, free.value, blocks.value

    m, K, n = 30000, 4000, 16  # 3.84 MB result
    p, j, x = powerlaw_csr(m, K, 12, seed=71, cap=2000)
    Y = np.asfortranarray(np.random.default_rng(71).standard_normal((n, K)))
    gc.collect()
    live0, _, _ = stats()
    res = rx.tcrossprod_csr_dense_numeric(p, j, x, Y)
    assert res.flags.f_contiguous and res.flags.writeable and res.shape == (m, n)
    live1, _, _ = stats()
    assert live1 - live0 >= res.nbytes  # its memory is a live pool block ...
    up, down = _bytes(lib)
    assert down == res.nbytes + 4
    want = port.tcrossprod_csr_dense_numeric(p, j, x, Y)
    assert rel_err(res, want) <= FP64_TOL
    heap = rx.tcrossprod_csr_dense_numeric(p, j, x, Y, out=np.empty((m, n), order="F"))  # a result on the caller's heap
    assert np.array_equal(res, heap) and stats()[0] == live1
    addr = res.ctypes.data
    view = res[:, 3]  # a view keeps the block alive
    del res
    gc.collect()
    assert stats()[0] == live1 and np.array_equal(view, heap[:, 3])
    del view
    gc.collect()
    live2, free2, blocks2 = stats()
    assert live2 == live0 and free2 >= heap.nbytes  # ... handed back when the array is collected ...
    again = rx.tcrossprod_csr_dense_numeric(p, j, x, Y)
    assert np.array_equal(again, heap)
    assert stats() == (live1, free2 - (live1 - live0), blocks2)  # ... and a free block is reused by the next result of that size
    del addr
    small = rx.matmul_csr_dvec_numeric(p, j, x, np.ones(K))  # 240 KB: stays on the ordinary heap
    assert stats()[0] == live1
    del again, small
    gc.collect()
    # a pool that may hold nothing: results fall back to ordinary arrays, same bits
    lib.call("mxg_trim")
    lib.set_option("host_result_pool_mb", 0)
    res = rx.tcrossprod_csr_dense_numeric(p, j, x, Y)
    assert np.array_equal(res, heap) and stats()[0] == 0
    # the raw entry points: alignment, unknown pointers, recycling with at most 25 % waste
    lib.set_option("host_result_pool_mb", 64)
    a, b = C.c_void_p(), C.c_void_p()
    lib.call("mxg_host_alloc", 5 << 20, C.byref(a))
    assert a.value % 4096 == 0
    with pytest.raises(lib.MxgError):
        lib.call("mxg_host_free", C.c_void_p(a.value + 64))
    lib.call("mxg_host_free", a)
    lib.call("mxg_host_alloc", 1 << 20, C.byref(b))  # far smaller: a block of its own, not the 6 MiB one
    assert b.value != a.value
    lib.call("mxg_host_free", b)
    with pytest.raises(lib.MxgError):
        lib.call("mxg_host_alloc", 65 << 20, C.byref(a))  # beyond the cap: the caller falls back
    lib.call("mxg_trim")
    assert stats()[1] == 0


@pytest.mark.parametrize("dt,n", [(np.float32, 64), (np.float64, 32), (np.float64, 64), (np.float32, 48)])
def test_warm_product_in_two_column_halves(rx, lib, port, dt, n):
    """host_colsplit: a result of >= 32 MiB leaves as two column halves so that the link's directions overlap
    (pipeline.cu: handle_spmm_host_split).  Default mode: taken only where every element is summed in the same order
    (256-byte rows: fp32 n = 64, fp64 n = 32) and then bit-identical to the unsplit product; mode 2 splits fp64 n = 64
    as well (other team geometry: equal within the tolerance only); 96-byte halves (fp32 n = 48) are never split.
    Every mix of pageable / page-locked operand and result, with long rows (pieces) present."""
    import torch
    lib.set_option("piece", 64)
    m, K = 180_000, 4000
    p, j, x = powerlaw_csr(m, K, 12, seed=21, cap=3000)
    sfx = _sfx(dt)
    call = getattr(rx, "gpu_csr_dense_tcrossprod_" + sfx)
    h = rx.as_gpu_csr(p, j, x, K, keep_float64=True, keep_float32=True)
    try:
        X = np.asfortranarray(np.random.default_rng(n).standard_normal((n, K)).astype(dt))  # (n x K): X %*% t(A)
        assert n * m * X.itemsize >= 32 << 20
        lib.set_option("host_colsplit", 0)
        whole = call(X, h, out=np.empty((n, m), dtype=dt, order="F"))
        want = getattr(port, "tcrossprod_dense_csr_" + sfx)(X, p, j, x, None, K)
        assert rel_err(whole, want) <= _tol(dt)
        t_dt = torch.float32 if dt == np.float32 else torch.float64
        Xp = torch.empty(n * K, dtype=t_dt).pin_memory().numpy().reshape((n, K), order="F")
        Xp[...] = X
        for mode in (1, 2):
            lib.set_option("host_colsplit", mode)
            for pin_x, pin_out in ((False, False), (True, True), (False, True), (True, False)):
                out = (torch.full((n * m,), float("nan"), dtype=t_dt).pin_memory().numpy().reshape((n, m), order="F") if pin_out
                       else np.full((n, m), np.nan, dtype=dt, order="F"))
                got = call(Xp if pin_x else X, h, out=out)
                assert _bytes(lib) == (X.nbytes, n * m * X.itemsize)
                if mode == 1 or n * X.itemsize != 512:
                    assert np.array_equal(got, whole), (mode, pin_x, pin_out)
                else:  # fp64 n = 64 in mode 2: two 32-column products on 4 sub-teams instead of one on 2
                    assert not np.isnan(got).any() and rel_err(got, want) <= _tol(dt)
                    assert np.array_equal(got[:, np.diff(p) <= 1], whole[:, np.diff(p) <= 1])  # nothing to re-order there
        # fewer output slots than chunks and more: the ring of page-locked slots wraps in both
        lib.set_option("host_colsplit", 2)
        for slots in (3, 8):
            lib.set_option("pipe_slots", slots)
            got = call(X, h, out=np.full((n, m), np.nan, dtype=dt, order="F"))
            assert rel_err(got, want) <= _tol(dt) and not np.isnan(got).any()
    finally:
        rx.gpu_csr_free(h)


@pytest.mark.parametrize("register", [1, 0])
def test_page_locked_blocks_of_both_kinds(rx, lib, port, register):
    """host_pin_register: new page-locked blocks (result pool, staging arena) are huge-page backed anonymous memory
    registered with the driver (1) or cudaHostAlloc blocks (0).  Either kind is page-locked to the library (DMA in place,
    no bounce), recycled by the pool, and released by mxg_trim; the products are the same bits."""
    import gc
    gc.collect()
    lib.call("mxg_trim")  # drop the arena and the free pool blocks of earlier tests: the next ones are of this kind
    lib.set_option("host_pin_register", register)
    try:
        m, K, n = 40000, 3000, 32
        p, j, x = powerlaw_csr(m, K, 10, seed=91, cap=1500)
        Y = np.asfortranarray(np.random.default_rng(91).standard_normal((n, K)))
        a = C.c_void_p()
        lib.call("mxg_host_alloc", 9 << 20, C.byref(a))
        buf = np.frombuffer((C.c_char * (9 << 20)).from_address(a.value), dtype=np.uint8)
        buf[:] = 7  # every page is there and writable
        assert a.value % (2 << 20) == 0 or register == 0
        import torch
        t = torch.empty(9 << 20, dtype=torch.uint8, device="cuda")
        try:
            rt = C.CDLL("libcudart.so.12")
        except OSError:
            rt = None
        if rt is not None:
            attr = (C.c_int * 16)()
            assert rt.cudaPointerGetAttributes(attr, a) == 0 and attr[0] == 1  # cudaMemoryTypeHost: page-locked to the driver
        lib.call("mxg_host_free", a)
        del t, buf
        res = rx.tcrossprod_csr_dense_numeric(p, j, x, Y)  # 10 MB result: the freed 10 MiB block, DMA'd into directly
        up, down = _bytes(lib)
        assert down == res.nbytes + 4
        assert rel_err(res, port.tcrossprod_csr_dense_numeric(p, j, x, Y)) <= FP64_TOL
        heap = rx.tcrossprod_csr_dense_numeric(p, j, x, Y, out=np.empty((m, n), order="F"))  # through the arena's slots
        assert np.array_equal(res, heap)
        del res
        gc.collect()
        lib.call("mxg_trim")
        live, free, blocks = C.c_size_t(), C.c_size_t(), C.c_int()
        lib.call("mxg_host_pool_stats", C.byref(live), C.byref(free), C.byref(blocks))
        assert free.value == 0
    finally:
        lib.set_option("host_pin_register", 1)
        lib.call("mxg_trim")
