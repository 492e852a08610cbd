"""GPU tests of the host staging engine of the streamed level-1 calls (csrc/hoststage.cu + csrc/pipeline.cu):
pageable caller memory bounced through the page-locked ring by the host threads, float32 values narrowed on the
host.  Whatever route the bytes take — pageable or page-locked operands, host or device narrowing, staging on or
off, one thread or many, few ring slots and many chunks — the result must be the SAME BITS, and within the
north_star tolerance of the CPU oracle (src/matmul.cpp:118-185, 381-483)."""
import ctypes as C
import itertools

import numpy as np
import pytest

from helpers import FP32_TOL, FP64_TOL, powerlaw_csr, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rx():
    from matrixextra_b200 import rcpp_exports
    return rcpp_exports


@pytest.fixture()
def options():
    from matrixextra_b200 import _lib
    names = ("pipe_chunk_nnz", "piece", "host_narrow", "host_stage", "host_threads", "pipe_slots", "host_arena_max_mb", "host_pack")
    old = {k: _lib.get_option(k) for k in names}
    yield _lib
    for k, v in old.items():
        _lib.set_option(k, v)


def _pinned(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t.numpy()


def _pinned_f(shape, dtype):
    """Page-locked Fortran-ordered matrix."""
    import torch
    flat = torch.empty(int(np.prod(shape)), dtype={np.float32: torch.float32, np.float64: torch.float64}[dtype]).pin_memory()
    return flat.numpy().reshape(shape, order="F")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_every_staging_route_gives_the_same_bits(rx, port, options, dtype):
    sfx = "float32" if dtype == np.float32 else "numeric"
    tol = FP32_TOL if dtype == np.float32 else FP64_TOL
    m, K, n = 4000, 1500, 24
    p, j, x = powerlaw_csr(m, K, 15, seed=77, cap=1400)
    options.set_option("piece", 64)
    options.set_option("pipe_chunk_nnz", 1500)  # ~40 chunks: the ring wraps many times
    rng = np.random.default_rng(3)
    X = np.asfortranarray(rng.standard_normal((n, K)).astype(dtype))
    want_rm = getattr(port, "tcrossprod_dense_csr_" + sfx)(X, p, j, x, 1, K)
    want_cm = getattr(port, "tcrossprod_csr_dense_" + sfx)(p, j, x, X, 1)
    pp, pj, px = _pinned(p), _pinned(j), _pinned(x)
    pX = _pinned_f((n, K), dtype)
    pX[...] = X
    ref_rm = ref_cm = None
    routes = itertools.product((0, 1), (0, 1), (1, 5), (False, True), (3, 8), (0, 2))
    for narrow, stage, threads, pinned, slots, pack in routes:
        options.set_option("host_pack", pack)  # 2: column ids packed on the host whatever the size (K = 1500: 2 bytes each)
        options.set_option("host_narrow", narrow)
        options.set_option("host_stage", stage)
        options.set_option("pipe_slots", slots)
        a = (pp, pj, px, pX) if pinned else (p, j, x, X)
        out_rm = _pinned_f((n, m), dtype) if pinned else None
        out_cm = _pinned_f((m, n), dtype) if pinned else None
        got_rm = getattr(rx, "tcrossprod_dense_csr_" + sfx)(a[3], a[0], a[1], a[2], threads, K, out=out_rm)
        got_cm = getattr(rx, "tcrossprod_csr_dense_" + sfx)(a[0], a[1], a[2], a[3], threads, out=out_cm)
        if ref_rm is None:
            ref_rm, ref_cm = got_rm.copy(), got_cm.copy()
            assert rel_err(ref_rm, want_rm) <= tol and rel_err(ref_cm, want_cm) <= tol
        route = dict(narrow=narrow, stage=stage, threads=threads, pinned=pinned, slots=slots, pack=pack)
        assert np.array_equal(got_rm, ref_rm), route
        assert np.array_equal(got_cm, ref_cm), route


def test_mixed_pinned_and_pageable_operands(rx, port, options):
    """Each operand is classified on its own: pageable indices with page-locked values, pageable result, ..."""
    m, K, n = 2500, 800, 16
    p, j, x = powerlaw_csr(m, K, 10, seed=5, cap=700)
    options.set_option("pipe_chunk_nnz", 2000)
    rng = np.random.default_rng(8)
    X = np.asfortranarray(rng.standard_normal((n, K)))
    want = port.tcrossprod_dense_csr_numeric(X, p, j, x, 1, K)
    ref = None
    for pin_j, pin_x, pin_X, pin_out in itertools.product((False, True), repeat=4):
        out = _pinned_f((n, m), np.float64) if pin_out else None
        XX = X
        if pin_X:
            XX = _pinned_f((n, K), np.float64)
            XX[...] = X
        got = rx.tcrossprod_dense_csr_numeric(XX, p, _pinned(j) if pin_j else j, _pinned(x) if pin_x else x, 1, K, out=out)
        if ref is None:
            ref = got.copy()
            assert rel_err(ref, want) <= FP64_TOL
        assert np.array_equal(got, ref), (pin_j, pin_x, pin_X, pin_out)


def test_staged_wide_operands_and_strided_result(rx, port, options):
    """A dense operand larger than one ring slot (cut into blocks of lines), a column-major operand whose lines
    are longer than the copy grain, and a caller result with ldc > n."""
    import ctypes as C
    from matrixextra_b200 import _lib
    from matrixextra_b200._lib import MXG_F64, MXG_ROWS_CONTIGUOUS, MXG_COLS_CONTIGUOUS
    m, K, n = 3000, 300_000, 8  # B: 300000 x 8 doubles = 19.2 MB > the 16 MiB slot
    rng = np.random.default_rng(11)
    lens = rng.integers(0, 12, size=m)
    p = np.zeros(m + 1, dtype=np.int32)
    np.cumsum(lens, out=p[1:])
    j = np.concatenate([np.sort(rng.choice(K, size=l, replace=False)) for l in lens]).astype(np.int32)
    x = rng.uniform(-1, 1, size=p[-1])
    X = np.asfortranarray(rng.standard_normal((n, K)))  # n x K column-major == K rows of n contiguous
    want = port.tcrossprod_dense_csr_numeric(X, p, j, x, 1, K)  # (n x m) column-major
    vp = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    for b_layout in (MXG_ROWS_CONTIGUOUS, MXG_COLS_CONTIGUOUS):
        Bsrc = X if b_layout == MXG_ROWS_CONTIGUOUS else np.ascontiguousarray(X)  # lines of K doubles (2.4 MB)
        ldb = n if b_layout == MXG_ROWS_CONTIGUOUS else K
        ldc = n + 3
        out = np.full((m, ldc), -7.0)
        _lib.call("mxg_spmm_csr_dense", MXG_F64, MXG_ROWS_CONTIGUOUS, b_layout, m, K, n, vp(p), vp(j), vp(x), vp(Bsrc), ldb,
                  vp(out), ldc)
        assert np.all(out[:, n:] == -7.0)  # the padding of the caller's rows is not touched
        assert rel_err(out[:, :n].T, want) <= FP64_TOL


@pytest.mark.parametrize("ytype", ["numeric", "integer", "logical", "float32"])
def test_spmv_staging_routes(rx, port, options, ytype):
    m, K = 6000, 2000
    p, j, x = powerlaw_csr(m, K, 12, seed=9, cap=1900)
    options.set_option("pipe_chunk_nnz", 3000)
    options.set_option("piece", 128)
    rng = np.random.default_rng(2)
    if ytype == "numeric":
        y = rng.standard_normal(K)
    elif ytype == "float32":
        y = rng.standard_normal(K).astype(np.float32)
    else:
        y = rng.integers(-3, 4, size=K).astype(np.int32)
        y[::97] = np.iinfo(np.int32).min  # NA
    fn = getattr(rx, "matmul_csr_dvec_" + ytype)
    want = getattr(port, "matmul_csr_dvec_" + ytype)(p, j, x, y, 1)
    ref = None
    for stage, threads, pinned in itertools.product((0, 1), (1, 4), (False, True)):
        options.set_option("host_stage", stage)
        a = (_pinned(p), _pinned(j), _pinned(x), _pinned(y)) if pinned else (p, j, x, y)
        got = np.asarray(fn(*a, threads))
        if ref is None:
            ref = got.copy()
            g64, w64 = np.asarray(ref, dtype=np.float64), np.asarray(want, dtype=np.float64)
            assert np.array_equal(np.isnan(g64), np.isnan(w64))
            ok = ~np.isnan(w64)
            assert rel_err(g64[ok], w64[ok]) <= (FP32_TOL if ytype == "float32" else FP64_TOL)
        assert np.array_equal(got.view(np.uint8), ref.view(np.uint8)), (stage, threads, pinned)


def test_bad_column_index_is_still_reported_through_the_staged_route(rx, options):
    from matrixextra_b200 import _lib
    p, j, x = powerlaw_csr(500, 200, 8, seed=1)
    j = j.copy()
    j[len(j) // 2] = 200  # == K: out of range
    options.set_option("pipe_chunk_nnz", 500)
    X = np.asfortranarray(np.ones((4, 200), dtype=np.float32))
    with pytest.raises(_lib.MxgError, match="outside"):
        rx.tcrossprod_dense_csr_float32(X, p, j, x, 1, 200)
    # and the library is usable afterwards
    j[len(j) // 2] = 0
    rx.tcrossprod_dense_csr_float32(X, p, j, x, 1, 200)


def _big_random_csr(m, K, per_row, seed):
    """Large enough (> 1 MiB per array) for the one-shot staged copies to engage."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(per_row // 2, per_row * 3 // 2, size=m)
    p = np.zeros(m + 1, dtype=np.int32)
    np.cumsum(lens, out=p[1:])
    j = np.empty(p[-1], dtype=np.int32)
    for r in range(m):
        j[p[r]:p[r + 1]] = np.sort(rng.choice(K, size=lens[r], replace=False))
    return p, j, rng.uniform(-1, 1, size=p[-1])


def test_one_shot_staged_copies_upload_crossprod_and_csr2csc(rx, port, options):
    """Handle upload, crossprod (device transpose + product) and CSR->CSC through the staged one-shot copies:
    identical bits with staging / host narrowing on and off, CSR->CSC equal to scipy's tocsc()."""
    import scipy.sparse as sp
    from matrixextra_b200._lib import MXG_F32, MXG_F64, MXG_KEEP_F32, MXG_KEEP_F64
    from matrixextra_b200.device import DeviceCSR
    m, K, n = 20_000, 3_000, 80  # ~400k entries: j 1.6 MB, x 3.2 MB, Y 6.4 MB (fp32) / 12.8 MB, result 1-2 MB
    p, j, x = _big_random_csr(m, K, 20, seed=31)
    assert j.nbytes > (1 << 20)
    rng = np.random.default_rng(31)
    Y = np.asfortranarray(rng.standard_normal((m, n)))
    q2, k2, y2 = port.csr2csc(m, K, p, j, x)
    want64 = port.matmul_dense_csc_numeric(np.asfortranarray(Y.T), q2, k2, y2).T
    want32 = port.matmul_dense_csc_float32(np.asfortranarray(Y.T.astype(np.float32)), q2, k2, y2).T
    ref = {}
    for stage, narrow in ((1, 1), (1, 0), (0, 0)):
        options.set_option("host_stage", stage)
        options.set_option("host_narrow", narrow)
        # level-2 upload: float32-only handles take the host-narrowing route
        for keep in (MXG_KEEP_F64, MXG_KEEP_F32, MXG_KEEP_F64 | MXG_KEEP_F32):
            A = DeviceCSR.upload(m, K, p, j, x, keep=keep)
            pp, jj, xx = A.to_host()
            A.free()
            assert np.array_equal(pp, p) and np.array_equal(jj, j)
            if keep & MXG_KEEP_F64:
                assert np.array_equal(xx, x)
            else:  # values come back widened from the float32 copy
                assert np.array_equal(xx, x.astype(np.float32).astype(np.float64))
        got64 = rx.crossprod_csr_dense(p, j, x, K, Y, MXG_F64)
        got32 = rx.crossprod_csr_dense(p, j, x, K, Y.astype(np.float32), MXG_F32)
        p2, i2, x2 = rx.csr_to_csc(m, K, p, j, x)
        if not ref:
            ref = dict(g64=got64.copy(), g32=got32.copy())
            assert rel_err(got64, want64) <= FP64_TOL and rel_err(got32, want32) <= FP32_TOL
            S = sp.csr_matrix((x, j, p), shape=(m, K)).tocsc()
            assert np.array_equal(p2, S.indptr) and np.array_equal(i2, S.indices) and np.array_equal(x2, S.data)
        assert np.array_equal(got64, ref["g64"]) and np.array_equal(got32, ref["g32"]), (stage, narrow)
        assert np.array_equal(p2, q2) and np.array_equal(i2, k2) and np.array_equal(x2, y2)


def test_without_page_locked_memory_the_driver_path_takes_over(rx, port, options):
    """host_arena_max_mb caps the page-locked arena; when a call would need more (or cudaHostAlloc fails) every copy
    falls back to the driver's own bounce and the values are narrowed on the device: same bits, no error."""
    from matrixextra_b200._lib import MXG_F32
    m, K, n = 20_000, 3_000, 48
    p, j, x = _big_random_csr(m, K, 20, seed=41)
    rng = np.random.default_rng(41)
    X = np.asfortranarray(rng.standard_normal((n, K)).astype(np.float32))
    y = rng.standard_normal(K)
    Y = np.asfortranarray(rng.standard_normal((m, 8)).astype(np.float32))
    ref_mm = rx.tcrossprod_dense_csr_float32(X, p, j, x, 4, K)
    ref_mv = rx.matmul_csr_dvec_numeric(p, j, x, y, 4)
    ref_cp = rx.crossprod_csr_dense(p, j, x, K, Y, MXG_F32)
    assert rel_err(ref_mm, port.tcrossprod_dense_csr_float32(X, p, j, x, 1, K)) <= FP32_TOL
    options.set_option("host_arena_max_mb", 1)  # every arena request (>= 3 slots of >= 1 MiB) is refused
    assert np.array_equal(rx.tcrossprod_dense_csr_float32(X, p, j, x, 4, K), ref_mm)
    assert np.array_equal(rx.matmul_csr_dvec_numeric(p, j, x, y, 4), ref_mv)
    assert np.array_equal(rx.crossprod_csr_dense(p, j, x, K, Y, MXG_F32), ref_cp)


def test_one_shot_staged_copies_wrap_their_ring(rx, options):
    """Arrays of many 16 MiB blocks: the 4-slot ring of the one-shot staged copies wraps in both directions
    (upload of 80 MB indices / 160 MB values, host-narrowed upload of 20 M values, download of the CSC arrays)."""
    import scipy.sparse as sp
    from matrixextra_b200._lib import MXG_KEEP_F32, MXG_KEEP_F64
    from matrixextra_b200.device import DeviceCSR
    m, K = 400_000, 50_000
    A = DeviceCSR.synth(m, K, 20_000_000, 1, 0, seed=99, keep=MXG_KEEP_F64)
    p, j, x = A.to_host()  # pageable numpy arrays
    A.free()
    assert x.nbytes > 4 * (16 << 20) and j.nbytes > 4 * (16 << 20)
    for keep in (MXG_KEEP_F64, MXG_KEEP_F32):
        H = DeviceCSR.upload(m, K, p, j, x, keep=keep)
        pp, jj, xx = H.to_host()
        H.free()
        assert np.array_equal(pp, p) and np.array_equal(jj, j)
        assert np.array_equal(xx, x if keep == MXG_KEEP_F64 else x.astype(np.float32).astype(np.float64))
    p2, i2, x2 = rx.csr_to_csc(m, K, p, j, x)
    S = sp.csr_matrix((x, j, p), shape=(m, K)).tocsc()
    assert np.array_equal(p2, S.indptr) and np.array_equal(i2, S.indices) and np.array_equal(x2, S.data)


@pytest.mark.parametrize("K", [40_000, 70_000, 1_000_000, 1_100_000, 16_777_216, 16_777_217])
def test_packed_column_ids_every_width(rx, port, options, K):
    """Column ids cross PCIe as 2 / 2.5 / 3 bytes per entry (host_pack, csrc/hoststage.cu + k_unpack_indices) or as
    int32 above 2^24 columns: same bits as the unpacked route for SpMM (float32, host-narrowed values sharing the
    slot) and SpMV (float64), odd chunk starts, ids at both ends of the range."""
    m, n = 3000, 4
    rng = np.random.default_rng(K % 977)
    rows = [np.unique(rng.integers(0, K, size=l)) for l in rng.integers(0, 9, size=m)]  # sorted, unique
    rows[0] = np.union1d(rows[0], [0])
    rows[-1] = np.union1d(rows[-1], [K - 1])
    p = np.zeros(m + 1, dtype=np.int32)
    np.cumsum([len(r) for r in rows], out=p[1:])
    j = np.concatenate(rows).astype(np.int32)
    x = rng.uniform(-1, 1, size=p[-1])
    v = rng.standard_normal(K)
    X = np.asfortranarray(rng.standard_normal((n, 1)).astype(np.float32) * np.ones((1, K), dtype=np.float32))
    X[:, ::7] *= -0.5
    options.set_option("pipe_chunk_nnz", 777)
    got, moved = {}, {}
    for pack in (0, 2, 3, 1):  # off, every chunk, a fixed mix of packed and raw chunks, decided from the upload stream
        options.set_option("host_pack", pack)
        got[pack] = (rx.tcrossprod_dense_csr_float32(X, p, j, x, 4, K), rx.matmul_csr_dvec_numeric(p, _pinned(j), _pinned(x), v, 4))
        h2d, d2h = C.c_size_t(0), C.c_size_t(0)
        options.call("mxg_last_call_bytes", C.byref(h2d), C.byref(d2h))  # of the SpMV call
        moved[pack] = (h2d.value, d2h.value)
    for pack in (2, 3, 1):
        assert np.array_equal(got[0][0], got[pack][0]) and np.array_equal(got[0][1], got[pack][1]), pack
    nnz = int(p[-1])
    raw = 4 * (m + 1) + 12 * nnz + 8 * K
    assert raw <= moved[0][0] <= raw + 4096 and moved[0][1] == moved[2][1] == 8 * m + 4  # + the validation flag
    if K <= 1 << 24:  # 2, 2.5 or 3 bytes per id, every chunk padded to 2 x 16 bytes
        per_id = 2 if K <= 1 << 16 else (2.5 if K <= 1 << 20 else 3)
        assert moved[0][0] - (4 - per_id) * nnz <= moved[2][0] <= moved[0][0] - (4 - per_id) * nnz + 33 * (nnz // 777 + m // 777 + 4)
        assert moved[2][0] < moved[3][0] < moved[0][0]
    else:
        assert moved[2][0] == moved[3][0] == moved[0][0]
    assert moved[1][0] == moved[0][0]  # the automatic mode leaves a call of this size alone
    assert rel_err(got[2][0], port.tcrossprod_dense_csr_float32(X, p, j, x, 1, K)) <= FP32_TOL
    assert rel_err(got[2][1], port.matmul_csr_dvec_numeric(p, j, x, v, 1)) <= FP64_TOL


@pytest.mark.parametrize("bad", [-1, "K", "K+70000"])
def test_packed_column_ids_out_of_range_is_an_error(rx, options, bad):
    """The packed route keeps the streamed call's validation: a negative id is caught by the host threads while
    packing, one at or beyond K by the device when the chunk is unpacked; both fail like the unpacked route."""
    from matrixextra_b200._lib import MxgError
    m, K, n = 500, 70_000, 4
    p, j, x = powerlaw_csr(m, K, 6, seed=3, cap=60)
    j = j.copy()
    j[len(j) // 2] = {-1: -1, "K": K, "K+70000": K + 70_000}[bad]
    X = np.asfortranarray(np.ones((n, K), dtype=np.float32))
    options.set_option("pipe_chunk_nnz", 500)
    for pack in (0, 2):
        options.set_option("host_pack", pack)
        with pytest.raises(MxgError, match="column index outside"):
            rx.tcrossprod_dense_csr_float32(X, p, j, x, 2, K)


def test_packed_column_ids_automatic_mode_on_a_long_call(rx, options):
    """A call long enough for the automatic mode (>= 2^20 entries, >= 8 host threads): chunks are packed or not
    depending on how far the upload stream lags behind the host threads — whatever the mix, same bits."""
    from matrixextra_b200._lib import MXG_KEEP_F64
    from matrixextra_b200.device import DeviceCSR
    m, K, n = 200_000, 300_000, 16
    A = DeviceCSR.synth(m, K, 6_000_000, 1, 1, seed=123, keep=MXG_KEEP_F64)
    p, j, x = A.to_host()
    A.free()
    rng = np.random.default_rng(123)
    v = rng.standard_normal(K)
    X = np.asfortranarray(rng.standard_normal((n, K)).astype(np.float32))
    pj, px = _pinned(j), _pinned(x)
    options.set_option("pipe_chunk_nnz", 250_000)  # ~24 chunks
    got, moved = {}, {}
    for pack in (0, 1, 1):
        options.set_option("host_pack", pack)
        mv = rx.matmul_csr_dvec_numeric(p, pj, px, v, 16)
        h2d = C.c_size_t(0)
        options.call("mxg_last_call_bytes", C.byref(h2d), None)
        mm = rx.tcrossprod_dense_csr_float32(X, p, pj, px, 16, K)
        if pack in got:
            assert np.array_equal(mv, got[pack][0]) and np.array_equal(mm, got[pack][1])
        got[pack], moved[pack] = (mv, mm), h2d.value
    assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1])
    import os
    if (os.cpu_count() or 1) >= 8:
        assert moved[1] < moved[0]  # at least the first two chunks (queued behind the vector) go packed
