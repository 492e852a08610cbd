"""CPU tests of the host-side mirror of R/matmul.R: dispatch table, dimension checks and their
messages, class handling — everything that runs before the first CUDA call."""
import os
import numpy as np
import pytest

from helpers import rsparsematrix


def test_dimension_checks_fire_before_any_device_work():
    from matrixextra_b200 import crossprod, dgCMatrix, dgRMatrix, float32, matmul, tcrossprod
    A = dgRMatrix.from_scipy(rsparsematrix(10, 7, 0.3, 0))
    Ac = dgCMatrix.from_scipy(rsparsematrix(7, 5, 0.3, 0, "csc"))
    with pytest.raises(ValueError, match="Matrix dimensions do not match."):  # R/matmul.R:144-145
        matmul(A, np.zeros((8, 3), order="F"))
    with pytest.raises(ValueError, match="Matrix dimensions do not match."):
        tcrossprod(A, np.zeros((3, 8), order="F"))
    with pytest.raises(ValueError, match="Matrix dimensions do not match."):
        tcrossprod(np.zeros((3, 8), order="F"), A)
    with pytest.raises(ValueError, match="Matrix dimensions do not match."):
        matmul(np.zeros((3, 8), order="F"), Ac)
    with pytest.raises(ValueError, match="Matrix dimensions do not match."):
        crossprod(A, np.zeros((11, 2), order="F"))
    with pytest.raises(ValueError, match="Matrix dimensions do not match."):
        tcrossprod(A, float32(np.zeros((3, 8), dtype=np.float32)))
    with pytest.raises(ValueError, match="Matrix-vector dimensions do not match."):  # R/matmul.R:546-547
        matmul(A, np.zeros(8))
    with pytest.raises(TypeError):
        matmul(A, A)
    with pytest.raises(NotImplementedError):  # single-column outer product: outside the scoped path
        matmul(dgRMatrix.from_scipy(rsparsematrix(5, 1, 0.9, 1)), np.zeros(1))


def test_check_valid_matrix_messages():
    from matrixextra_b200.classes import check_valid_matrix, dgRMatrix
    ok = dgRMatrix([0, 1, 2], [0, 1], [1.0, 2.0], (2, 2))
    check_valid_matrix(ok)
    with pytest.raises(ValueError, match="'p' doesn't match with dimension"):
        check_valid_matrix(dgRMatrix([0, 1, 2], [0, 1], [1.0, 2.0], (3, 2)))
    with pytest.raises(ValueError, match="'p' has bad start/end"):
        check_valid_matrix(dgRMatrix([1, 1, 2], [0, 1], [1.0, 2.0], (2, 2)))
    with pytest.raises(ValueError, match="lengths of indices and values differ"):
        check_valid_matrix(dgRMatrix([0, 1, 2], [0, 1], [1.0], (2, 2)))


def test_shallow_transpose_relabels_without_copying():
    from matrixextra_b200 import dgCMatrix, dgRMatrix, t_shallow
    A = dgRMatrix.from_scipy(rsparsematrix(6, 4, 0.5, 2))
    At = t_shallow(A)
    assert isinstance(At, dgCMatrix) and At.Dim == (4, 6)
    assert At.p is A.p and At.i is A.j and At.x is A.x
    assert np.array_equal(At.to_scipy().toarray(), A.to_scipy().toarray().T)
    assert isinstance(t_shallow(At), dgRMatrix)


def test_float32_container_keeps_column_major_binary32():
    from matrixextra_b200 import float32
    f = float32(np.arange(6, dtype=np.float64).reshape(2, 3))
    assert f.Data.dtype == np.float32 and f.Data.flags.f_contiguous and not f.is_vector()
    assert float32(np.zeros(4)).is_vector()


# ---- host staging engine (csrc/hoststage.cu): worker threads, no GPU involved ---------------------------------
def _vp(a):
    import ctypes as C
    return C.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("threads", [1, 3, 16])
def test_host_narrow_is_the_reference_cast_bit_for_bit(threads):
    """(float)values[ix] of src/matmul.cpp:53-57 == numpy astype(float32): round-to-nearest-even, every class of
    double (normal, subnormal result, overflow to inf, NaN, signed zero), any alignment, any thread count."""
    from matrixextra_b200 import _lib
    _lib.set_option("host_threads", threads)
    try:
        rng = np.random.default_rng(threads)
        n = 1_000_003
        x = rng.standard_normal(n) * np.exp(rng.uniform(-110, 95, n))
        x[:12] = [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, 1e-46, 3.4028235677973366e38, 3.5e38, -1e-310,
                  1.0000000596046448, 1.0000001788139343]  # ties and just-above ties at float32 precision
        want = x.astype(np.float32)
        for off in (0, 1, 3):
            buf = np.empty(n + 4, dtype=np.float32)
            dst = buf[off:off + n]
            _lib.call("mxg_host_narrow", _vp(x), _vp(dst), n)
            assert np.array_equal(dst.view(np.uint32), want.view(np.uint32))
        _lib.call("mxg_host_narrow", None, None, 0)
    finally:
        _lib.set_option("host_threads", 0)


def test_host_copy_2d_pitched_lines():
    from matrixextra_b200 import _lib
    rng = np.random.default_rng(5)
    for width, height, sp, dp in [(256, 70_000, 256, 256), (256, 5_000, 320, 272), (3 << 20, 3, (3 << 20) + 64, 3 << 20),
                                  (1, 1, 1, 1), (0, 10, 8, 8), (100, 1, 50, 60)]:
        for streaming, off in ((0, 0), (1, 0), (1, 5)):  # cache-bypassing stores too, also to an unaligned destination
            src = rng.integers(0, 255, size=max(sp * height, width, 1), dtype=np.uint8)
            buf = np.full(max(dp * height, width, 1) + 8, 7, dtype=np.uint8)
            dst = buf[off:off + max(dp * height, width, 1)]
            _lib.call("mxg_host_copy_2d", _vp(dst), dp, _vp(src), sp, width, height, streaming)
            want = np.full_like(dst, 7)
            for l in range(height):
                want[l * dp:l * dp + width] = src[l * sp:l * sp + width]
            assert np.array_equal(dst, want)
            assert np.all(buf[:off] == 7) and np.all(buf[off + dst.size:] == 7)
    with pytest.raises(_lib.MxgError):
        _lib.call("mxg_host_copy_2d", _vp(dst), 4, _vp(src), 4, 8, 2, 0)


def _unpack_indices(buf, n, K):
    """numpy restatement of k_unpack_indices (csrc/layout.cu): the wire format documented in include/mxgpu.h."""
    lo_bytes = (2 * n + 15) & ~15
    lo = buf[:2 * n].view(np.uint16).astype(np.int32)
    hi = buf[lo_bytes:]
    if K <= 1 << 16:
        return lo
    if K <= 1 << 20:
        e = np.arange(n)
        return lo | (((hi[e >> 1] >> ((e & 1) * 4)) & 15).astype(np.int32) << 16)
    return lo | (hi[:n].astype(np.int32) << 16)


@pytest.mark.parametrize("threads", [1, 5, 16])
@pytest.mark.parametrize("K", [1, 300, 1 << 16, (1 << 16) + 1, 1_000_000, 1 << 20, (1 << 20) + 1, 10_000_000, 1 << 24])
def test_host_pack_indices_round_trip(threads, K):
    """Column ids packed for the wire (2 / 2.5 / 3 bytes per entry) come back bit for bit; sizes with tails of every
    length modulo 32 and several tasks; ids outside [0, K) are reported."""
    import ctypes as C
    from matrixextra_b200 import _lib
    _lib.set_option("host_threads", threads)
    try:
        rng = np.random.default_rng(K % 1000 + threads)
        for n in (0, 1, 2, 31, 32, 33, 63, 1000, 32768, 32769, 100_003):
            j = rng.integers(0, K, size=n, dtype=np.int32)
            if n > 2:
                j[0], j[-1], j[n // 2] = K - 1, K - 1, 0
            nbytes, ok = C.c_size_t(0), C.c_int(-1)
            _lib.call("mxg_host_pack_indices", None, n, K, None, C.byref(nbytes), None)
            hi = 0 if K <= 1 << 16 else ((n + 1) // 2 if K <= 1 << 20 else n)
            assert nbytes.value == ((2 * n + 15) & ~15) + ((hi + 15) & ~15)
            raw = np.full(nbytes.value + 16, 0xA5, dtype=np.uint8)
            off = (-raw.ctypes.data) % 16
            buf = raw[off:off + nbytes.value]
            _lib.call("mxg_host_pack_indices", _vp(j) if n else None, n, K, _vp(buf) if nbytes.value else _vp(raw), C.byref(nbytes), C.byref(ok))
            assert ok.value == 1
            assert np.array_equal(_unpack_indices(buf, n, K), j)
            for bad_at, bad in ((0, -1), (n - 1, K), (n // 2, np.iinfo(np.int32).min)):
                if n == 0:
                    break
                jb = j.copy()
                jb[bad_at] = bad
                _lib.call("mxg_host_pack_indices", _vp(jb), n, K, _vp(buf), C.byref(nbytes), C.byref(ok))
                assert ok.value == 0, (n, bad_at, bad)
    finally:
        _lib.set_option("host_threads", 0)


def test_host_pack_indices_wide_matrices_do_not_pack():
    import ctypes as C
    from matrixextra_b200 import _lib
    nbytes = C.c_size_t(7)
    _lib.call("mxg_host_pack_indices", None, 1000, (1 << 24) + 1, None, C.byref(nbytes), None)
    assert nbytes.value == 0


def _plan_by_the_rule(p, target_nnz, target_rows, taper):
    """Row-by-row restatement of the chunk rule of csrc/pipeline.cu build_plan: a chunk ends with the first row that
    brings it to the target; in the automatic mode targets shrink to a third of what is left (>= target / 8)."""
    m, nnz = len(p) - 1, int(p[-1])
    rows, start = [0], 0
    while start < m:
        first = int(p[start])
        tn, tr = target_nnz, target_rows
        if taper:
            tn = min(tn, max(target_nnz // 8, (nnz - first) // 3))
            tr = min(tr, max(target_rows // 8, (m - start) // 3))
        r = start
        while True:
            r += 1
            if r == m or int(p[r]) - first >= tn or r - start >= tr:
                break
        rows.append(r)
        start = r
    return rows


@pytest.mark.parametrize("forced", [0, 700, 5000])
def test_chunk_plan_follows_the_rule(forced):
    import ctypes as C
    from matrixextra_b200 import _lib
    from helpers import powerlaw_csr
    old = _lib.get_option("pipe_chunk_nnz"), _lib.get_option("piece")
    try:
        _lib.set_option("pipe_chunk_nnz", forced)
        _lib.set_option("piece", 256)
        rng = np.random.default_rng(forced)
        for m, mean in ((0, 0), (1, 0), (1, 5), (50_000, 3), (400_000, 9), (2_500_000, 1)):
            lens = np.minimum(np.floor(mean * 0.34 * (1 - rng.random(m)) ** (-1 / 1.5)), 40_000).astype(np.int64) if mean else np.zeros(m, np.int64)
            p = np.zeros(m + 1, dtype=np.int32)
            np.cumsum(lens, out=p[1:])
            nnz = int(p[-1])
            n = [C.c_int(0) for _ in range(4)]
            _lib.call("mxg_host_chunk_plan", m, _vp(p), 256, None, 0, *[C.byref(v) for v in n])
            rows = np.zeros(n[0].value + 1, dtype=np.int32)
            _lib.call("mxg_host_chunk_plan", m, _vp(p), 256, _vp(rows), len(rows), *[C.byref(v) for v in n])
            if forced:
                want = _plan_by_the_rule(p, forced, forced, taper=False)
            else:
                tn = min(max(nnz // 16, 1 << 20), 16 << 20)
                tr = min(max(m // 16, 1 << 16), max(1 << 16, (64 << 20) // 256))
                want = _plan_by_the_rule(p, tn, tr, taper=True)
            assert rows.tolist() == want, (m, mean)
            long_rows = lens[lens > 256]
            assert n[1].value == len(long_rows) and n[2].value == int(np.sum((long_rows + 255) // 256))
            assert n[3].value == (int(lens.max()) if m else 0)
        bad = np.array([0, 3, 2, 5], dtype=np.int32)
        with pytest.raises(_lib.MxgError, match="indptr"):
            _lib.call("mxg_host_chunk_plan", 3, _vp(bad), 8, None, 0, None, None, None, None)
    finally:
        _lib.set_option("pipe_chunk_nnz", old[0])
        _lib.set_option("piece", old[1])


def test_bench_traffic_table_is_tied_to_the_current_kernel_sources():
    """bench.py's roofline.traffic comes from committed ncu captures; every entry names the capture, the kernel and the
    SHA-256 of the source it was captured from.  A source that changed since must yield null, never a stale number — and
    the committed table is expected to be current (re-capture with tools/gpu_session.sh ncufull after touching a kernel)."""
    import bench
    for wl, e in bench.NCU_TRAFFIC.items():
        assert os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), e["capture"])), e["capture"]
        nbytes, why = bench.ncu_traffic(wl)
        assert nbytes == e["bytes"] and e["src_sha16"] in why, (wl, why)
    saved = dict(bench.NCU_TRAFFIC["cfg3"])
    try:
        bench.NCU_TRAFFIC["cfg3"]["src_sha16"] = "0" * 16
        nbytes, why = bench.ncu_traffic("cfg3")
        assert nbytes is None and "another version" in why
    finally:
        bench.NCU_TRAFFIC["cfg3"] = saved
