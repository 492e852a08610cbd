"""CPU tests of the host-side mirror of R/matmul.R: dispatch table, dimension checks and their
messages, class handling — everything that runs before the first CUDA call."""
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
