"""CPU check of the arithmetic shortcut the CUDA kernels use: division by the
fitness denominators through their correctly rounded reciprocals plus one fused
correction step must equal IEEE division bit for bit ("exactness of f64")."""
import ctypes as C


def test_reciprocal_division_equals_ieee_division(oracle):
    fn = oracle.lib.oracle_check_fastdiv
    fn.restype = C.c_int64
    fn.argtypes = [C.c_int64, C.c_uint64]
    assert fn(20_000_000, 12345) == 0
