"""Host-side type table of the Python mirror (src/core/types.zig:36-104): 20 SUPPORTED_TYPES, Complex(T) storage."""
import numpy as np
import pytest

from wekua_b200 import core


def test_type_indices():
    assert len(core.SUPPORTED_TYPES) == 20
    for i, t in enumerate(core.REAL_TYPES):
        assert core.get_type_index(t) == i
        assert core.get_type_index(core.Complex(t)) == 10 + i
        assert core.base_type(core.Complex(t)) == np.dtype(t)
        assert core.storage_dtype(core.Complex(t)).itemsize == 2 * np.dtype(t).itemsize
    assert core.get_type_index(np.complex64) == 18 and core.get_type_index(np.complex128) == 19
    assert core.is_complex(np.complex64) and not core.is_complex(np.float32)
    with pytest.raises(Exception):
        core.get_type_index(np.float16)


def test_as_elements():
    dt = core.Complex(np.int16)
    a = core.as_elements([1, 2, 3], dt)
    assert a.dtype == dt and list(a["re"]) == [1, 2, 3] and not a["im"].any()
    b = core.as_elements(np.array([1 + 2j, 3 - 4j]), core.Complex(np.float32))
    assert list(b["re"]) == [1, 3] and list(b["im"]) == [2, -4]
    s = core.as_elements((7, 2), dt).reshape(1)
    assert s["re"][0] == 7 and s["im"][0] == 2
    assert core.as_elements(np.arange(4), np.float64).dtype == np.float64
    c = core.as_elements(b, np.complex64)
    assert c.dtype == core.Complex(np.float32)
