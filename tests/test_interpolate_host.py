"""Host-only parts of sfsim_b200.interpolate against the reference's facts (t_interpolate.clj:22-72,107-117):
linear-space, compose-space, clip, mix and make-lookup-table of plain functions need no device."""
import math

import numpy as np
import pytest

from sfsim_b200 import interpolate as itp


@pytest.mark.parametrize("x,result", [(-2.0, 0.0), (4.0, 15.0), (0.0, 5.0)])
def test_linear_forward_1d(x, result):
    assert itp.linear_space([-2.0], [4.0], [16]).forward(x) == [result]


def test_linear_mappings():
    space = itp.linear_space([-2.0, -1.0], [4.0, 1.0], [16, 5])
    assert space.forward(4.0, 0.0) == [15.0, 2.0]
    assert space.backward(15, 2) == [4.0, 0.0]
    assert space.shape == (16, 5)
    one = itp.linear_space([-2.0], [4.0], [16])
    assert [one.backward(i) for i in (0.0, 15.0, 5.0)] == [[-2.0], [4.0], [0.0]]
    assert one.shape == (16,)


def test_make_lookup_table_of_plain_functions():
    table = itp.make_lookup_table(lambda x: x * x, itp.linear_space([-3.0], [2.0], [6]))
    np.testing.assert_array_equal(table, [9.0, 4.0, 1.0, 0.0, 1.0, 4.0])
    table2 = itp.make_lookup_table(lambda a, b: a * b, itp.linear_space([1.0, 3.0], [2.0, 5.0], [2, 3]))
    np.testing.assert_array_equal(table2, [[3.0, 4.0, 5.0], [6.0, 8.0, 10.0]])
    vec = itp.make_lookup_table(lambda x: (x, 2 * x, 3 * x), itp.linear_space([0.0], [1.0], [3]))
    np.testing.assert_array_equal(vec, [[0, 0, 0], [0.5, 1.0, 1.5], [1, 2, 3]])


@pytest.mark.parametrize("i,result", [(0.0, 0.0), (15.0, 15.0), (-2.0, 0.0), (16.0, 15.0)])
def test_clip(i, result):
    assert itp.clip(i, 16) == result


@pytest.mark.parametrize("s,result", [(0.0, -2.0), (1.0, 4.0), (0.5, 1.0), (0.25, -0.5)])
def test_mix(s, result):
    assert itp.mix(-2.0, 4.0, s) == result


def test_compose_space():
    radius_space = type("R", (), {"shape": None, "forward": staticmethod(lambda a, b: [math.hypot(a, b)]),
                                  "backward": staticmethod(lambda r: [r, 0.0])})
    combined = itp.compose_space(itp.linear_space([0.0], [1.0], [101]), radius_space)
    assert combined.shape == (101,)
    assert combined.forward(3.0, 4.0) == [500.0]
    assert combined.backward(500.0) == [5.0, 0.0]


def test_unknown_function_objects_are_rejected():
    with pytest.raises(TypeError):
        itp.make_lookup_table(42, itp.linear_space([0.0], [1.0], [3]))


def test_tables_of_the_wrong_shape_are_refused_before_the_library_reads_them():
    """The C entry points read prod(shape) * 3 floats from every table pointer: a mismatched table must never reach
    them (it would be an out-of-bounds host read)."""
    import numpy as np
    import pytest
    from sfsim_b200 import atmosphere, atmosphere_lut, interpolate as itp
    earth = atmosphere_lut.earth
    space4 = atmosphere.ray_scatter_space(earth, (4, 5, 3, 2))
    small = itp.interpolation_table(np.zeros((2, 2, 2, 2, 3), np.float32), atmosphere.ray_scatter_space(earth, (2, 2, 2, 2)))
    with pytest.raises(TypeError, match="shape"):
        itp.make_lookup_table(itp.RayScatter(earth, [atmosphere_lut.mie, atmosphere_lut.rayleigh], 10, small), space4)
    with pytest.raises(TypeError, match="different space"):
        itp.make_lookup_table(small, space4)
    other_planet = dict(earth, radius=1000.0, height=10.0)
    same_shape = itp.interpolation_table(np.zeros((4, 5, 3, 2, 3), np.float32),
                                         atmosphere.ray_scatter_space(other_planet, (4, 5, 3, 2)))
    with pytest.raises(TypeError, match="different space"):
        itp.make_lookup_table(same_shape, space4)
    source = atmosphere.FirstOrder(atmosphere.FirstOrder.COMPONENT, earth, [atmosphere_lut.mie, atmosphere_lut.rayleigh],
                                   atmosphere_lut.rayleigh, 10, (1, 1, 1))
    with pytest.raises(TypeError, match="scatter"):
        atmosphere.ray_scatter(earth, [atmosphere_lut.rayleigh, atmosphere_lut.mie], 10, source, (6378000.0, 0, 0),
                               (1, 0, 0), (0, 1, 0), True)
