"""Known answers of the numpy Generator streams the path depends on,
captured with numpy 2.3.5.  numpy does not promise stream stability across versions: if one of
these fails, every seeded map/task changes and the golden fixtures must be regenerated."""
import numpy as np


def test_binomial_stream():
    exp = [[1, 0, 1, 0], [0, 1, 1, 1], [0, 0, 0, 1], [0, 1, 0, 0]]
    assert np.random.default_rng(42).binomial(1, 0.3, (4, 4)).tolist() == exp


def test_shuffle_stream():
    order = [(x, y) for x in range(3) for y in range(3)]
    np.random.default_rng(42).shuffle(order)
    assert order == [(1, 0), (0, 0), (2, 1), (0, 2), (1, 1), (2, 0), (0, 1), (1, 2), (2, 2)]


def test_integers_streams():
    assert np.random.default_rng(42).integers(5, size=6).tolist() == [0, 3, 3, 2, 2, 4]
    assert np.random.default_rng(42).integers(np.iinfo(np.int32).max, size=3).tolist() == [191664963, 1662057957,
                                                                                          1405681631]


def test_choice_stream():
    assert tuple(int(v) for v in tuple(*np.random.default_rng(7).choice([(1, 2), (3, 4), (5, 6)], 1))) == (5, 6)


def test_pcg64_seed_state():
    st = np.random.default_rng(42).bit_generator.state
    assert st['state']['state'] == 274674114334540486603088602300644985544
    assert st['state']['inc'] == 332724090758049132448979897138935081983
