"""CPU: algebraic checks of the packed marching-cubes table shipped in the product kernel."""
import numpy as np

from helpers import mc_tri_table

CONN = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]


def test_triangles_use_exactly_the_crossed_edges():
    tab = mc_tri_table()
    assert tab.shape == (256, 16)
    for c in range(256):
        crossed = {e for e, (a, b) in enumerate(CONN) if ((c >> a) & 1) != ((c >> b) & 1)}
        used = {int(x) for x in tab[c] if x >= 0}
        assert used == crossed, c      # this is why the kernel needs no edge-flag table
        n = int((tab[c] >= 0).sum())
        assert n % 3 == 0 and n <= 15
        assert (tab[c][n:] == -1).all()


def test_case_counts():
    tab = mc_tri_table()
    ntri = (tab >= 0).sum(1) // 3
    assert ntri[0] == 0 and ntri[255] == 0
    assert ntri.max() == 5 and ntri.sum() == 820
    # complementary cases triangulate the same edge set
    for c in range(256):
        assert set(tab[c][tab[c] >= 0]) == set(tab[255 - c][tab[255 - c] >= 0])
