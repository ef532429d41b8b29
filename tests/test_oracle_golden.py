"""The C oracle against the committed golden vectors (independent numpy restatement, see
tests/golden/make_golden.py).  CPU only."""
import pytest

import golden_util as G


@pytest.mark.parametrize("name", G.NAMES)
def test_oracle_matches_golden(name, oracle_lattice_factory):
    g = G.load(name)
    be = G.replay(g, oracle_lattice_factory)
    G.check(name, be, g)
