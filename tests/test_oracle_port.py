"""pins the CPU restatement (oracle/tpch_port.py) against the REAL reference's outputs (tests/golden, oracle/_ref)."""
import pytest

import tpch_port
from compare import compare
from sdqlpy_b200.tpch.gen import TPCH
from util import QUERIES, golden

_db = {}


def db(sf):
    if sf not in _db:
        _db[sf] = tpch_port.make_db(TPCH(sf))
    return _db[sf]


@pytest.mark.parametrize("q", QUERIES)
def test_port_matches_reference_sf001(q):
    assert compare(tpch_port.QUERIES[q](db(0.01)), golden(0.01)[q]) is None


@pytest.mark.parametrize("q", QUERIES)
def test_port_matches_reference_sf005(q):
    assert compare(tpch_port.QUERIES[q](db(0.05)), golden(0.05)[q]) is None
