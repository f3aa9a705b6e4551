"""the C-ABI library loads and exports every symbol include/sdqlb200.h declares (no compute without a GPU)."""
import ctypes
import json
import os
import re

import pytest

from sdqlpy_b200 import build
from util import QUERY_SCRIPT, ROOT


@pytest.fixture(scope="module")
def lib():
    so = build.compile_file(QUERY_SCRIPT)
    return ctypes.CDLL(so)


def test_exports_declared_symbols(lib):
    hdr = open(os.path.join(ROOT, "include", "sdqlb200.h")).read()
    names = re.findall(r"\b(sdqlb200_[a-z_]+)\s*\(", hdr)
    assert set(names) >= {"sdqlb200_run", "sdqlb200_manifest", "sdqlb200_num_queries", "sdqlb200_query_name",
                          "sdqlb200_result_free", "sdqlb200_last_error"}
    for n in set(names):
        assert hasattr(lib, n), n


def test_manifest_lists_all_queries(lib):
    lib.sdqlb200_manifest.restype = ctypes.c_char_p
    lib.sdqlb200_query_name.restype = ctypes.c_char_p
    man = json.loads(lib.sdqlb200_manifest().decode())
    names = [q["name"] for q in man["queries"]]
    assert names == ["q%d" % i for i in range(1, 23)]
    assert lib.sdqlb200_num_queries() == 22
    assert lib.sdqlb200_query_name(5).decode() == "q6"
    q6 = man["queries"][5]
    assert [tuple(i) for i in q6["inputs"]] == [("li", "l_shipdate", "i32"), ("li", "l_discount", "f64"),
                                                 ("li", "l_quantity", "f64"), ("li", "l_extendedprice", "f64")]


def test_unknown_query_is_an_error(lib):
    from sdqlpy_b200.runtime import Args
    lib.sdqlb200_last_error.restype = ctypes.c_char_p
    a = Args()
    assert lib.sdqlb200_run(b"nope", ctypes.byref(a)) == -4
    assert b"nope" in lib.sdqlb200_last_error()


def test_product_path_has_no_cpu_fallback():
    import torch
    from sdqlpy_b200 import runtime
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    runtime.set_backend(None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        runtime.backend()
