"""front end: the 22 workload programs translate to the IR shapes the reference compiler produces."""
from sdqlpy_b200 import frontend, ir
from util import QUERY_SCRIPT


def _ir(name):
    funcs, consts = frontend.parse_module(open(QUERY_SCRIPT).read())
    return frontend.function_to_ir(funcs[name][0], consts)


def _sums(e, out):
    if isinstance(e, ir.SumExpr):
        out.append(e)
    for c in e.children():
        _sums(c, out)
    return out


def test_all_queries_translate():
    funcs, consts = frontend.parse_module(open(QUERY_SCRIPT).read())
    assert sorted(funcs) == sorted("q%d" % i for i in range(1, 23))
    for f in funcs.values():
        root, args = frontend.function_to_ir(f[0], consts)
        assert isinstance(root, ir.LetExpr)


def test_q6_shape():
    root, args = _ir("q6")
    assert args == ["li"]
    s = root.valExpr
    assert isinstance(s, ir.SumExpr) and s.dictExpr.name == "db->li_dataset" and not s.isAssignmentSum
    assert isinstance(s.bodyExpr, ir.IfExpr) and isinstance(s.bodyExpr.condExpr, ir.MulExpr)  # `and` -> *  (comp:277-292)
    assert root.bodyExpr.varExpr.name == "out"


def test_join_build_probe_sugar():
    root, _ = _ir("q3")
    sums = _sums(root, [])
    assert [s.isAssignmentSum for s in sums] == [True, True, False, True]   # build, probe(False), probe, unique
    probe = root.bodyExpr.bodyExpr.valExpr                                    # order_probed = LetExpr(probeVar, ..)
    assert isinstance(probe, ir.LetExpr) and isinstance(probe.bodyExpr, ir.SumExpr)
    inner = probe.bodyExpr.bodyExpr
    assert isinstance(inner, ir.IfExpr) and isinstance(inner.thenBodyExpr, ir.IfExpr)
    assert inner.thenBodyExpr.condExpr.compareType == ir.CompareSymbol.NE    # probe != None (ir:450)


def test_dense_and_unique_flags():
    root, _ = _ir("q4")
    s = root.valExpr
    assert s.isAssignmentSum and s.dictType == "dense_array(6000000)"
    root, _ = _ir("q21")
    sums = _sums(root, [])
    dense = [s for s in sums if s.dictType.startswith("dense_array")]
    assert len(dense) == 3 and [s.isAssignmentSum for s in dense] == [True, False, False]


def test_contains_and_negative_literal():
    root, _ = _ir("q9")
    found = []

    def walk(e):
        if isinstance(e, ir.ExtFuncExpr) and e.symbol == ir.ExtFuncSymbol.StringContains:
            found.append(e)
        for c in e.children():
            walk(c)
    walk(root)
    assert len(found) == 1 and found[0].inp2.value == -1
