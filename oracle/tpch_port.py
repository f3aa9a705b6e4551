"""TEST INFRASTRUCTURE ONLY -- a CPU restatement (numpy / pandas) of what the reference computes for the 22 TPC-H
programs of /root/reference/test/test_all.py.  It follows the reference's *semantics* (the C++ its generator emits,
sdql_ir_cpp_generator_par.py:176-570 + :712-795, string semantics of include/varchar.h) including the quirks listed
in SURVEY.md section 8(a); every function cites the test_all.py lines it restates.

Pinned: tests/test_oracle_port.py checks every function against tests/golden/*.json, i.e. against outputs of the REAL
reference build (oracle/_ref) on the same seeded inputs.  It exists so that parity can still be checked where the
reference module is unavailable, and at sizes where building UCS4 inputs for the reference is impractical.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module.
Input: ``db`` = {table name: {column name: numpy array}} in the reference layout (int64 / float64 / '<U n').
Output: float, or list of tuples in the reference's result-field order.
"""
import numpy as np
import pandas as pd


def _df(db, table, cols):
    return pd.DataFrame({c: db[table][c] for c in cols})


def _rev(d):
    return d.l_extendedprice * (1.0 - d.l_discount)


def _first_index(col, kw):
    """varchar.h:91-97 firstIndex (wcsstr; -1 if absent)."""
    return np.array([s.find(kw) for s in col], dtype=np.int64)


def q1(db):  # t:46-62
    li = _df(db, "lineitem", ["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax",
                              "l_shipdate"])
    li = li[li.l_shipdate <= 19980902]
    disc = li.l_extendedprice * (1.0 - li.l_discount)
    g = pd.DataFrame({"rf": li.l_returnflag, "ls": li.l_linestatus, "q": li.l_quantity, "b": li.l_extendedprice,
                      "d": disc, "c": disc * (1.0 + li.l_tax), "n": 1}).groupby(["rf", "ls"], sort=False).sum()
    return [(rf, ls, r.q, r.b, r.d, r.c, int(r.n)) for (rf, ls), r in g.iterrows()]


def q2(db):  # t:66-141  (the "min" supplycost is a SUM compared with ==, t:115, 136)
    re = _df(db, "region", ["r_regionkey", "r_name"])
    na = _df(db, "nation", ["n_nationkey", "n_name", "n_regionkey"])
    su = _df(db, "supplier", ["s_suppkey", "s_name", "s_address", "s_nationkey", "s_phone", "s_acctbal", "s_comment"])
    pa = _df(db, "part", ["p_partkey", "p_mfgr", "p_type", "p_size"])
    ps = _df(db, "partsupp", ["ps_partkey", "ps_suppkey", "ps_supplycost"])
    na = na[na.n_regionkey.isin(re[re.r_name == "EUROPE"].r_regionkey)]
    su = su.merge(na, left_on="s_nationkey", right_on="n_nationkey")
    pa = pa[(pa.p_size == 15) & pa.p_type.str.endswith("BRASS")]
    cand = ps[ps.ps_partkey.isin(pa.p_partkey) & ps.ps_suppkey.isin(su.s_suppkey)]
    tot = cand.groupby("ps_partkey", sort=False).ps_supplycost.sum()
    r = ps[ps.ps_partkey.isin(tot.index) & ps.ps_suppkey.isin(su.s_suppkey)]
    r = r[r.ps_supplycost.values == tot.reindex(r.ps_partkey).values]
    r = r.merge(su, left_on="ps_suppkey", right_on="s_suppkey").merge(pa, left_on="ps_partkey", right_on="p_partkey")
    return [(x.s_acctbal, x.s_name, x.n_name, int(x.ps_partkey), x.p_mfgr, x.s_address, x.s_phone, x.s_comment)
            for x in r.itertuples()]


def q3(db):  # t:145-176
    cu = _df(db, "customer", ["c_custkey", "c_mktsegment"])
    od = _df(db, "orders", ["o_orderkey", "o_custkey", "o_orderdate", "o_shippriority"])
    li = _df(db, "lineitem", ["l_orderkey", "l_extendedprice", "l_discount", "l_shipdate"])
    od = od[(od.o_orderdate < 19950315) & od.o_custkey.isin(cu[cu.c_mktsegment == "BUILDING"].c_custkey)]
    li = li[li.l_shipdate > 19950315].merge(od, left_on="l_orderkey", right_on="o_orderkey")
    li["rev"] = _rev(li)
    g = li.groupby(["l_orderkey", "o_orderdate", "o_shippriority"], sort=False).rev.sum()
    return [(int(k[0]), int(k[1]), int(k[2]), v) for k, v in g.items()]


def q4(db):  # t:180-211
    li = _df(db, "lineitem", ["l_orderkey", "l_commitdate", "l_receiptdate"])
    od = _df(db, "orders", ["o_orderkey", "o_orderdate", "o_orderpriority"])
    late = li[li.l_commitdate < li.l_receiptdate].l_orderkey.unique()
    od = od[(od.o_orderdate >= 19930701) & (od.o_orderdate < 19931001) & od.o_orderkey.isin(late)]
    return [(k, int(v)) for k, v in od.groupby("o_orderpriority", sort=False).size().items()]


def q5(db):  # t:215-281
    re = _df(db, "region", ["r_regionkey", "r_name"])
    na = _df(db, "nation", ["n_nationkey", "n_name", "n_regionkey"])
    cu = _df(db, "customer", ["c_custkey", "c_nationkey"])
    od = _df(db, "orders", ["o_orderkey", "o_custkey", "o_orderdate"])
    su = _df(db, "supplier", ["s_suppkey", "s_nationkey"])
    li = _df(db, "lineitem", ["l_orderkey", "l_suppkey", "l_extendedprice", "l_discount"])
    na = na[na.n_regionkey.isin(re[re.r_name == "ASIA"].r_regionkey)]
    cu = cu.merge(na, left_on="c_nationkey", right_on="n_nationkey")
    od = od[(od.o_orderdate < 19950101) & (od.o_orderdate >= 19940101)].merge(cu, left_on="o_custkey", right_on="c_custkey")
    li = li.merge(od, left_on="l_orderkey", right_on="o_orderkey")
    li = li.merge(su, left_on=["l_suppkey", "c_nationkey"], right_on=["s_suppkey", "s_nationkey"])
    li["rev"] = _rev(li)
    return [(k, v) for k, v in li.groupby("n_name", sort=False).rev.sum().items()]


def q6(db):  # t:285-295
    t = db["lineitem"]
    m = ((t["l_shipdate"] >= 19940101) & (t["l_shipdate"] < 19950101) & (t["l_discount"] >= 0.05)
         & (t["l_discount"] <= 0.07) & (t["l_quantity"] < 24.0))
    return float(np.sum(t["l_extendedprice"][m] * t["l_discount"][m]))


def q7(db):  # t:299-367
    na = _df(db, "nation", ["n_nationkey", "n_name"])
    cu = _df(db, "customer", ["c_custkey", "c_nationkey"])
    od = _df(db, "orders", ["o_orderkey", "o_custkey"])
    su = _df(db, "supplier", ["s_suppkey", "s_nationkey"])
    li = _df(db, "lineitem", ["l_orderkey", "l_suppkey", "l_extendedprice", "l_discount", "l_shipdate"])
    na = na[na.n_name.isin(["FRANCE", "GERMANY"])]
    cu = cu.merge(na, left_on="c_nationkey", right_on="n_nationkey").rename(columns={"n_name": "cust_nation"})
    od = od.merge(cu, left_on="o_custkey", right_on="c_custkey")
    su = su.merge(na, left_on="s_nationkey", right_on="n_nationkey").rename(columns={"n_name": "supp_nation"})
    li = li[(li.l_shipdate >= 19950101) & (li.l_shipdate <= 19961231)]
    li = li.merge(od[["o_orderkey", "cust_nation"]], left_on="l_orderkey", right_on="o_orderkey")
    li = li.merge(su[["s_suppkey", "supp_nation"]], left_on="l_suppkey", right_on="s_suppkey")
    li = li[((li.cust_nation == "FRANCE") & (li.supp_nation == "GERMANY"))
            | ((li.cust_nation == "GERMANY") & (li.supp_nation == "FRANCE"))]
    li["rev"], li["yr"] = _rev(li), li.l_shipdate // 10000
    g = li.groupby(["supp_nation", "cust_nation", "yr"], sort=False).rev.sum()
    return [(k[0], k[1], int(k[2]), v) for k, v in g.items()]


def q8(db):  # t:371-427
    re = _df(db, "region", ["r_regionkey", "r_name"])
    na = _df(db, "nation", ["n_nationkey", "n_name", "n_regionkey"])
    su = _df(db, "supplier", ["s_suppkey", "s_nationkey"])
    cu = _df(db, "customer", ["c_custkey", "c_nationkey"])
    pa = _df(db, "part", ["p_partkey", "p_type"])
    od = _df(db, "orders", ["o_orderkey", "o_custkey", "o_orderdate"])
    li = _df(db, "lineitem", ["l_orderkey", "l_partkey", "l_suppkey", "l_extendedprice", "l_discount"])
    america = na[na.n_regionkey.isin(re[re.r_name == "AMERICA"].r_regionkey)].n_nationkey
    li = li[li.l_partkey.isin(pa[pa.p_type == "ECONOMY ANODIZED STEEL"].p_partkey)]
    od = od[(od.o_orderdate >= 19950101) & (od.o_orderdate <= 19961231)]
    li = li.merge(od, left_on="l_orderkey", right_on="o_orderkey").merge(cu, left_on="o_custkey", right_on="c_custkey")
    li = li[li.c_nationkey.isin(america)]
    li = li.merge(su, left_on="l_suppkey", right_on="s_suppkey").merge(na, left_on="s_nationkey", right_on="n_nationkey")
    li["B"] = _rev(li)
    li["A"] = np.where(li.n_name == "BRAZIL", li.B, 0.0)
    g = li.assign(yr=li.o_orderdate // 10000).groupby("yr", sort=False)[["A", "B"]].sum()
    return [(int(k), r.A / r.B) for k, r in g.iterrows()]


def q9(db):  # t:431-491
    na = _df(db, "nation", ["n_nationkey", "n_name"])
    su = _df(db, "supplier", ["s_suppkey", "s_nationkey"]).merge(na, left_on="s_nationkey", right_on="n_nationkey")
    pa = _df(db, "part", ["p_partkey", "p_name"])
    ps = _df(db, "partsupp", ["ps_partkey", "ps_suppkey", "ps_supplycost"])
    od = _df(db, "orders", ["o_orderkey", "o_orderdate"])
    li = _df(db, "lineitem", ["l_orderkey", "l_partkey", "l_suppkey", "l_quantity", "l_extendedprice", "l_discount"])
    green = pa[pa.p_name.str.contains("green", regex=False)].p_partkey
    ps = ps[ps.ps_partkey.isin(green)].merge(su[["s_suppkey", "n_name"]], left_on="ps_suppkey", right_on="s_suppkey")
    li = li.merge(ps, left_on=["l_partkey", "l_suppkey"], right_on=["ps_partkey", "ps_suppkey"])
    li = li.merge(od, left_on="l_orderkey", right_on="o_orderkey")
    li["profit"] = li.l_extendedprice * (1.0 - li.l_discount) - li.ps_supplycost * li.l_quantity
    g = li.assign(yr=li.o_orderdate // 10000).groupby(["n_name", "yr"], sort=False).profit.sum()
    return [(k[0], int(k[1]), v) for k, v in g.items()]


def q10(db):  # t:495-558
    na = _df(db, "nation", ["n_nationkey", "n_name"])
    cu = _df(db, "customer", ["c_custkey", "c_name", "c_acctbal", "c_address", "c_nationkey", "c_phone", "c_comment"])
    od = _df(db, "orders", ["o_orderkey", "o_custkey", "o_orderdate"])
    li = _df(db, "lineitem", ["l_orderkey", "l_extendedprice", "l_discount", "l_returnflag"])
    od = od[(od.o_orderdate >= 19931001) & (od.o_orderdate < 19940101)]
    od = od.merge(cu, left_on="o_custkey", right_on="c_custkey").merge(na, left_on="c_nationkey", right_on="n_nationkey")
    li = li[li.l_returnflag == "R"].merge(od, left_on="l_orderkey", right_on="o_orderkey")
    li["rev"] = _rev(li)
    keys = ["c_custkey", "c_name", "c_acctbal", "n_name", "c_address", "c_phone", "c_comment"]
    g = li.groupby(keys, sort=False).rev.sum()
    return [(int(k[0]), k[1], v, k[2], k[3], k[4], k[5], k[6]) for k, v in g.items()]


def q11(db):  # t:562-604  (A sums per-row products already scaled by 0.0001, t:587)
    na = _df(db, "nation", ["n_nationkey", "n_name"])
    su = _df(db, "supplier", ["s_suppkey", "s_nationkey"])
    ps = _df(db, "partsupp", ["ps_partkey", "ps_suppkey", "ps_availqty", "ps_supplycost"])
    su = su[su.s_nationkey.isin(na[na.n_name == "GERMANY"].n_nationkey)]
    ps = ps[ps.ps_suppkey.isin(su.s_suppkey)]
    val = ps.ps_supplycost * ps.ps_availqty
    a = float(np.sum(val.values * 0.0001))
    g = val.groupby(ps.ps_partkey, sort=False).sum()
    return [(int(k), v) for k, v in g.items() if v > a]


def q12(db):  # t:608-652
    od = _df(db, "orders", ["o_orderkey", "o_orderpriority"])
    li = _df(db, "lineitem", ["l_orderkey", "l_shipdate", "l_commitdate", "l_receiptdate", "l_shipmode"])
    li = li[li.l_shipmode.isin(["MAIL", "SHIP"]) & (li.l_receiptdate >= 19940101) & (li.l_receiptdate < 19950101)
            & (li.l_shipdate < li.l_commitdate) & (li.l_commitdate < li.l_receiptdate)]
    li = li.merge(od, left_on="l_orderkey", right_on="o_orderkey")
    hi = li.o_orderpriority.isin(["1-URGENT", "2-HIGH"])
    g = pd.DataFrame({"m": li.l_shipmode, "h": hi.astype(np.int64), "l": (~hi).astype(np.int64)}).groupby("m", sort=False).sum()
    return [(k, int(r.h), int(r.l)) for k, r in g.iterrows()]


def q13(db):  # t:656-691
    cu = _df(db, "customer", ["c_custkey"])
    od = _df(db, "orders", ["o_custkey", "o_comment"])
    sp, rq = _first_index(od.o_comment, "special"), _first_index(od.o_comment, "requests")
    keep = ~((sp != -1) & (rq > sp + 6))
    cnt = od[keep].groupby("o_custkey", sort=False).size()
    c_count = cnt.reindex(cu.c_custkey).fillna(0).astype(np.int64)
    return [(int(k), int(v)) for k, v in c_count.groupby(c_count.values, sort=False).size().items()]


def q14(db):  # t:695-716
    pa = _df(db, "part", ["p_partkey", "p_type"])
    li = _df(db, "lineitem", ["l_partkey", "l_extendedprice", "l_discount", "l_shipdate"])
    li = li[(li.l_shipdate >= 19950901) & (li.l_shipdate < 19951001)]
    rev = _rev(li)
    promo = li.l_partkey.isin(pa[pa.p_type.str.startswith("PROMO")].p_partkey)
    return float(100.0 * np.sum(np.where(promo, rev, 0.0)) / np.sum(rev.values))


def q15(db, max_revenue=1772627.2087):  # t:720-756
    li = _df(db, "lineitem", ["l_suppkey", "l_extendedprice", "l_discount", "l_shipdate"])
    su = _df(db, "supplier", ["s_suppkey", "s_name", "s_address", "s_phone"]).set_index("s_suppkey")
    li = li[(li.l_shipdate >= 19960101) & (li.l_shipdate < 19960401)]
    g = _rev(li).groupby(li.l_suppkey, sort=False).sum()
    return [(int(k), su.s_name[k], su.s_address[k], su.s_phone[k], v) for k, v in g.items() if v == max_revenue]


def q16(db):  # t:760-827
    pa = _df(db, "part", ["p_partkey", "p_brand", "p_type", "p_size"])
    su = _df(db, "supplier", ["s_suppkey", "s_comment"])
    ps = _df(db, "partsupp", ["ps_partkey", "ps_suppkey"])
    pa = pa[(pa.p_brand != "Brand#45") & ~pa.p_type.str.startswith("MEDIUM POLISHED")
            & pa.p_size.isin([49, 14, 23, 45, 19, 3, 36, 9])]
    cu, co = _first_index(su.s_comment, "Customer"), _first_index(su.s_comment, "Complaints")
    bad = su[(cu != -1) & (co > cu + 7)].s_suppkey
    ps = ps[~ps.ps_suppkey.isin(bad)].merge(pa, left_on="ps_partkey", right_on="p_partkey")
    g = ps.groupby(["p_brand", "p_type", "p_size"], sort=False).ps_suppkey.nunique()
    return [(k[0], k[1], int(k[2]), int(v)) for k, v in g.items()]


def q17(db):  # t:831-870
    pa = _df(db, "part", ["p_partkey", "p_brand", "p_container"])
    li = _df(db, "lineitem", ["l_partkey", "l_quantity", "l_extendedprice"])
    li = li[li.l_partkey.isin(pa[(pa.p_brand == "Brand#23") & (pa.p_container == "MED BOX")].p_partkey)]
    g = li.groupby("l_partkey", sort=False).l_quantity.agg(["sum", "count"])
    avg = (g["sum"] / g["count"].astype(np.float64)).reindex(li.l_partkey).values
    return float(np.sum(li.l_extendedprice.values[(0.2 * avg) > li.l_quantity.values]) / 7.0)


def q18(db):  # t:874-913
    li = _df(db, "lineitem", ["l_orderkey", "l_quantity"])
    cu = _df(db, "customer", ["c_custkey", "c_name"])
    od = _df(db, "orders", ["o_orderkey", "o_custkey", "o_orderdate", "o_totalprice"])
    tot = li.groupby("l_orderkey", sort=False).l_quantity.sum()
    od = od[od.o_orderkey.isin(tot[tot > 300].index)].merge(cu, left_on="o_custkey", right_on="c_custkey")
    li = li.merge(od, left_on="l_orderkey", right_on="o_orderkey")
    g = li.groupby(["c_name", "o_custkey", "o_orderkey", "o_orderdate", "o_totalprice"], sort=False).l_quantity.sum()
    return [(k[0], int(k[1]), int(k[2]), int(k[3]), k[4], v) for k, v in g.items()]


def q19(db):  # t:917-981  (note "AIR REG", which dbgen data never contains, t:940)
    pa = _df(db, "part", ["p_partkey", "p_brand", "p_size", "p_container"])
    li = _df(db, "lineitem", ["l_partkey", "l_quantity", "l_extendedprice", "l_discount", "l_shipinstruct", "l_shipmode"])
    pa = pa[((pa.p_brand == "Brand#12") & pa.p_container.isin(["SM CASE", "SM BOX", "SM PACK", "SM PKG"]) & (pa.p_size >= 1) & (pa.p_size <= 5))
            | ((pa.p_brand == "Brand#23") & pa.p_container.isin(["MED BAG", "MED BOX", "MED PACK", "MED PKG"]) & (pa.p_size >= 1) & (pa.p_size <= 10))
            | ((pa.p_brand == "Brand#34") & pa.p_container.isin(["LG CASE", "LG BOX", "LG PACK", "LG PKG"]) & (pa.p_size >= 1) & (pa.p_size <= 15))]
    li = li[(li.l_shipinstruct == "DELIVER IN PERSON") & li.l_shipmode.isin(["AIR", "AIR REG"])]
    li = li.merge(pa, left_on="l_partkey", right_on="p_partkey")
    ok = (((li.p_brand == "Brand#12") & (li.l_quantity >= 1) & (li.l_quantity <= 11))
          | ((li.p_brand == "Brand#23") & (li.l_quantity >= 10) & (li.l_quantity <= 20))
          | ((li.p_brand == "Brand#34") & (li.l_quantity >= 20) & (li.l_quantity <= 30)))
    return [(float(np.sum(_rev(li).values[ok.values])),)]


def q20(db):  # t:985-1028
    su = _df(db, "supplier", ["s_suppkey", "s_name", "s_address", "s_nationkey"])
    na = _df(db, "nation", ["n_nationkey", "n_name"])
    ps = _df(db, "partsupp", ["ps_partkey", "ps_suppkey", "ps_availqty"])
    pa = _df(db, "part", ["p_partkey", "p_name"])
    li = _df(db, "lineitem", ["l_partkey", "l_suppkey", "l_quantity", "l_shipdate"])
    forest = pa[pa.p_name.str.startswith("forest")].p_partkey
    canada = su[su.s_nationkey.isin(na[na.n_name == "CANADA"].n_nationkey)].s_suppkey
    li = li[(li.l_shipdate >= 19940101) & (li.l_shipdate < 19950101) & li.l_partkey.isin(forest) & li.l_suppkey.isin(canada)]
    half = (0.5 * li.l_quantity).groupby([li.l_partkey, li.l_suppkey], sort=False).sum().rename("half").reset_index()
    ps = ps.merge(half, left_on=["ps_partkey", "ps_suppkey"], right_on=["l_partkey", "l_suppkey"])
    keys = ps[ps.ps_availqty > ps.half].ps_suppkey.unique()
    su = su[su.s_suppkey.isin(keys)]
    return list(dict.fromkeys(zip(su.s_name, su.s_address)))


def q21(db):  # t:1032-1109  (list lengths, not distinct suppliers, t:1101-1102)
    su = _df(db, "supplier", ["s_suppkey", "s_name", "s_nationkey"])
    na = _df(db, "nation", ["n_nationkey", "n_name"])
    od = _df(db, "orders", ["o_orderkey", "o_orderstatus"])
    li = _df(db, "lineitem", ["l_orderkey", "l_suppkey", "l_commitdate", "l_receiptdate"])
    su = su[su.s_nationkey.isin(na[na.n_name == "SAUDI ARABIA"].n_nationkey)]
    l2 = li.groupby("l_orderkey", sort=False).size()
    late = li[li.l_receiptdate > li.l_commitdate]
    l3 = late.groupby("l_orderkey", sort=False).size()
    l1 = late[late.l_orderkey.isin(od[od.o_orderstatus == "F"].o_orderkey)].merge(su, left_on="l_suppkey", right_on="s_suppkey")
    n2 = l2.reindex(l1.l_orderkey).fillna(0).values
    n3 = l3.reindex(l1.l_orderkey).fillna(0).values
    l1 = l1[(n2 > 1) & ~((n3 > 0) & (n3 > 1))]
    return [(k, int(v)) for k, v in l1.groupby("s_name", sort=False).size().items()]


def q22(db):  # t:1113-1183
    cu = _df(db, "customer", ["c_custkey", "c_phone", "c_acctbal"])
    od = _df(db, "orders", ["o_custkey"])
    code = cu.c_phone.str[:2]
    sel = code.isin(["13", "31", "23", "29", "30", "18", "17"])
    inner = cu[(cu.c_acctbal > 0.0) & sel]
    avg = inner.c_acctbal.sum() / float(len(inner))
    r = cu[(cu.c_acctbal > avg) & ~cu.c_custkey.isin(od.o_custkey) & sel]
    g = r.groupby(code[r.index], sort=False).c_acctbal.agg(["size", "sum"])
    return [(k, int(x["size"]), x["sum"]) for k, x in g.iterrows()]


QUERIES = {"q%d" % i: globals()["q%d" % i] for i in range(1, 23)}


def make_db(tpch, tables=None):
    """{table: {column: array}} in the reference layout from the repo's generator."""
    from sdqlpy_b200.tpch.gen import SCHEMAS
    out = {}
    for t in tables or SCHEMAS:
        cols = tpch.columns(t)
        out[t] = {c: v.to_ref() for c, v in cols.items()}
    return out
