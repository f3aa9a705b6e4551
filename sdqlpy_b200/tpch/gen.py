"""Synthetic TPC-H-schema generator (no dbgen and no network in this environment).

dbgen-like distributions, driven by a *counter-based* generator: every value
is a pure function ``mix(seed, stream, index)`` of its row (or order) index, so any row range can be
produced independently -- on the host in numpy (this file) or on the device (csrc/tpch_gen.cu uses the
same integer arithmetic, so both produce bit-identical columns).

Two layouts:
  * ``compact``  -- what the B200 backend keeps resident: int32 keys/dates, float64 measures,
                    uint8 dictionary codes (+ the dictionary) for low-cardinality strings,
                    fixed-width ASCII byte matrices for free text.
  * ``ref``      -- what the reference's ``read_csv`` produces (sdql_lib.py:83-97): int64 / float64 /
                    numpy ``<U n`` (UCS4).  Schemas follow test/test_all.py:26-33 incl. the trailing
                    ``*_NA`` column.

Row order: orders by o_orderkey, lineitem clustered by order (as dbgen), so both can be range
partitioned on order boundaries for multi-GPU runs.
"""
import numpy as np

SEED = 20231017
U64 = np.uint64

# ----------------------------------------------------------------------------------------------
# schemas (column name, kind) -- kind: 'int' | 'date' | 'float' | ('str', n)
# ----------------------------------------------------------------------------------------------
SCHEMAS = {
    "lineitem": [("l_orderkey", "int"), ("l_partkey", "int"), ("l_suppkey", "int"), ("l_linenumber", "int"),
                 ("l_quantity", "float"), ("l_extendedprice", "float"), ("l_discount", "float"), ("l_tax", "float"),
                 ("l_returnflag", ("str", 1)), ("l_linestatus", ("str", 1)), ("l_shipdate", "date"),
                 ("l_commitdate", "date"), ("l_receiptdate", "date"), ("l_shipinstruct", ("str", 25)),
                 ("l_shipmode", ("str", 10)), ("l_comment", ("str", 44)), ("l_NA", ("str", 1))],
    "customer": [("c_custkey", "int"), ("c_name", ("str", 25)), ("c_address", ("str", 40)), ("c_nationkey", "int"),
                 ("c_phone", ("str", 15)), ("c_acctbal", "float"), ("c_mktsegment", ("str", 10)),
                 ("c_comment", ("str", 117)), ("c_NA", ("str", 1))],
    "orders": [("o_orderkey", "int"), ("o_custkey", "int"), ("o_orderstatus", ("str", 1)), ("o_totalprice", "float"),
               ("o_orderdate", "date"), ("o_orderpriority", ("str", 15)), ("o_clerk", ("str", 15)),
               ("o_shippriority", "int"), ("o_comment", ("str", 79)), ("o_NA", ("str", 1))],
    "nation": [("n_nationkey", "int"), ("n_name", ("str", 25)), ("n_regionkey", "int"), ("n_comment", ("str", 152)),
               ("n_NA", ("str", 1))],
    "region": [("r_regionkey", "int"), ("r_name", ("str", 25)), ("r_comment", ("str", 152)), ("r_NA", ("str", 1))],
    "part": [("p_partkey", "int"), ("p_name", ("str", 55)), ("p_mfgr", ("str", 25)), ("p_brand", ("str", 10)),
             ("p_type", ("str", 25)), ("p_size", "int"), ("p_container", ("str", 10)), ("p_retailprice", "float"),
             ("p_comment", ("str", 23)), ("p_NA", ("str", 1))],
    "partsupp": [("ps_partkey", "int"), ("ps_suppkey", "int"), ("ps_availqty", "float"), ("ps_supplycost", "float"),
                 ("ps_comment", ("str", 199)), ("ps_NA", ("str", 1))],
    "supplier": [("s_suppkey", "int"), ("s_name", ("str", 25)), ("s_address", ("str", 40)), ("s_nationkey", "int"),
                 ("s_phone", ("str", 15)), ("s_acctbal", "float"), ("s_comment", ("str", 101)), ("s_NA", ("str", 1))],
}

NATIONS = [("ALGERIA", 0), ("ARGENTINA", 1), ("BRAZIL", 1), ("CANADA", 1), ("EGYPT", 4), ("ETHIOPIA", 0),
           ("FRANCE", 3), ("GERMANY", 3), ("INDIA", 2), ("INDONESIA", 2), ("IRAN", 4), ("IRAQ", 4), ("JAPAN", 2),
           ("JORDAN", 4), ("KENYA", 0), ("MOROCCO", 0), ("MOZAMBIQUE", 0), ("PERU", 1), ("CHINA", 2), ("ROMANIA", 3),
           ("SAUDI ARABIA", 4), ("VIETNAM", 2), ("RUSSIA", 3), ("UNITED KINGDOM", 3), ("UNITED STATES", 1)]
REGIONS = ["AFRICA", "AMERICA", "ASIA", "EUROPE", "MIDDLE EAST"]
SEGMENTS = ["AUTOMOBILE", "BUILDING", "FURNITURE", "MACHINERY", "HOUSEHOLD"]
PRIORITIES = ["1-URGENT", "2-HIGH", "3-MEDIUM", "4-NOT SPECIFIED", "5-LOW"]
INSTRUCTS = ["DELIVER IN PERSON", "COLLECT COD", "NONE", "TAKE BACK RETURN"]
MODES = ["REG AIR", "AIR", "RAIL", "SHIP", "TRUCK", "MAIL", "FOB"]
RFLAGS = ["R", "A", "N"]
LSTATUS = ["O", "F"]
OSTATUS = ["F", "O", "P"]
TYPE_S1 = ["STANDARD", "SMALL", "MEDIUM", "LARGE", "ECONOMY", "PROMO"]
TYPE_S2 = ["ANODIZED", "BURNISHED", "PLATED", "POLISHED", "BRUSHED"]
TYPE_S3 = ["TIN", "NICKEL", "BRASS", "STEEL", "COPPER"]
TYPES = [a + " " + b + " " + c for a in TYPE_S1 for b in TYPE_S2 for c in TYPE_S3]
CONT_S1 = ["SM", "LG", "MED", "JUMBO", "WRAP"]
CONT_S2 = ["CASE", "BOX", "BAG", "JAR", "PKG", "PACK", "CAN", "DRUM"]
CONTAINERS = [a + " " + b for a in CONT_S1 for b in CONT_S2]
BRANDS = ["Brand#%d%d" % (m, n) for m in range(1, 6) for n in range(1, 6)]
MFGRS = ["Manufacturer#%d" % m for m in range(1, 6)]
COLORS = ("almond antique aquamarine azure beige bisque black blanched blue blush brown burlywood burnished "
          "chartreuse chiffon chocolate coral cornflower cornsilk cream cyan dark deep dim dodger drab firebrick "
          "floral forest frosted gainsboro ghost goldenrod green grey honeydew hot indian ivory khaki lace lavender "
          "lawn lemon light lime linen magenta maroon medium metallic midnight mint misty moccasin navajo navy olive "
          "orange orchid pale papaya peach peru pink plum powder puff purple red rose rosy royal saddle salmon sandy "
          "seashell sienna sky slate smoke snow spring steel tan thistle tomato turquoise violet wheat white "
          "yellow").split()
assert len(COLORS) == 92
WORDS = ("furiously carefully quickly blithely slyly boldly fluffily ironic regular express pending final "
         "bold silent even unusual deposits accounts packages theodolites foxes ideas pinto beans platelets "
         "asymptotes courts dolphins sleep wake nag haggle cajole detect integrate among above beside the").split()

# stream ids ------------------------------------------------------------------------------------
(S_ORD_NL, S_ORD_CUST, S_ORD_DATE, S_ORD_PRIO, S_ORD_CMT, S_ORD_CMTK, S_L_PART, S_L_SUPPJ, S_L_QTY, S_L_DISC,
 S_L_TAX, S_L_SHIP, S_L_COMMIT, S_L_RECEIPT, S_L_RFLAG, S_L_INSTR, S_L_MODE, S_PS_QTY, S_PS_COST, S_P_NAME,
 S_P_MFGR, S_P_BRAND, S_P_TYPE, S_P_SIZE, S_P_CONT, S_C_NAT, S_C_PHONE, S_C_BAL, S_C_SEG, S_C_TXT, S_S_NAT,
 S_S_PHONE, S_S_BAL, S_S_CMT, S_S_CMTK, S_S_TXT) = range(1, 37)


def mix(stream, idx, seed=SEED):
    """splitmix64 finaliser over (seed, stream, idx) -- uint64 in, uint64 out (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        x = np.asarray(idx).astype(U64) + U64((seed * 0x9E3779B97F4A7C15 + stream * 0xD1B54A32D192ED03) & (2**64 - 1))
        x = (x ^ (x >> U64(30))) * U64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> U64(27))) * U64(0x94D049BB133111EB)
        x = x ^ (x >> U64(31))
    return x


def rint(stream, idx, lo, hi, seed=SEED):
    """uniform integer in [lo, hi] as int64."""
    return (lo + ((mix(stream, idx, seed) >> U64(11)) % U64(hi - lo + 1)).astype(np.int64))


def days_to_ymd(z):
    """days since 1970-01-01 -> int YYYYMMDD (civil-from-days, pure integer arithmetic)."""
    z = np.asarray(z, dtype=np.int64) + 719468
    era = z // 146097
    doe = z - era * 146097
    yoe = (doe - doe // 1460 + doe // 36524 - doe // 146096) // 365
    y = yoe + era * 400
    doy = doe - (365 * yoe + yoe // 4 - yoe // 100)
    mp = (5 * doy + 2) // 153
    d = doy - (153 * mp + 2) // 5 + 1
    m = np.where(mp < 10, mp + 3, mp - 9)
    y = np.where(m <= 2, y + 1, y)
    return y * 10000 + m * 100 + d


DAY_1992_01_01 = 8035  # days since epoch
DATE_SPAN = 2406  # 1992-01-01 .. 1998-08-02 inclusive
DAY_1995_06_17 = 9298


def retail_cents(pk):
    return 90000 + (pk // 10) % 20001 + 100 * (pk % 1000)


def _bytes_table(words, width):
    t = np.zeros((len(words), width), dtype=np.uint8)
    for i, w in enumerate(words):
        b = w.encode("ascii")
        t[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    return t


def bytes_to_ustr(mat, n):
    """uint8 [rows, w<=n] zero padded -> numpy '<U n' array (reference layout)."""
    rows, w = mat.shape
    out = np.zeros((rows, n), dtype=np.uint32)
    out[:, :w] = mat
    return out.view("<U%d" % n).reshape(rows)


def _digits(vals, ndig):
    """int array -> uint8 [rows, ndig] ASCII digits (zero padded)."""
    vals = np.asarray(vals, dtype=np.int64)
    out = np.empty((len(vals), ndig), dtype=np.uint8)
    for k in range(ndig):
        out[:, ndig - 1 - k] = 48 + (vals // (10 ** k)) % 10
    return out


def _prefixed_number(prefix, vals, ndig):
    p = np.frombuffer(prefix.encode(), dtype=np.uint8)
    out = np.empty((len(vals), len(p) + ndig), dtype=np.uint8)
    out[:, :len(p)] = p
    out[:, len(p):] = _digits(vals, ndig)
    return out


def _phone(stream, idx, nationkey, seed):
    r = mix(stream, idx, seed)
    out = np.empty((len(idx), 15), dtype=np.uint8)
    out[:, 0:2] = _digits(nationkey + 10, 2)
    out[:, 2] = 45
    out[:, 3:6] = _digits(100 + (r % U64(900)).astype(np.int64), 3)
    out[:, 6] = 45
    out[:, 7:10] = _digits(100 + ((r >> U64(16)) % U64(900)).astype(np.int64), 3)
    out[:, 10] = 45
    out[:, 11:15] = _digits(1000 + ((r >> U64(32)) % U64(9000)).astype(np.int64), 4)
    return out


def _text(stream, idx, nslots, width, seed):
    """free text: nslots vocabulary words in 10-byte slots, cut to a random length, zero padded to width."""
    slot = 12
    vocab = _bytes_table([w + " " for w in WORDS], slot)
    vocab[vocab == 0] = 32
    n = len(idx)
    out = np.zeros((n, width), dtype=np.uint8)
    base = np.asarray(idx).astype(U64) * U64(nslots)
    for s in range(nslots):
        code = (mix(stream, base + U64(s), seed) % U64(len(WORDS))).astype(np.int64)
        out[:, s * slot:(s + 1) * slot] = vocab[code]
    return out


def _cut(mat, stream, idx, lo, hi, seed, keep=None, keep_len=0):
    ln = rint(stream, idx, lo, hi, seed)
    if keep is not None:
        ln = np.where(keep, np.maximum(ln, keep_len), ln)
    col = np.arange(mat.shape[1])[None, :]
    mat[col >= ln[:, None]] = 0
    return mat


def _put(mat, rows, start, word):
    b = np.frombuffer(word.encode(), dtype=np.uint8)
    mat[rows, start:start + len(b)] = b


class Column:
    """One generated column in compact form.  kind: 'i32' | 'i64' | 'f64' | 'code' | 'bytes'."""
    __slots__ = ("name", "kind", "data", "dictionary", "width", "wire")

    def __init__(self, name, kind, data, dictionary=None, width=None):
        self.name, self.kind, self.data, self.dictionary, self.width = name, kind, data, dictionary, width
        self.wire = None  # optional lossless packed image for the host -> device link (sdqlpy_b200.wire.Packed)

    def to_ref(self):
        if self.kind in ("i32", "i64"):
            return self.data.astype(np.int64)
        if self.kind == "f64":
            return self.data
        if self.kind == "code":
            d = np.array(self.dictionary, dtype="<U%d" % self.width)
            return d[self.data]
        if self.kind == "bytes":
            return bytes_to_ustr(self.data, self.width)
        raise ValueError(self.kind)


class TPCH:
    """Lazy generator of TPC-H tables at scale factor ``sf`` (float; 0.01 is fine for tests)."""

    def __init__(self, sf=1.0, seed=SEED):
        self.sf, self.seed = float(sf), int(seed)
        self.S = max(4, int(round(10000 * sf)))
        self.P = max(4, int(round(200000 * sf)))
        self.C = max(3, int(round(150000 * sf)))
        self.O = max(1, int(round(1500000 * sf)))
        self._nl_cache = None

    # -- sizes ---------------------------------------------------------------------------------
    def order_lines(self):
        """(lines per order [O], exclusive prefix offsets [O+1])."""
        if self._nl_cache is None:
            nl = rint(S_ORD_NL, np.arange(self.O), 1, 7, self.seed)
            off = np.zeros(self.O + 1, dtype=np.int64)
            np.cumsum(nl, out=off[1:])
            self._nl_cache = (nl, off)
        return self._nl_cache

    def rows(self, table):
        if table == "lineitem":
            return int(self.order_lines()[1][-1])
        return {"orders": self.O, "customer": self.C, "supplier": self.S, "part": self.P,
                "partsupp": 4 * self.P, "nation": 25, "region": 5}[table]

    def max_orderkey(self):
        i = self.O - 1
        return (i // 8) * 32 + i % 8 + 1

    # -- column generation ---------------------------------------------------------------------
    def columns(self, table, cols=None, order_range=None):
        """dict name -> Column for the requested columns (None = all columns queries can touch).
        ``order_range=(o0, o1)`` restricts orders/lineitem to that range of order indices."""
        fn = getattr(self, "_gen_" + table)
        want = None if cols is None else set(cols)
        if table in ("orders", "lineitem"):
            return fn(want, order_range or (0, self.O))
        return fn(want)

    def _ord_keys(self, i):
        return (i // 8) * 32 + i % 8 + 1

    def _ord_dates_days(self, i):
        return DAY_1992_01_01 + rint(S_ORD_DATE, i, 0, DATE_SPAN - 1, self.seed)

    def _gen_lineitem(self, want, orng):
        o0, o1 = orng
        nl, off = self.order_lines()
        oi = np.repeat(np.arange(o0, o1), nl[o0:o1])
        r = np.arange(off[o0], off[o1])  # global lineitem row ids
        out = {}

        def need(c):
            return want is None or c in want

        seed = self.seed
        if need("l_orderkey"):
            out["l_orderkey"] = Column("l_orderkey", "i32", self._ord_keys(oi).astype(np.int32))
        if need("l_linenumber"):
            out["l_linenumber"] = Column("l_linenumber", "i32", (r - off[oi] + 1).astype(np.int32))
        pk = None
        if any(need(c) for c in ("l_partkey", "l_suppkey", "l_extendedprice")):
            pk = rint(S_L_PART, r, 1, self.P, seed)
        if need("l_partkey"):
            out["l_partkey"] = Column("l_partkey", "i32", pk.astype(np.int32))
        if need("l_suppkey"):
            j = rint(S_L_SUPPJ, r, 0, 3, seed)
            sk = (pk + j * (self.S // 4 + (pk - 1) // self.S)) % self.S + 1
            out["l_suppkey"] = Column("l_suppkey", "i32", sk.astype(np.int32))
        qty = None
        if need("l_quantity") or need("l_extendedprice"):
            qty = rint(S_L_QTY, r, 1, 50, seed)
        if need("l_quantity"):
            out["l_quantity"] = Column("l_quantity", "f64", qty.astype(np.float64))
        if need("l_extendedprice"):
            out["l_extendedprice"] = Column("l_extendedprice", "f64", (qty * retail_cents(pk)) / 100.0)
        if need("l_discount"):
            out["l_discount"] = Column("l_discount", "f64", rint(S_L_DISC, r, 0, 10, seed) / 100.0)
        if need("l_tax"):
            out["l_tax"] = Column("l_tax", "f64", rint(S_L_TAX, r, 0, 8, seed) / 100.0)
        ship = receipt = None
        if any(need(c) for c in ("l_shipdate", "l_commitdate", "l_receiptdate", "l_returnflag", "l_linestatus")):
            od = self._ord_dates_days(oi)
            ship = od + rint(S_L_SHIP, r, 1, 121, seed)
            receipt = ship + rint(S_L_RECEIPT, r, 1, 30, seed)
            if need("l_shipdate"):
                out["l_shipdate"] = Column("l_shipdate", "i32", days_to_ymd(ship).astype(np.int32))
            if need("l_commitdate"):
                out["l_commitdate"] = Column("l_commitdate", "i32",
                                             days_to_ymd(od + rint(S_L_COMMIT, r, 30, 90, seed)).astype(np.int32))
            if need("l_receiptdate"):
                out["l_receiptdate"] = Column("l_receiptdate", "i32", days_to_ymd(receipt).astype(np.int32))
        if need("l_returnflag"):
            ra = (mix(S_L_RFLAG, r, seed) & U64(1)).astype(np.uint8)
            out["l_returnflag"] = Column("l_returnflag", "code",
                                         np.where(receipt <= DAY_1995_06_17, ra, 2).astype(np.uint8), RFLAGS, 1)
        if need("l_linestatus"):
            out["l_linestatus"] = Column("l_linestatus", "code",
                                         np.where(ship > DAY_1995_06_17, 0, 1).astype(np.uint8), LSTATUS, 1)
        if need("l_shipinstruct"):
            out["l_shipinstruct"] = Column("l_shipinstruct", "code",
                                           rint(S_L_INSTR, r, 0, 3, seed).astype(np.uint8), INSTRUCTS, 25)
        if need("l_shipmode"):
            out["l_shipmode"] = Column("l_shipmode", "code", rint(S_L_MODE, r, 0, 6, seed).astype(np.uint8), MODES, 10)
        return out

    def _gen_orders(self, want, orng):
        o0, o1 = orng
        i = np.arange(o0, o1)
        seed = self.seed
        out = {}

        def need(c):
            return want is None or c in want

        if need("o_orderkey"):
            out["o_orderkey"] = Column("o_orderkey", "i32", self._ord_keys(i).astype(np.int32))
        if need("o_custkey"):
            # uniform over custkeys with key % 3 != 0 (a third of customers never order, as in dbgen)
            m = self.C - self.C // 3
            k = rint(S_ORD_CUST, i, 0, m - 1, seed)
            out["o_custkey"] = Column("o_custkey", "i32", (k + k // 2 + 1).astype(np.int32))
        if need("o_orderdate"):
            out["o_orderdate"] = Column("o_orderdate", "i32", days_to_ymd(self._ord_dates_days(i)).astype(np.int32))
        if need("o_orderpriority"):
            out["o_orderpriority"] = Column("o_orderpriority", "code", rint(S_ORD_PRIO, i, 0, 4, seed).astype(np.uint8),
                                            PRIORITIES, 15)
        if need("o_shippriority"):
            out["o_shippriority"] = Column("o_shippriority", "i32", np.zeros(len(i), dtype=np.int32))
        if need("o_orderstatus") or need("o_totalprice"):
            li = self._gen_lineitem({"l_linestatus", "l_extendedprice", "l_discount", "l_tax"}, orng)
            nl, off = self.order_lines()
            seg = (off[o0:o1] - off[o0])
            if need("o_orderstatus"):
                nf = np.add.reduceat((li["l_linestatus"].data == 1).astype(np.int64), seg)
                cnt = nl[o0:o1]
                st = np.where(nf == cnt, 0, np.where(nf == 0, 1, 2)).astype(np.uint8)
                out["o_orderstatus"] = Column("o_orderstatus", "code", st, OSTATUS, 1)
            if need("o_totalprice"):
                v = np.round(li["l_extendedprice"].data * (1 + li["l_tax"].data) * (1 - li["l_discount"].data) * 100)
                tot = np.add.reduceat(v.astype(np.int64), seg)
                out["o_totalprice"] = Column("o_totalprice", "f64", tot / 100.0)
        if need("o_comment"):
            m = _text(S_ORD_CMT, i, 6, 79, seed)
            k = (mix(S_ORD_CMTK, i + 7777777, seed) % U64(100)).astype(np.int64)
            r0 = np.nonzero(k == 0)[0]  # special .. requests  (excluded by Q13)
            _put(m, r0, 10, "special ")
            _put(m, r0, 40, "requests ")
            r1 = np.nonzero(k == 1)[0]  # only 'special'
            _put(m, r1, 20, "special ")
            r2 = np.nonzero(k == 2)[0]  # 'requests' before 'special'
            _put(m, r2, 0, "requests ")
            _put(m, r2, 30, "special ")
            r3 = np.nonzero(k == 3)[0]  # too close: requests starts at special+6 exactly -> not excluded
            _put(m, r3, 10, "specialrequests")
            m = _cut(m, S_ORD_CMTK, i, 19, 70, seed, k < 4, 60)
            out["o_comment"] = Column("o_comment", "bytes", m, None, 79)
        return out

    def _gen_customer(self, want):
        i = np.arange(self.C)
        seed = self.seed
        out = {}

        def need(c):
            return want is None or c in want

        nat = rint(S_C_NAT, i, 0, 24, seed)
        if need("c_custkey"):
            out["c_custkey"] = Column("c_custkey", "i32", (i + 1).astype(np.int32))
        if need("c_name"):
            out["c_name"] = Column("c_name", "bytes", _prefixed_number("Customer#", i + 1, 9), None, 25)
        if need("c_address"):
            out["c_address"] = Column("c_address", "bytes", _cut(_text(S_C_TXT, i, 3, 40, seed), S_C_TXT, i + 99, 10, 30, seed), None, 40)
        if need("c_nationkey"):
            out["c_nationkey"] = Column("c_nationkey", "i32", nat.astype(np.int32))
        if need("c_phone"):
            out["c_phone"] = Column("c_phone", "bytes", _phone(S_C_PHONE, i, nat, seed), None, 15)
        if need("c_acctbal"):
            out["c_acctbal"] = Column("c_acctbal", "f64", rint(S_C_BAL, i, -99999, 999999, seed) / 100.0)
        if need("c_mktsegment"):
            out["c_mktsegment"] = Column("c_mktsegment", "code", rint(S_C_SEG, i, 0, 4, seed).astype(np.uint8), SEGMENTS, 10)
        if need("c_comment"):
            out["c_comment"] = Column("c_comment", "bytes", _cut(_text(S_C_TXT, i + (1 << 40), 9, 117, seed), S_C_TXT, i + 55, 29, 108, seed), None, 117)
        return out

    def _gen_supplier(self, want):
        i = np.arange(self.S)
        seed = self.seed
        out = {}

        def need(c):
            return want is None or c in want

        nat = rint(S_S_NAT, i, 0, 24, seed)
        if need("s_suppkey"):
            out["s_suppkey"] = Column("s_suppkey", "i32", (i + 1).astype(np.int32))
        if need("s_name"):
            out["s_name"] = Column("s_name", "bytes", _prefixed_number("Supplier#", i + 1, 9), None, 25)
        if need("s_address"):
            out["s_address"] = Column("s_address", "bytes", _cut(_text(S_S_TXT, i, 3, 40, seed), S_S_TXT, i + 99, 10, 30, seed), None, 40)
        if need("s_nationkey"):
            out["s_nationkey"] = Column("s_nationkey", "i32", nat.astype(np.int32))
        if need("s_phone"):
            out["s_phone"] = Column("s_phone", "bytes", _phone(S_S_PHONE, i, nat, seed), None, 15)
        if need("s_acctbal"):
            out["s_acctbal"] = Column("s_acctbal", "f64", rint(S_S_BAL, i, -99999, 999999, seed) / 100.0)
        if need("s_comment"):
            m = _text(S_S_CMT, i, 8, 101, seed)
            k = (mix(S_S_CMTK, i + 7777777, seed) % U64(2000)).astype(np.int64)
            r0 = np.nonzero(k == 0)[0]  # Customer .. Complaints (Q16 excludes these suppliers)
            _put(m, r0, 10, "Customer ")
            _put(m, r0, 40, "Complaints ")
            r1 = np.nonzero(k == 1)[0]  # Customer .. Recommends
            _put(m, r1, 10, "Customer ")
            _put(m, r1, 40, "Recommends ")
            r2 = np.nonzero(k == 2)[0]  # Complaints before Customer
            _put(m, r2, 0, "Complaints ")
            _put(m, r2, 30, "Customer ")
            m = _cut(m, S_S_CMTK, i, 25, 95, seed, k < 3, 60)
            out["s_comment"] = Column("s_comment", "bytes", m, None, 101)
        return out

    def _gen_part(self, want):
        i = np.arange(self.P)
        pk = i + 1
        seed = self.seed
        out = {}

        def need(c):
            return want is None or c in want

        if need("p_partkey"):
            out["p_partkey"] = Column("p_partkey", "i32", pk.astype(np.int32))
        if need("p_name"):
            # 5 colour words (as dbgen); <= 54 chars so the reference's wcsstr stays inside the row
            tab = _bytes_table(COLORS, 10)
            lens = np.array([len(c) for c in COLORS])
            m = np.zeros((self.P, 55), dtype=np.uint8)
            pos = np.zeros(self.P, dtype=np.int64)
            rows = np.arange(self.P)
            for w in range(5):
                code = (mix(S_P_NAME, i * 5 + w, seed) % U64(92)).astype(np.int64)
                for k in range(10):
                    ch = tab[code, k]
                    ok = k < lens[code]
                    m[rows[ok], (pos + k)[ok]] = ch[ok]
                pos = pos + lens[code]
                if w < 4:
                    m[rows, pos] = 32
                    pos = pos + 1
            out["p_name"] = Column("p_name", "bytes", m, None, 55)
        mf = rint(S_P_MFGR, i, 1, 5, seed)
        if need("p_mfgr"):
            out["p_mfgr"] = Column("p_mfgr", "code", (mf - 1).astype(np.uint8), MFGRS, 25)
        if need("p_brand"):
            bn = rint(S_P_BRAND, i, 1, 5, seed)
            out["p_brand"] = Column("p_brand", "code", ((mf - 1) * 5 + bn - 1).astype(np.uint8), BRANDS, 10)
        if need("p_type"):
            out["p_type"] = Column("p_type", "code", rint(S_P_TYPE, i, 0, 149, seed).astype(np.uint8), TYPES, 25)
        if need("p_size"):
            out["p_size"] = Column("p_size", "i32", rint(S_P_SIZE, i, 1, 50, seed).astype(np.int32))
        if need("p_container"):
            out["p_container"] = Column("p_container", "code", rint(S_P_CONT, i, 0, 39, seed).astype(np.uint8), CONTAINERS, 10)
        if need("p_retailprice"):
            out["p_retailprice"] = Column("p_retailprice", "f64", retail_cents(pk) / 100.0)
        return out

    def _gen_partsupp(self, want):
        r = np.arange(4 * self.P)
        pk = r // 4 + 1
        j = r % 4
        seed = self.seed
        out = {}

        def need(c):
            return want is None or c in want

        if need("ps_partkey"):
            out["ps_partkey"] = Column("ps_partkey", "i32", pk.astype(np.int32))
        if need("ps_suppkey"):
            sk = (pk + j * (self.S // 4 + (pk - 1) // self.S)) % self.S + 1
            out["ps_suppkey"] = Column("ps_suppkey", "i32", sk.astype(np.int32))
        if need("ps_availqty"):
            out["ps_availqty"] = Column("ps_availqty", "f64", rint(S_PS_QTY, r, 1, 9999, seed).astype(np.float64))
        if need("ps_supplycost"):
            out["ps_supplycost"] = Column("ps_supplycost", "f64", rint(S_PS_COST, r, 100, 100000, seed) / 100.0)
        return out

    def _gen_nation(self, want):
        out = {"n_nationkey": Column("n_nationkey", "i32", np.arange(25, dtype=np.int32)),
               "n_name": Column("n_name", "code", np.arange(25, dtype=np.uint8), [n for n, _ in NATIONS], 25),
               "n_regionkey": Column("n_regionkey", "i32", np.array([r for _, r in NATIONS], dtype=np.int32))}
        return {k: v for k, v in out.items() if want is None or k in want}

    def _gen_region(self, want):
        out = {"r_regionkey": Column("r_regionkey", "i32", np.arange(5, dtype=np.int32)),
               "r_name": Column("r_name", "code", np.arange(5, dtype=np.uint8), REGIONS, 25)}
        return {k: v for k, v in out.items() if want is None or k in want}

    # -- reference layout ----------------------------------------------------------------------
    def ref_table(self, table, cols, order_range=None):
        """list of numpy arrays in schema order (sdql_lib.py:115 ``data``); columns not in ``cols`` are
        1-element placeholders (the reference never dereferences columns a query does not use; the row
        count is taken from column 0, sdql_compiler.py:644), so column 0 is always materialised."""
        schema = SCHEMAS[table]
        want = set(cols) | {schema[0][0]}
        gen = self.columns(table, want, order_range)
        out = []
        for name, kind in schema:
            if name in gen:
                out.append(np.ascontiguousarray(gen[name].to_ref()))
            elif isinstance(kind, tuple):
                out.append(np.zeros(1, dtype="<U%d" % kind[1]))
            elif kind == "float":
                out.append(np.zeros(1, dtype=np.float64))
            else:
                out.append(np.zeros(1, dtype=np.int64))
        return out
