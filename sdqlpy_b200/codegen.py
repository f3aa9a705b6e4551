"""SDQL IR -> sm_100a CUDA.   Replaces the reference's C++/TBB generator
(/root/reference/src/sdqlpy/lib/sdql_ir_cpp_generator_par.py:12-637, ``GenerateCPPCode``).

The reference lowers every ``SumExpr`` over a relation to one of five TBB loop templates (gen:188-568) that fill
thread-local phmap tables and merge them serially.  Here every sum becomes ONE fused CUDA kernel
(scan + predicates + probes + aggregation inlined, the operator fusion SDQL exists for), assembled from the
hand-written device runtime in csrc/sdqlb200_rt.cuh:

  reference template (gen: lines)                     this generator
  ------------------------------------------------    -------------------------------------------------------------
  dense-array sum              gen:191-224, 472-495   BuildSink/GroupSink on a direct-indexed Tbl (domain from column
                                                      statistics at run time; the literal N of dense(N, ..) is ignored)
  scalar / record reduction    gen:258-291            ReduceSink: register accumulators -> block tree -> last-block tree
  dictionary build ("unique")  gen:293-369            BuildSink: key -> representative row id (CAS claim); payload
                                                      fields are re-evaluated from the representative row on use
  dictionary aggregation       gen:402-440            GroupSink: 3 tiers picked at run time -- thread-private shared
                                                      memory, CTA-shared shared memory atomics, global atomics
  local / finalisation sums    gen:520-568            the same kernels over table slots; result sets -> ResultSink
  probes  .contains()/.at()    gen:85-96              tbl_find on packed keys (mixed radix from column statistics)

The lowering is a symbolic evaluation of the IR: expressions evaluate to ``S*`` values carrying C++ text plus
provenance, which is what allows group keys to be minimised by functional dependency (e.g. Q3's
<l_orderkey, o_orderdate, o_shippriority> is keyed by l_orderkey alone) and wide payloads (strings) to stay
on the host until the result is boxed (late materialisation by row id).
"""
import json
import os
from collections import OrderedDict

from . import ir
from .ir import CompareSymbol as CS
from .ir import ExtFuncSymbol as XF


class CodegenError(Exception):
    pass


class _SplitDone(Exception):
    """unwinds run_body after the kernel body was split into a filter phase and a compacted phase"""


# scan-loop latency hiding:
#   "tma" = bulk-async (TMA engine) copies of whole column tiles into a multi-stage shared-memory ring tracked by
#           mbarriers: bytes in flight = stages x tile, no registers held by outstanding loads
#   "reg" = 128-bit LDGs, the next row group double-buffered in registers
#   "l2"  = 128-bit LDGs, single register buffer plus prefetch.global.L2 of the group PF_DIST iterations ahead
PIPELINE = os.environ.get("SDQLB200_PIPELINE", "auto")  # auto | tma | reg | l2 (see Kernel.pipe_mode)
PF_DIST = int(os.environ.get("SDQLB200_PF_DIST", "2"))
AUTO_PIPE = os.environ.get("SDQLB200_AUTO_PIPE", "reg")  # what "auto" picks for tiered (group-by) kernels
RING_ROWS_PER_THREAD = int(os.environ.get("SDQLB200_RING_ROWS", "0"))  # 0 = per kernel (Kernel.ring_rows)
RING_MAX_ROW_BYTES = 100     # wider scans cannot keep two stages in 227 KB of shared memory: they use LDGs
# string columns of the scanned row staged through shared memory: measured on B200 (SF10) Q13 2.25 ms staged vs 1.80 ms
# with plain (L1-cached) byte loads and the same look-ahead search, Q16/Q2 within noise -> opt-in
BYTE_STAGING = os.environ.get("SDQLB200_BYTE_STAGING", "0") == "1"
# firstIndex / contains: word-wise search with a two-character prefix filter (sdqlrt::str_find); 0 = byte loop
STRFIND_W = os.environ.get("SDQLB200_STRFIND_W", "1") == "1"
# narrow scans with 8 / 16 rows per thread per iteration (groups of 4 rows one CTA-width apart, so loads stay coalesced):
# lost the B200 A/B at SF10 on 17 of 20 queries (profiles/r01_codegen_variants_ab2.json: Q16 0.47 vs 0.29 ms, Q5 0.67 vs
# 0.56 ms), so 4 rows per thread stays the default
ROWS_AUTO = os.environ.get("SDQLB200_ROWS_AUTO", "0") == "1"
# firstIndex / contains on a scanned string column: the warp first streams its 128 rows' bytes with coalesced 128-bit loads
# and marks the rows that contain the pattern's first four characters anywhere (sdqlrt::warp_text_scan); the exact
# per-row search then runs for those candidate rows only
TEXTSCAN = os.environ.get("SDQLB200_TEXTSCAN", "1") != "0"
# ... and the exact firstIndex of the candidate rows is computed by the whole warp per (row, pattern)
# (sdqlrt::warp_text_resolve) instead of by the lane that owns the row; 0 = per-lane str_find on candidate rows
TEXTRESOLVE = os.environ.get("SDQLB200_TEXTRESOLVE", "1") != "0"
# patterns of >= 7 characters: the scan compares the run's aligned 32-bit words with the pattern's four 4-character
# substrings (sdqlrt::warp_text_scan_aligned) instead of every byte window with the first four characters
# (B200, SF100: q13_k0 10.23 ms per-lane search -> 9.58 ms warp resolve -> 7.80 ms + aligned words, profiles/r02_visit7)
TEXTALIGNED = os.environ.get("SDQLB200_TEXTALIGNED", "1") != "0"
# Every scan-loop iteration ends with a full-warp sync (all lanes of a warp run the same number of iterations).  Lanes that
# take a data-dependent slow path (an insertion with its probe loop, a hit behind a probe) otherwise do not rejoin their
# warp: ncu showed q12_k0's main loop executing with 13 of 32 lanes active and 2.5x the warp-level instructions of the
# cardinality pass over the same rows (profiles/r01_q12_k0_main_*.txt; SF100: 15.0 ms vs 1.3 ms).
RECONVERGE = os.environ.get("SDQLB200_RECONVERGE", "1") != "0"
# probes of single-part tables with an int32 column value use 32-bit key arithmetic (sdqlrt::pack_key1 / tbl_find1): the
# narrow lineitem scans are bound by instruction issue (~3 warp instructions per row, profiles/r01_q5_k5_narrow_*.txt),
# most of them 64-bit key packing and the generic presence test
PROBE32 = os.environ.get("SDQLB200_PROBE32", "1") != "0"
# 32-bit row / group indices in the relation-scan loops (row ids are int32 everywhere else already: Tbl.rep, the hit
# queues); the host driver refuses relations of more than 2'000'000'000 rows.  B200 A/B on identical data
# (profiles/r02_visit1/r02_ab_candidates_sf10.json, _sf100.log): Q5 -6 %, Q9 -6 %, Q10 -5 %, Q19 -7 %, Q20 -10 % at SF10, Q5
# 3.98 -> 3.76 ms and Q10 5.22 -> 5.03 ms at SF100; Q17 +8 % at SF10 is the one regression.  SDQLB200_IDX32=0 restores
# 64-bit indices (relations beyond 2^31 rows).
IDX32 = os.environ.get("SDQLB200_IDX32", "1") == "1"
# Tier 0 of a group-by (tiny key domain) keeps the thread-private accumulators in SHARED memory: cell (slot, field) of
# thread t lives at sm[(slot * nf + field) * kBlock + t] (conflict-free: consecutive lanes, consecutive 8-byte words).  A
# row updates only the nf cells of its slot (LDS + DADD + STS each) instead of running rcap x nf predicated adds, and the
# kernel needs ~60 fewer registers -- which is what lets the scan loop double-buffer the next row group in registers
# ("reg" pipeline).  B200 (profiles/r02_visit2): q1_k0 SF10 0.455 -> 0.373 ms (6.1 TB/s, 0.93 of the measured copy peak),
# SF100 4.19 -> 3.42 ms (6.67 TB/s); with register accumulators the same loop needs 172 registers (1 CTA per SM).
# SDQLB200_TIER0_SMEM=0 restores register accumulators.
TIER0_SMEM = os.environ.get("SDQLB200_TIER0_SMEM", "1") == "1"


COMPACT = os.environ.get("SDQLB200_COMPACT", "1") != "0"  # rows surviving a selective probe are queued and re-dealt to all lanes
BITS_FILTER = os.environ.get("SDQLB200_BITS", "1") != "0"  # presence bitmaps in front of selective, probed tables
COUNT_PASS = os.environ.get("SDQLB200_COUNT_PASS", "1") != "0"  # cardinality passes in front of selective table builds


# =============================================================================================
# symbolic values
# =============================================================================================
E = frozenset()


class SScalar:
    def __init__(self, ctype, code, prov=E, det=E, stats=None):
        self.ctype, self.code, self.prov, self.det, self.stats = ctype, code, prov, det, stats


class SStr:
    """kind: 'const' (value) | 'ref' (arg, col, row) | 'codeval' (arg, col, code) | 'pack' (n, code)"""

    def __init__(self, kind, **kw):
        self.kind = kind
        self.prov, self.det = kw.pop("prov", E), kw.pop("det", E)
        self.__dict__.update(kw)


class SRec:
    def __init__(self, fields):
        self.fields = OrderedDict(fields)

    def field(self, K, name):
        if name not in self.fields:
            raise CodegenError("record has no field '%s' (has %s)" % (name, list(self.fields)))
        v = self.fields[name]
        return v() if callable(v) else v

    def items(self, K):
        return [(n, self.field(K, n)) for n in self.fields]


class SRow:
    """record = row ``row`` of relation argument ``arg``."""

    def __init__(self, q, arg, row, scan=False, prov=E, keycode=None, token=None, gather=False):
        self.q, self.arg, self.row, self.scan, self.prov, self.keycode, self.token = q, arg, row, scan, prov, keycode, token
        self.gather = gather  # the scanned row seen from another lane (hit compaction): values are loaded by row id

    def field(self, K, name):
        return self.q.col_value(K, self, name)

    def items(self, K):
        return [(n, self.field(K, n)) for n, _ in self.q.schemas[self.arg]]


class SPair:
    def __init__(self, k, v):
        self.k, self.v = k, v


class SNone:
    pass


class STable:
    def __init__(self, desc):
        self.desc = desc


class SDictLit:
    def __init__(self, k, v):
        self.k, self.v = k, v


class SVecLit:
    def __init__(self, elem):
        self.elem = elem


class SLookup:
    def __init__(self, K, table, slot, token):
        self.K, self.table, self.slot, self.token = K, table, slot, token
        self.found = "(%s >= 0)" % slot
        self._val = None

    def value(self):
        if self._val is None:
            self._val = self.table.value_at(self.K, self.slot, self.token, safe=True)
        return self._val


TRUE = SScalar("bool", "true")
CT = {"i64": "long long", "f64": "double", "bool": "bool"}


def lit_f64(v):
    r = repr(float(v))
    if "e" not in r and "." not in r and "inf" not in r and "nan" not in r:
        r += ".0"
    return r


def cstr(s):
    return '"' + s.replace("\\", "\\\\").replace('"', '\\"') + '"'


# =============================================================================================
# tables
# =============================================================================================
class TableDesc:
    """compile-time description of one device dictionary."""

    def __init__(self, q, name, kind):
        self.q, self.name, self.kind = q, name, kind  # kind: 'build' | 'agg'
        self.parts = []         # by-value key parts: stats tuples
        self.fields = []        # aggregate fields [(name or None, ctype)]
        self.scalar_value = False
        self.count_only = False
        self.src = None         # ('rel', arg) | ('tbl', TableDesc)
        self.key_fn = None      # (K, idx_code) -> key SValue   (rep-evaluation)
        self.val_fn = None      # build: (K, idx_code) -> value SValue
        self.inner = None       # nested dict value: (n_outer_parts, [inner stats], inner key template SValue)
        self.probed = False     # looked up by some kernel (tbl_find): candidates for a presence filter
        self.distinct_of = None

    # -- element access ---------------------------------------------------------------------
    def src_elem(self, K, idx, prov=E):
        """the (key, value) pair of this table's *source* at source index idx."""
        if self.src[0] == "rel":
            return SPair(SRow(self.q, self.src[1], idx, scan=False, prov=prov), TRUE)
        t = self.src[1]
        return SPair(t.key_at(K, idx, prov), t.value_at(K, idx, None, prov=prov))

    def key_at(self, K, slot, prov=E):
        """key of the entry in ``slot``: parts kept by value are decoded from the packed key, functionally
        dependent parts are re-evaluated (lazily) at the slot's representative source index."""
        rp_code = "sdqlrt::rep_of(c.%s, %s)" % (self.name, slot)

        def rep_leaf(i):
            def f():
                rp = K.let("int", rp_code)
                return flatten(K, self.key_fn(K, rp, prov))[i]
            return f

        shape = getattr(self, "key_shape", None)
        if shape is None:  # never keyed through setup_key
            rp = K.let("int", rp_code)
            return self.key_fn(K, rp, prov)
        kk = K.let("unsigned long long", "sdqlrt::tbl_key(c.%s, %s)" % (self.name, slot))
        leaves = []
        for i in range(self.full_arity):
            if i in self.kept_pos:
                j = self.kept_pos.index(i)
                code = "sdqlrt::unpack_part(%s, c.%s_mn[%d], c.%s_rng[%d], c.%s_mul[%d], c.%s_sb[%d], c.%s_sk[%d])" % (
                    kk, self.name, j, self.name, j, self.name, j, self.name, j, self.name, j)
                leaves.append(decode_leaf(self.leaf_kinds[j], code, prov, self.parts[j]))
            else:
                leaves.append(rep_leaf(i))
        if shape == "scalar":
            v = leaves[0]
            return v() if callable(v) else v
        return SRec(list(zip(shape, leaves)))

    def value_at(self, K, slot, token, safe=False, prov=E):
        q = self.q
        if token is not None:
            prov = frozenset([token])
        sl = "(%s < 0 ? 0 : %s)" % (slot, slot) if safe else slot
        if self.kind == "agg":
            vals = []
            for j, (fname, ct) in enumerate(self.fields):
                ld = "ld1" if slot == "i" else "ldg1"  # sequential when the table itself is being scanned
                # an int64 aggregate used as a key part elsewhere: its value range is known once this table is complete
                vals.append((fname, SScalar(ct, "sdqlrt::%s(c.%s_a%d + %s)" % (ld, self.name, j, sl), prov=prov,
                                            stats=("agg", self.name, j) if ct == "i64" else None)))
            if self.inner is not None:
                raise CodegenError("nested dictionary value used as a plain value")
            if self.scalar_value or self.count_only:
                return vals[0][1]
            return SRec(vals)
        rp = K.let("int", "sdqlrt::rep_of(c.%s, %s)" % (self.name, slot))
        keycode = None
        if token is not None:
            # a probe reads this entry's payload through its representative source row: only valid where that row is
            # local (after a cross-GPU merge the entries of other ranks have no local representative, merge_code)
            self.rep_probed = True
            kv = flatten(K, self.key_fn(K, rp, prov))
            if len(kv) == 1 and hasattr(kv[0], "code"):
                keycode = kv[0].code
            elif len(kv) == 1 and kv[0].kind == "ref":
                keycode = ("ref", kv[0].arg, kv[0].col, kv[0].row)
        v = self.val_fn(K, rp, prov)
        return mark(K, v, prov, keycode, token)


def leaf_kind(x):
    if isinstance(x, SScalar):
        return (x.ctype,)
    if x.kind in ("ref", "codeval"):
        return ("codeval", x.arg, x.col)
    if x.kind == "pack":
        return ("pack", x.n)
    raise CodegenError("unsupported key part")


def decode_leaf(kind, code, prov, stats):
    if kind[0] == "i64":
        return SScalar("i64", code, prov, E, stats)
    if kind[0] == "f64":
        return SScalar("f64", "__longlong_as_double(%s)" % code, prov)
    if kind[0] == "bool":
        return SScalar("bool", "(%s != 0)" % code, prov)
    if kind[0] == "codeval":
        return SStr("codeval", arg=kind[1], col=kind[2], code=code, prov=prov)
    return SStr("pack", n=kind[1], code=code, prov=prov)


def mark(K, v, prov, keycode, token):
    """stamp provenance on everything reachable from a lookup result (lazily for records)."""
    if isinstance(v, SScalar):
        det = frozenset([token]) if (token is not None and keycode is not None and v.code == keycode) else v.det
        return SScalar(v.ctype, v.code, v.prov if v.prov else prov, det, v.stats)
    if isinstance(v, SStr):
        if v.kind == "const":
            return v
        d = dict(v.__dict__)
        kind = d.pop("kind")
        d["prov"] = v.prov if v.prov else prov
        if token is not None and keycode == ("ref", d.get("arg"), d.get("col"), d.get("row")):
            d["det"] = frozenset([token])
        return SStr(kind, **d)
    if isinstance(v, SRec):
        def lazy(n):
            return lambda: mark(K, v.field(K, n), prov, keycode, token)
        return SRec([(n, lazy(n)) for n in v.fields])
    if isinstance(v, SRow):
        return SRow(v.q, v.arg, v.row, False, v.prov if v.prov else prov, keycode, token)
    if isinstance(v, SLookup):
        return mark(K, v.value(), prov, keycode, token)
    return v


def flatten(K, v):
    """leaf values of a key in field order."""
    if isinstance(v, (SScalar, SStr)):
        return [v]
    if isinstance(v, SLookup):
        return flatten(K, v.value())
    if isinstance(v, (SRec, SRow)):
        out = []
        for _, f in v.items(K):
            out += flatten(K, f)
        return out
    raise CodegenError("cannot use %s as a dictionary key" % type(v).__name__)


# =============================================================================================
# kernels
# =============================================================================================
class Kernel:
    def __init__(self, q, name, src):
        self.q, self.name, self.src = q, name, src
        self.body, self.pre, self.post = [], [], []
        self.depth = 0
        self.scan_cols = OrderedDict()   # (col, rep) -> (array name, input idx)
        self.cse = [{}]
        self.ntmp = 0
        self.sink = None
        self.tiered = False
        self.smem_expr = "0"
        self.tier_expr = "2"
        self.scan_var = "i"
        self.body2 = None         # hit compaction: the full body re-evaluated for a queued row id `iq` (all lanes busy)
        self.nprobe_sel = 0       # selective probes evaluated so far (lookups into tables built behind predicates)
        self.byte_cols = OrderedDict()  # input idx -> width: fixed-width string columns staged through shared memory
        self.text_cols = OrderedDict()  # input idx -> [width, [patterns]]: scanned string columns with a warp text scan
        self.iter_pre, self.iter_post = [], []  # code in front of / behind the unrolled row loop of one iteration
        self.counted = False      # has a cardinality-pass variant (TIER == 3): predicates in front of a table build
        self.pred_cols = None     # scan columns the predicates read (the only ones the cardinality pass loads)

    @property
    def templated(self):
        return self.tiered or self.counted

    def enable_count(self):
        """cardinality pass (TIER == 3): the same scan + predicates + probes, but every row that reaches the sink is only
        counted.  The host sizes the kernel's tables from that count instead of the source's row count, so selective
        builds get small (cache-resident) tables and cheap initialisation."""
        if self.counted:
            return
        self.counted = True
        self.count_slot = self.q.ntcount
        self.q.ntcount += 1
        self.pre.append("unsigned long long cnt_rows = 0;")
        self.post.append("if (TIER == 3) {")
        self.post.append("    const long long s_ = sdqlrt::block_sum((long long)cnt_rows);")
        self.post.append("    if (threadIdx.x == 0 && s_) atomicAdd(c.tcount + %d, (unsigned long long)s_);" % self.count_slot)
        self.post.append("    return;")
        self.post.append("}")

    def count_guard(self, col, rep):
        """loads of columns that only the sink reads are skipped by the cardinality pass"""
        if self.counted and self.pred_cols is not None and (col, rep) not in self.pred_cols:
            return "if (TIER != 3) "
        return ""

    def emit(self, s):
        self.body.append("    " * self.depth + s)

    def ring_rows(self):
        """rows per thread per tile (tile = kBlock x rows): wide scans use smaller tiles so that two CTAs per SM can
        each keep several stages in flight and the per-thread row buffer stays small."""
        if RING_ROWS_PER_THREAD:
            return RING_ROWS_PER_THREAD
        row_bytes = sum({"i32": 4, "f64": 8, "code": 1}[rep] for (_, rep) in self.scan_cols)
        return 4 if row_bytes <= 24 else 2

    def pipe_mode(self):
        if self.src[0] != "rel" or not self.scan_cols:
            return None
        pipe = PIPELINE
        if pipe == "auto":
            # measured on B200 (profiles/r01_pipeline_ab.json, profiles/r02_visit2): the LDG pipelines beat the TMA ring on
            # every kernel class.  Group-by kernels keep their tier-0 accumulators in shared memory, so they can afford the
            # register double buffer (q1_k0: 120 registers, 2 CTAs per SM, 6.1-6.7 TB/s against 5.2 TB/s with the L2-prefetch
            # loop); wide scans without a tier use the single buffer + L2 prefetch.  SDQLB200_AUTO_PIPE=l2: the round-1 rule.
            if AUTO_PIPE == "l2" or self.body2 is not None:
                # kernels with hit compaction (filter phase + compacted phase) keep the single buffer: Q3 / Q7 / Q9 / Q10 / Q20
                # are 3-8 % faster with it at SF10 (profiles/r02_visit4/r02_ab_pipes_sf10.json)
                pipe = "l2" if (self.tiered or len(self.scan_cols) >= 6) else "reg"
            else:
                pipe = "l2" if (not self.tiered and len(self.scan_cols) >= 6) else "reg"
        if pipe == "tma" and sum({"i32": 4, "f64": 8, "code": 4}[rep] for (_, rep) in self.scan_cols) > RING_MAX_ROW_BYTES:
            pipe = "l2"
        return pipe

    def tmp(self, p="t"):
        self.ntmp += 1
        return "%s%d" % (p, self.ntmp)

    def let(self, ctype, expr):
        key = (ctype, expr)
        for scope in self.cse:
            if key in scope:
                return scope[key]
        v = self.tmp()
        self.emit("const %s %s = %s;" % (ctype, v, expr))
        self.cse[-1][key] = v
        return v

    def open_if(self, cond):
        self.emit("if (%s) {" % cond)
        self.depth += 1
        self.cse.append({})

    def open_block(self, head):
        self.emit(head + " {")
        self.depth += 1
        self.cse.append({})

    def close(self):
        self.cse.pop()
        self.depth -= 1
        self.emit("}")

    def rows_per_thread(self):
        """rows one thread handles per iteration (a multiple of 4: whole 128-bit loads).  Narrow scans (the filter phase
        of a compacted kernel streams a key column or two) take 8 or 16 so that enough bytes are in flight per thread;
        only where the per-row code is short (it is unrolled)."""
        if self.src[0] != "rel" or self.byte_cols or self.text_cols or not ROWS_AUTO:
            return 4
        if self.body2 is None and len(self.body) > 12:
            return 4
        b = sum({"i32": 4, "f64": 8, "code": 1}[rep] for (_, rep) in self.scan_cols)
        return 16 if b <= 4 else 8 if b <= 12 else 4

    def scan_col(self, col, rep):
        k = (col, rep)
        if k not in self.scan_cols:
            idx = self.q.input(self.src[1], col, rep)
            self.scan_cols[k] = ("r_%s%s" % (col, "_c" if rep == "code" else ""), idx)
        return self.scan_cols[k][0] + "[u]"

    # -- text -------------------------------------------------------------------------------
    def render(self):
        q = self.q
        L = []
        tmpl = "template <int TIER>\n" if self.templated else ""
        L.append("%s__global__ void __launch_bounds__(sdqlrt::kBlock) %s(const __grid_constant__ %s_ctx c) {" %
                 (tmpl, self.name, q.name))
        if self.tiered or self.pipe_mode() == "tma" or self.byte_cols or self.text_cols or self.body2 is not None:
            L.append("    SDQL_EXTERN_SMEM(sm);")
        L += ["    " + s for s in self.pre]
        if self.src[0] == "rel" and self.pipe_mode() == "tma":
            L += self.render_ring()
        elif self.src[0] == "rel":
            # columns the evaluator touched but whose values no emitted line reads (payload fields of a build whose value is
            # re-evaluated at the representative row later, vector elements that only count): not loaded at all
            # (q21_k3 / q21_k4 streamed 2.4 GB of l_suppkey each for nothing)
            text_ = "\n".join(self.pre + self.iter_pre + self.body + (self.body2 or []) + self.iter_post + self.post)
            for key_ in [k_ for k_, (arr_, _) in self.scan_cols.items() if (arr_ + "[") not in text_]:
                del self.scan_cols[key_]
            # software-pipelined streaming loop: the next group's column loads are issued before the current
            # group is processed, so every thread keeps two groups (2 x 4 rows x all columns) in flight
            ety = {"i32": "int", "f64": "double", "code": "int"}
            IT = "int" if IDX32 else "long long"  # type of row / row-group indices in this loop
            R = self.rows_per_thread()
            self.R = R

            G = R // 4  # groups of 4 consecutive rows per thread per iteration, blockDim.x groups apart: every 128-bit
            #             load instruction of a warp stays one contiguous run of bytes

            def loads(prefix, gvar, ind):
                o = []
                for k in range(G):
                    o.append(ind + "{")
                    o.append(ind + "    const %s j0 = (%s + %d * (%s)blockDim.x) << 2;" % (IT, gvar, k, IT))
                    o.append(ind + "    if (j0 + 4 <= n) {")
                    for (col, rep), (arr, idx) in self.scan_cols.items():
                        dst = prefix + arr[2:]
                        cg = self.count_guard(col, rep)
                        d4 = "reinterpret_cast<%s(&)[4]>(%s[%d])" % (ety[rep], dst, 4 * k)
                        if rep == "code":
                            o.append(ind + "        %ssdqlrt::ld4_code(c.in%d, j0, c.in%d_w, %s);" % (cg, idx, idx, d4))
                        else:
                            o.append(ind + "        %ssdqlrt::ld4(c.in%d + j0, %s);" % (cg, idx, d4))
                    o.append(ind + "    } else {")
                    o.append(ind + "        for (int u = 0; u < 4; ++u) {")
                    o.append(ind + "            const %s ii = (j0 + u < n) ? j0 + u : n - 1;" % IT)
                    for (col, rep), (arr, idx) in self.scan_cols.items():
                        dst = prefix + arr[2:]
                        cg = self.count_guard(col, rep)
                        if rep == "code":
                            o.append(ind + "            %s%s[%d + u] = sdqlrt::ld1_code(c.in%d, ii, c.in%d_w);" % (cg, dst, 4 * k, idx, idx))
                        else:
                            o.append(ind + "            %s%s[%d + u] = sdqlrt::ld1(c.in%d + ii);" % (cg, dst, 4 * k, idx))
                    o.append(ind + "        }")
                    o.append(ind + "    }")
                    o.append(ind + "}")
                return o

            L.append("    const %s n = %sc.n_%s;" % (IT, "(int)" if IDX32 else "", self.src[1]))
            L.append("    const %s ngrp = (n + 3) >> 2;" % IT)
            L.append("    const %s gstride = (%s)gridDim.x * blockDim.x * %d;" % (IT, IT, G))
            L.append("    %s g = (%s)blockIdx.x * blockDim.x * %d + threadIdx.x;" % (IT, IT, G))
            loop_cond = "g < ngrp"
            if RECONVERGE:
                loop_cond = "g - (%s)(threadIdx.x & (sdqlrt::kLanes - 1)) < ngrp" % IT
            if self.body2 is not None:
                loop_cond = "g - (%s)(threadIdx.x & (sdqlrt::kLanes - 1)) < ngrp" % IT
                L.append("    int* const wq = (int*)((unsigned char*)sm + c.%s_qo) + (threadIdx.x / sdqlrt::kLanes) * (sdqlrt::kLanes * %d);" %
                         (self.name, R + 1))
                L.append("    const int lane_ = (int)(threadIdx.x & (sdqlrt::kLanes - 1));")
                L.append("    int wcnt = 0;  // rows queued by this warp (the same value in every lane)")
            if self.byte_cols:
                # every lane of a warp runs the same number of iterations (the warp stages its rows cooperatively)
                loop_cond = "g - (%s)(threadIdx.x & (sdqlrt::kLanes - 1)) < ngrp" % IT
                boff = 0
                for idx, w in self.byte_cols.items():
                    L.append("    unsigned char* const bs%d = (unsigned char*)sm + c.%s_bo + %du + (threadIdx.x / sdqlrt::kLanes) * %du;" %
                             (idx, self.name, 16 + boff, 128 * w))
                    boff += 1024 * w
            row0 = "(g - (long long)(threadIdx.x & (sdqlrt::kLanes - 1))) << 2"  # first row of the warp's run (64-bit)
            stage = ["        sdqlrt::stage_rows(bs%d, c.in%d, %s, n, %d);" % (idx, idx, row0, w)
                     for idx, w in self.byte_cols.items()]
            if self.text_cols:
                # every lane of a warp runs the same number of iterations (the warp scans its rows' bytes cooperatively)
                loop_cond = "g - (%s)(threadIdx.x & (sdqlrt::kLanes - 1)) < ngrp" % IT
                toff = 0
                for idx, (w, pats) in self.text_cols.items():
                    L.append("    unsigned* const tm%d = (unsigned*)((unsigned char*)sm + c.%s_to + %du) + (threadIdx.x / sdqlrt::kLanes) * (%d * sdqlrt::kTextWords);" %
                             (idx, self.name, toff, len(pats)))
                    toff += 256 * 4 * len(pats)  # (kBlock / kLanes) warps x kTextWords words: at most 256 words per pattern
                    if TEXTALIGNED and all(len(p_) >= 7 for p_ in pats):
                        L.append("    const unsigned tp%d[%d][4] = {%s};" % (idx, len(pats), ", ".join(
                            "{%s}" % ", ".join("0x%08xu" % int.from_bytes(p_[o:o + 4].encode("latin1"), "little") for o in range(4))
                            for p_ in pats)))
                        stage.append("        sdqlrt::warp_text_scan_aligned<%d>(c.in%d, %s, n, %d, tp%d, tm%d);" %
                                     (len(pats), idx, row0, w, idx, idx))
                    else:
                        L.append("    const unsigned tp%d[%d] = {%s};" % (idx, len(pats), ", ".join(
                            "0x%08xu /* %s */" % (int.from_bytes(p_[:4].encode("latin1"), "little"), p_[:4].replace("*/", "")) for p_ in pats)))
                        stage.append("        sdqlrt::warp_text_scan<%d>(c.in%d, %s, n, %d, tp%d, tm%d);" %
                                     (len(pats), idx, row0, w, idx, idx))
                    if TEXTRESOLVE:
                        # exact positions: kStageRows int16 per warp and pattern behind the masks
                        L.append("    short* const tx%d = (short*)((unsigned char*)sm + c.%s_to + %du) + (threadIdx.x / sdqlrt::kLanes) * (%d * sdqlrt::kStageRows);" %
                                 (idx, self.name, toff, len(pats)))
                        L.append("    const sdqlrt::TextPat tq%d[%d] = {%s};" % (idx, len(pats), ", ".join(text_pat_init(p_) for p_ in pats)))
                        toff += 8 * 128 * 2 * len(pats)
                        stage.append("        sdqlrt::warp_text_resolve<%d>(c.in%d, %s, n, %d, tq%d, tm%d, tx%d);" %
                                     (len(pats), idx, row0, w, idx, idx, idx))

            def loop_head(cond):
                if self.body2 is None:
                    return ["    while (%s) {" % cond]
                # compaction: the drain below must also run once after the last scan iteration (one copy of the body)
                return ["    for (;;) {", "    const bool more_ = %s;" % cond, "    if (more_) {"]

            def prefetches():
                o = ["        { const %s gp = g + %d * gstride; if (gp < ngrp) {" % (IT, PF_DIST)]
                for (col, rep), (arr, idx) in self.scan_cols.items():
                    cg = self.count_guard(col, rep)
                    for k in range(G):
                        if rep == "code":
                            o.append("            %ssdqlrt::prefetch_l2((const char*)c.in%d + %s((gp + %d * (%s)blockDim.x) << 2) * c.in%d_w);" % (cg, idx, "(long long)" if IDX32 else "", k, IT, idx))
                        else:
                            o.append("            %ssdqlrt::prefetch_l2(c.in%d + ((gp + %d * (%s)blockDim.x) << 2));" % (cg, idx, k, IT))
                o.append("        } }")
                return o

            pipe = self.pipe_mode() or "reg"
            if pipe == "reg":
                for (col, rep), (arr, idx) in self.scan_cols.items():
                    L.append("    %s %s[%d], %s[%d];" % (ety[rep], arr, R, "q_" + arr[2:], R))
                L.append("    if (g < ngrp)")
                L += loads("r_", "g", "    ")
                L += loop_head(loop_cond)
                L.append("        const %s gn = g + gstride;" % IT)
                L.append("        if (gn < ngrp)")
                L += loads("q_", "gn", "        ")
                L += stage
            else:  # "l2": single register buffer, the group PF_DIST iterations ahead is prefetched into L2
                for (col, rep), (arr, idx) in self.scan_cols.items():
                    L.append("    %s %s[%d];" % (ety[rep], arr, R))
                L += loop_head(loop_cond)
                L.append("        const %s gn = g + gstride;" % IT)
                L += stage
                L += prefetches()
                L += loads("r_", "g", "        ")
            L += ["        " + x for x in self.iter_pre]
            if self.body2 is not None:
                L.append("        unsigned pm_ = 0u;  // bit u: row u of this lane survived the filter")
            L.append("#pragma unroll")
            L.append("        for (int u = 0; u < %d; ++u) {" % R)
            L.append("            const %s i = ((g + (u >> 2) * (%s)blockDim.x) << 2) + (u & 3);" % (IT, IT))
            if self.body2 is not None:
                L.append("            bool pass_ = false;")
            L.append("            if (i < n) {")
            L += ["                " + x for x in self.body]
            L.append("            }")
            if self.body2 is not None:
                L.append("            pm_ |= (pass_ ? 1u : 0u) << u;")
            L.append("        }")
            if self.body2 is not None:
                # warp-aggregated push of the surviving row ids.  One vote decides whether any of the warp's 32 x R rows
                # survived: behind a selective filter (Q17: 0.1 %) nearly every step skips the R ballots / popcounts / stores
                L.append("        if (sdqlrt::warp_ballot(pm_ != 0u)) {")
                L.append("#pragma unroll")
                L.append("            for (int u = 0; u < %d; ++u) {" % R)
                L.append("                const %s i = ((g + (u >> 2) * (%s)blockDim.x) << 2) + (u & 3);" % (IT, IT))
                L.append("                const bool p_ = (pm_ >> u) & 1u;")
                L.append("                const unsigned m_ = sdqlrt::warp_ballot(p_);")
                L.append("                if (p_) wq[wcnt + __popc(m_ & ((1u << lane_) - 1u))] = (int)i;")
                L.append("                wcnt += __popc(m_);")
                L.append("            }")
                L.append("        }")
            L += ["        " + x for x in self.iter_post]
            if RECONVERGE:
                L.append("        sdqlrt::warp_sync();  // lanes that took a slow path rejoin the warp here")
            L.append("        g = gn;")
            if self.scan_cols and pipe == "reg":
                L.append("#pragma unroll")
                L.append("        for (int u = 0; u < %d; ++u) {" % R)
                for (col, rep), (arr, idx) in self.scan_cols.items():
                    L.append("            %s[u] = %s[u];" % (arr, "q_" + arr[2:]))
                L.append("        }")
            L.append("    }")
            if self.body2 is not None:
                # deal the queued rows to the lanes: full groups of 32 while the scan runs, the remainder after it
                L.append("    for (;;) {")
                L.append("        const int take_ = wcnt >= sdqlrt::kLanes ? sdqlrt::kLanes : (more_ ? 0 : wcnt);")
                L.append("        if (take_ == 0) break;")
                L.append("        wcnt -= take_;")
                L.append("        sdqlrt::warp_sync();")
                L.append("        const bool act_ = lane_ < take_;")
                L.append("        const long long iq = act_ ? (long long)wq[wcnt + lane_] : 0ll;")
                L.append("        sdqlrt::warp_sync();")
                L.append("        if (act_) {")
                L += ["            " + x for x in self.body2]
                L.append("        }")
                L.append("    }")
                L.append("    if (!more_) break;")
                L.append("    }")
        elif self.src[0] == "tbl":
            t = self.src[1]
            L.append("    const long long n = c.%s.cap;" % t.name)
            L.append("    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; "
                     "i += (long long)gridDim.x * blockDim.x) {")
            if iterates_without_rep(self):
                # entries owned by another rank (rep == -2 after a cross-GPU merge) carry the merged aggregates and a key that
                # is decoded from the slot alone: every rank can emit them, the result needs no concatenation (c.<t>_all)
                L.append("        { const int r_ = c.%s.rep[i]; if (r_ == -1 || (r_ < 0 && !c.%s_all)) continue; }" % (t.name, t.name))
            else:
                L.append("        if (c.%s.rep[i] < 0) continue;" % t.name)
            L += ["        " + s for s in self.body]
            L.append("    }")
        else:  # single-thread finalisation kernel
            L.append("    if (blockIdx.x == 0 && threadIdx.x == 0) {")
            L += ["        " + s for s in self.body]
            L.append("    }")
        L += ["    " + s for s in self.post]
        L.append("}")
        return "\n".join(L)


def _render_ring(self):
    """rel-scan loop over a shared-memory ring of column tiles filled by bulk-async (TMA) copies.

    tile = blockDim.x * R rows; thread t of the CTA owns rows t, t + blockDim.x, ... of a tile (conflict-free shared
    memory reads).  CTA b processes tiles b, b + gridDim.x, ...; thread 0 keeps S tiles in flight: it arms full[s]
    with the stage's byte count, issues one bulk copy per scanned column, and refills a stage as soon as every warp
    has arrived on empty[s] (the warps arrive right after copying their rows to registers, before computing).
    The last, partial tile of the relation is read with plain guarded loads.  Under SDQLB200_EMU (tests) every tile
    takes that path."""
    R = self.ring_rows()
    ety = {"i32": "int", "f64": "double", "code": "int"}
    cols = list(self.scan_cols.items())
    L = []
    L.append("    const long long n = c.n_%s;" % self.src[1])
    L.append("    const int TILE = (int)blockDim.x * %d;" % R)
    L.append("    const long long ntile = (n + TILE - 1) / TILE;")
    for (col, rep), (arr, idx) in cols:
        L.append("    %s %s[%d];" % (ety[rep], arr, R))
    L.append("#ifndef SDQLB200_EMU")
    L.append("    const long long nfull = n / TILE;")
    L.append("    unsigned char* const ring = (unsigned char*)sm + c.%s_ro;" % self.name)
    L.append("    const int S = c.%s_rs;" % self.name)
    L.append("    unsigned long long* const full = (unsigned long long*)ring;")
    L.append("    unsigned long long* const empty = full + 8;")
    L.append("    unsigned char* const stage0 = ring + 128;")
    off = "0u"
    for j, ((col, rep), (arr, idx)) in enumerate(cols):
        w = {"i32": "4u", "f64": "8u", "code": "(unsigned)c.in%d_w" % idx}[rep]
        L.append("    const unsigned o%d = %s, b%d = %s * (unsigned)TILE;" % (j, off, j, w))
        off = "o%d + b%d" % (j, j)
    L.append("    const unsigned sbytes = %s;" % off)
    L.append("    const unsigned sstride = (sbytes + 127u) & ~127u;")
    L.append("    auto issue = [&](int s, long long k) {")
    L.append("        unsigned char* const st = stage0 + (size_t)s * sstride;")
    L.append("        sdqlrt::mbar_expect_tx(full + s, sbytes);")
    for j, ((col, rep), (arr, idx)) in enumerate(cols):
        if rep == "code":
            L.append("        sdqlrt::bulk_g2s(st + o%d, (const unsigned char*)c.in%d + k * TILE * c.in%d_w, b%d, full + s);" % (j, idx, idx, j))
        else:
            L.append("        sdqlrt::bulk_g2s(st + o%d, c.in%d + k * TILE, b%d, full + s);" % (j, idx, j))
    L.append("    };")
    L.append("    if (threadIdx.x == 0) {")
    L.append("        for (int s = 0; s < S; ++s) { sdqlrt::mbar_init(full + s, 1u); sdqlrt::mbar_init(empty + s, blockDim.x >> 5); }")
    L.append("        sdqlrt::mbar_fence_init();")
    L.append("    }")
    L.append("    __syncthreads();")
    L.append("    if (threadIdx.x == 0)")
    L.append("        for (int s = 0; s < S; ++s) { const long long k = blockIdx.x + (long long)s * gridDim.x; if (k < nfull) issue(s, k); }")
    L.append("    int rs = 0; unsigned rph = 0;  // ring stage / phase parity of the current tile")
    L.append("#else")
    L.append("    const long long nfull = 0;")
    L.append("#endif")
    L.append("    for (long long k = blockIdx.x; k < ntile; k += gridDim.x) {")
    L.append("        const long long i0 = k * TILE + threadIdx.x;")
    L.append("        if (k < nfull) {")
    L.append("#ifndef SDQLB200_EMU")
    L.append("            sdqlrt::mbar_wait(full + rs, rph);")
    L.append("            const unsigned char* const st = stage0 + (size_t)rs * sstride;")
    L.append("#pragma unroll")
    L.append("            for (int u = 0; u < %d; ++u) {" % R)
    L.append("                const int e = u * (int)blockDim.x + (int)threadIdx.x;")
    for j, ((col, rep), (arr, idx)) in enumerate(cols):
        if rep == "code":
            L.append("                %s[u] = c.in%d_w == 1 ? (int)(st + o%d)[e] : ((const int*)(st + o%d))[e];" % (arr, idx, j, j))
        else:
            L.append("                %s[u] = ((const %s*)(st + o%d))[e];" % (arr, ety[rep], j))
    L.append("            }")
    # The stage may be refilled by the TMA engine as soon as every warp has arrived on empty[rs], so the arrive must
    # not be issued before this warp's shared-memory loads have RETURNED (an LDS can sit behind global atomics /
    # probes in the load-store queue for a long time).  Folding every loaded word into the arrive's operand makes the
    # arrive wait on the loads' scoreboard.
    L.append("            unsigned dep = 0;")
    L.append("#pragma unroll")
    L.append("            for (int u = 0; u < %d; ++u) {" % R)
    for j, ((col, rep), (arr, idx)) in enumerate(cols):
        if rep == "f64":
            L.append("                dep ^= (unsigned)__double2loint(%s[u]) ^ (unsigned)__double2hiint(%s[u]);" % (arr, arr))
        else:
            L.append("                dep ^= (unsigned)%s[u];" % arr)
    L.append("            }")
    L.append("            dep = __reduce_or_sync(0xffffffffu, dep);  // all lanes' loads have returned")
    L.append("            // (S >> 8) is 0 at run time (S <= 8) but not to the compiler: the address really depends on dep")
    L.append("            if ((threadIdx.x & 31) == 0) sdqlrt::mbar_arrive(empty + rs + (dep & (unsigned)(S >> 8)));")
    L.append("            if (threadIdx.x == 0) {")
    L.append("                const long long kn = k + (long long)S * gridDim.x;")
    L.append("                if (kn < nfull) { sdqlrt::mbar_wait(empty + rs, rph); issue(rs, kn); }")
    L.append("            }")
    L.append("            if (++rs == S) { rs = 0; rph ^= 1u; }")
    L.append("#endif")
    L.append("        } else {")
    L.append("            for (int u = 0; u < %d; ++u) {" % R)
    L.append("                const long long iu = i0 + (long long)u * blockDim.x;")
    L.append("                const long long ii = iu < n ? iu : n - 1;")
    for (col, rep), (arr, idx) in cols:
        if rep == "code":
            L.append("                %s[u] = sdqlrt::ld1_code(c.in%d, ii, c.in%d_w);" % (arr, idx, idx))
        else:
            L.append("                %s[u] = sdqlrt::ld1(c.in%d + ii);" % (arr, idx))
    L.append("            }")
    L.append("        }")
    L.append("#pragma unroll")
    L.append("        for (int u = 0; u < %d; ++u) {" % R)
    L.append("            const long long i = i0 + (long long)u * blockDim.x;")
    L.append("            if (i < n) {")
    L += ["                " + x for x in self.body]
    L.append("            }")
    L.append("        }")
    L.append("    }")
    return L


Kernel.render_ring = _render_ring


def text_pat_init(pat):
    """C++ initialiser of sdqlrt::TextPat: the pattern as four masked little-endian words + its length"""
    b = pat.encode("latin1")[:16]
    ws, ms = [], []
    for k in range(4):
        chunk = b[4 * k:4 * k + 4]
        ws.append(int.from_bytes(chunk, "little"))
        ms.append(int.from_bytes(b"\xff" * len(chunk), "little"))
    return "{{%s}, {%s}, %d}" % (", ".join("0x%08xu" % x for x in ws), ", ".join("0x%08xu" % x for x in ms), len(pat))


def iterates_without_rep(K):
    """does kernel K, which iterates a table, get by without the representative source row of the entries it visits (no
    functionally dependent key part / late-materialised payload re-evaluated there)?  Then any rank can visit any entry."""
    if K.src[0] != "tbl":
        return False
    needle = "rep_of(c.%s, i)" % K.src[1].name
    return not any(needle in line for line in K.body)


# =============================================================================================
# sinks
# =============================================================================================
def cast_to(v, ctype):
    if v.ctype == ctype:
        return v.code
    return "(%s)(%s)" % (CT[ctype], v.code)


class ReduceSink:
    """A2: scalar / record reduction (gen:258-291)."""

    def __init__(self, q, K, first):
        self.q, self.K = q, K
        self.fields = []     # (name, ctype, slot)
        self.sub = {}        # field name -> GroupSink (dictionary-valued record field, Q11)
        self.scalar = not isinstance(first, SRec)

    def _acc(self, name, v):
        for f in self.fields:
            if f[0] == name:
                return f
        slot = self.q.new_scalar()
        f = (name, v.ctype if v.ctype != "bool" else "i64", slot, "acc%d" % len(self.fields))
        self.fields.append(f)
        self.K.pre.append("%s %s = 0;" % (CT[f[1]], f[3]))
        return f

    def produce(self, v):
        K = self.K
        items = [(None, v)] if self.scalar else v.items(K)
        for name, x in items:
            if isinstance(x, SLookup):
                x = x.value()
            if isinstance(x, SDictLit):
                if name not in self.sub:
                    self.sub[name] = GroupSink(self.q, K, self.q.new_table("agg"), False)
                self.sub[name].produce(x)
                continue
            if not isinstance(x, SScalar):
                raise CodegenError("cannot reduce a %s" % type(x).__name__)
            f = self._acc(name, x)
            K.emit("%s += %s;" % (f[3], cast_to(x, f[1])))

    def finish(self):
        q, K = self.q, self.K
        nf = len(self.fields)
        if nf:
            part = q.new_partials(nf)
            cnt = q.new_counter()
            K.post.append("{")
            for j, f in enumerate(self.fields):
                K.post.append("    %s b%d = sdqlrt::block_sum(%s);" % (CT[f[1]], j, f[3]))
                K.post.append("    if (threadIdx.x == 0) c.part[%s + %d * (long long)gridDim.x + blockIdx.x] = %s;" %
                              (part, j, "b%d" % j if f[1] == "f64" else "__longlong_as_double(b%d)" % j))
            K.post.append("    if (sdqlrt::last_block(c.cnt + %d)) {" % cnt)
            for j, f in enumerate(self.fields):
                if f[1] == "f64":
                    K.post.append("        double s%d = 0; for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) "
                                  "s%d += c.part[%s + %d * (long long)gridDim.x + b];" % (j, j, part, j))
                    K.post.append("        s%d = sdqlrt::block_sum(s%d);" % (j, j))
                    K.post.append("        if (threadIdx.x == 0) c.sc[%d] = s%d;" % (f[2], j))
                else:
                    K.post.append("        long long s%d = 0; for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) "
                                  "s%d += __double_as_longlong(c.part[%s + %d * (long long)gridDim.x + b]);" % (j, j, part, j))
                    K.post.append("        s%d = sdqlrt::block_sum(s%d);" % (j, j))
                    K.post.append("        if (threadIdx.x == 0) c.sc[%d] = __longlong_as_double(s%d);" % (f[2], j))
            K.post.append("    }")
            K.post.append("}")
        for s in self.sub.values():
            s.finish()

        def dev(f):
            if f[1] == "f64":
                return SScalar("f64", "c.sc[%d]" % f[2])
            return SScalar("i64", "__double_as_longlong(c.sc[%d])" % f[2])
        if self.scalar:
            return dev(self.fields[0])
        out = [(f[0], dev(f)) for f in self.fields]
        out += [(n, STable(s.t)) for n, s in self.sub.items()]
        return SRec(out)


def unwrap(v):
    return v.value() if isinstance(v, SLookup) else v


def is_true(v):
    return isinstance(v, SScalar) and v.code == "true"


class KeyedSink:
    """shared by BuildSink / GroupSink: FD-minimised, mixed-radix packed keys."""

    def count_open(self):
        K = self.K
        if K.sink is self and K.counted:  # cardinality pass: the row is counted instead of inserted
            K.emit("if (TIER == 3) { ++cnt_rows; }")
            K.open_block("else")

    def count_close(self):
        K = self.K
        if K.sink is self and K.counted:
            K.close()

    def setup_key(self, t, keyval, extra_parts=()):
        q, K = self.q, self.K
        leaves = flatten(K, keyval)
        kept = minimise_key(q, leaves)
        if not kept:
            kept = leaves[:1]
        parts = [q.part_code(K, x) for x in kept] + list(extra_parts)
        if not t.parts:
            t.full_arity = len(leaves)
            t.kept_pos = [i for i, x in enumerate(leaves) if any(x is y for y in kept)]
            t.leaf_kinds = [leaf_kind(x) for x in kept]
            kv = unwrap(keyval)
            if isinstance(kv, (SRec, SRow)):
                names = [n for n, _ in kv.items(K)]
                t.key_shape = names if len(names) == len(leaves) else None
            else:
                t.key_shape = "scalar"
            t.parts = [p[1] for p in parts]
            if any(p[0] in ("raw", "agg", "union") for p in t.parts) and len(t.parts) > 1:
                raise CodegenError("%s: key with an unbounded part cannot be packed with other parts" % t.name)
        elif len(t.parts) != len(parts):
            raise CodegenError("%s: inconsistent key shapes" % t.name)
        return [p[0] for p in parts]

    def pack(self, t, codes):
        K = self.K
        kk = K.tmp("kk")
        K.emit("unsigned long long %s = 0; bool %s_ok = true;" % (kk, kk))
        for j, code in enumerate(codes):
            K.emit("%s_ok &= sdqlrt::pack_part(%s, c.%s_mn[%d], c.%s_rng[%d], c.%s_mul[%d], c.%s_sb[%d], c.%s_sk[%d], %s);" %
                   (kk, code, t.name, j, t.name, j, t.name, j, t.name, j, t.name, j, kk))
        return kk


class BuildSink(KeyedSink):
    """A1/A3: assignment ("unique") dictionary: key -> representative source index."""

    def __init__(self, q, K, t):
        self.q, self.K, self.t = q, K, t
        t.src = K.src

    def produce(self, d):
        K, t = self.K, self.t
        self.count_open()
        kk = self.pack(t, self.setup_key(t, d.k))
        K.emit("if (%s_ok) { bool nw; sdqlrt::tbl_upsert(c.%s, %s, (int)%s, nw); }" % (kk, t.name, kk, K.scan_var))
        self.count_close()

    def finish(self):
        return STable(self.t)


class GroupSink(KeyedSink):
    """A4: aggregating dictionary (GROUP BY)."""

    def __init__(self, q, K, t, tiered=True):
        self.q, self.K, self.t = q, K, t
        t.src = K.src
        self.tiered = tiered
        self.tc = None

    def produce(self, d):
        self.count_open()
        self.produce_(d)
        self.count_close()

    def produce_(self, d):
        q, K, t = self.q, self.K, self.t
        v = unwrap(d.v)
        extra, inner_leaves = [], None
        if isinstance(v, SDictLit):  # nested dictionary value
            inner_leaves = flatten(K, v.k)
            extra = [q.part_code(K, x) for x in inner_leaves]
            iv = unwrap(v.v)
            if is_true(iv):
                return self.produce_distinct(d, extra)
            v = iv
        if isinstance(v, SVecLit):
            items, t.count_only = [(None, SScalar("i64", "1ll"))], True
        elif isinstance(v, SScalar):
            items, t.scalar_value = [(None, v)], True
        elif isinstance(v, (SRec, SRow)):
            items = [(n, unwrap(x)) for n, x in v.items(K)]
        else:
            raise CodegenError("cannot aggregate a %s" % type(v).__name__)
        codes = self.setup_key(t, d.k, extra)
        if inner_leaves is not None and t.inner is None:
            t.inner = (len(codes) - len(extra), [p[1] for p in extra], inner_leaves)
            self.tiered = False
        if not t.fields:
            for n, x in items:
                if not isinstance(x, SScalar):
                    raise CodegenError("cannot aggregate field %s of type %s" % (n, type(x).__name__))
                t.fields.append((n, "i64" if x.ctype == "bool" else x.ctype))
            if self.tiered:
                self._prologue()
        kk = self.pack(t, codes)
        K.open_if("%s_ok" % kk)
        nf = len(t.fields)
        vals = [cast_to(x, ct) for (n, x), (_, ct) in zip(items, t.fields)]
        if self.tiered:
            # TIER 0: accumulators of every (slot, field) cell live in registers; the row's slot is matched against the
            # statically unrolled slot index (predicated adds, no shared-memory traffic, no atomics)
            K.emit("if (TIER == 0) {")
            if TIER0_SMEM:
                K.emit("    unsigned long long* const pc_ = pa_ + (int)%s * (%d * sdqlrt::kBlock);" % (kk, nf))
                for j, (_, ct) in enumerate(t.fields):
                    if ct == "f64":
                        K.emit("    ((double*)pc_)[%d * sdqlrt::kBlock] += %s;" % (j, vals[j]))
                    else:
                        K.emit("    pc_[%d * sdqlrt::kBlock] += (unsigned long long)(%s);" % (j, vals[j]))
                K.emit("    pr_[(int)%s * sdqlrt::kBlock] = (int)%s;" % (kk, K.scan_var))
            for sl_ in range(self.rcap if not TIER0_SMEM else 0):  # unrolled here (scalar accumulators: nothing can end up in local memory)
                upd = " ".join(
                    "ra%d_%d += %s;" % (j, sl_, ("(unsigned long long)(%s)" % vals[j]) if ct != "f64" else vals[j])
                    for j, (_, ct) in enumerate(t.fields))
                K.emit("    if ((int)%s == %d) { %s rrep_%d = (int)%s; }" % (kk, sl_, upd, sl_, K.scan_var))
            K.emit("} else if (TIER == 1) {")
            K.emit("    const int sb = (int)%s * %d;" % (kk, nf))
            for j, (_, ct) in enumerate(t.fields):
                idx = "sb + %d" % j
                if ct == "f64":
                    K.emit("    atomicAdd((double*)&sm[%s], %s);" % (idx, vals[j]))
                else:
                    K.emit("    atomicAdd(&sm[%s], (unsigned long long)(%s));" % (idx, vals[j]))
            K.emit("    smrep[(int)%s] = (int)%s;" % (kk, K.scan_var))
            K.emit("} else {")
            K.depth += 1
        K.emit("bool nw; const int sl = sdqlrt::tbl_upsert(c.%s, %s, (int)%s, nw);" % (t.name, kk, K.scan_var))
        for j in range(nf):
            K.emit("sdqlrt::red_add(c.%s_a%d + sl, %s);" % (t.name, j, vals[j]))
        if self.tiered:
            K.depth -= 1
            K.emit("}")
        K.close()

    def produce_distinct(self, d, inner_parts):
        """{k: {x: True}} aggregated: a set of (k, x) plus a per-k counter bumped on first insertion (Q16 dictSize)."""
        q, K, t = self.q, self.K, self.t
        self.tiered = False
        if self.tc is None:
            self.tc = q.new_table("build")
            self.tc.src = K.src
            t.fields = [(None, "i64")]
            t.scalar_value = True
            t.distinct_of = self.tc
        tc = self.tc
        ccodes = self.setup_key(tc, d.k, inner_parts)
        ocodes = self.setup_key(t, d.k)
        kc = self.pack(tc, ccodes)
        K.open_if("%s_ok" % kc)
        K.emit("bool nwc; sdqlrt::tbl_upsert(c.%s, %s, (int)%s, nwc);" % (tc.name, kc, K.scan_var))
        K.open_if("nwc")
        ko = self.pack(t, ocodes)
        K.emit("bool nw; const int sl = sdqlrt::tbl_upsert(c.%s, %s, (int)%s, nw);" % (t.name, ko, K.scan_var))
        K.emit("sdqlrt::red_add(c.%s_a0 + sl, 1ll);" % t.name)
        K.close()
        K.close()

    def _prologue(self):
        K, t = self.K, self.t
        nf = len(t.fields)
        K.tiered = True
        self.rcap = max(1, min(8, 32 // nf))  # register tier: <= 8 groups and <= 32 accumulator cells per thread
        K.rcap = self.rcap
        K.pre.append("const long long ncap = c.%s.cap;" % t.name)
        if TIER0_SMEM:
            K.pre.append("unsigned long long* const pa_ = sm + threadIdx.x;  // this thread's cells: pa_[cell * kBlock]")
            K.pre.append("int* const pr_ = (int*)(sm + %d * sdqlrt::kBlock) + threadIdx.x;  // ... and its representative rows" % (self.rcap * nf))
            K.pre.append("if (TIER == 0) { for (int k = 0; k < %d; ++k) pa_[k * sdqlrt::kBlock] = 0; "
                         "for (int k = 0; k < %d; ++k) pr_[k * sdqlrt::kBlock] = -1; }" % (self.rcap * nf, self.rcap))
        else:
            for j, (_, ct) in enumerate(t.fields):
                K.pre.append("%s %s;" % ("double" if ct == "f64" else "unsigned long long",
                                         ", ".join("ra%d_%d = 0" % (j, sl_) for sl_ in range(self.rcap))))
        if not TIER0_SMEM:
            K.pre.append("int %s;" % ", ".join("rrep_%d = -1" % sl_ for sl_ in range(self.rcap)))
        K.pre.append("int* smrep = (int*)(sm + ncap * %d);" % nf)
        K.pre.append("if (TIER == 1) { for (long long k = threadIdx.x; k < ncap * %d; k += blockDim.x) sm[k] = 0; "
                     "for (long long k = threadIdx.x; k < ncap; k += blockDim.x) smrep[k] = -1; __syncthreads(); }" % nf)
        K.smem_nf = nf
        K.smem_tbl = t.name

    def finish(self):
        K, t = self.K, self.t
        if self.tiered and K.tiered and t.fields:
            nf = len(t.fields)
            P = K.post
            P.append("if (TIER == 0) {")
            for sl_ in range(self.rcap):
                P.append("    if (%d < ncap) {" % sl_)
                P.append("        const int r = sdqlrt::block_max(%s);" % (("pr_[%d * sdqlrt::kBlock]" % sl_) if TIER0_SMEM else "rrep_%d" % sl_))
                for j, (_, ct) in enumerate(t.fields):
                    acc = "ra%d_%d" % (j, sl_)
                    if TIER0_SMEM:
                        acc = ("((double*)pa_)[%d * sdqlrt::kBlock]" if ct == "f64" else "pa_[%d * sdqlrt::kBlock]") % (sl_ * nf + j)
                    if ct == "f64":
                        P.append("        const double v%d = sdqlrt::block_sum(%s);" % (j, acc))
                    else:
                        P.append("        const long long v%d = sdqlrt::block_sum((long long)%s);" % (j, acc))
                P.append("        if (threadIdx.x == 0 && r >= 0) {")
                P.append("            atomicMax(c.%s.rep + %d, r);" % (t.name, sl_))
                for j, (_, ct) in enumerate(t.fields):
                    P.append("            sdqlrt::red_add(c.%s_a%d + %d, v%d);" % (t.name, j, sl_, j))
                P.append("        }")
                P.append("    }")
            P.append("} else if (TIER == 1) {")
            P.append("    __syncthreads();")
            P.append("    for (long long k = threadIdx.x; k < ncap; k += blockDim.x) {")
            P.append("        int r = smrep[k];")
            P.append("        if (r >= 0) {")
            P.append("            atomicMax(c.%s.rep + k, r);" % t.name)
            for j, (_, ct) in enumerate(t.fields):
                if ct == "f64":
                    P.append("            sdqlrt::red_add(c.%s_a%d + k, __longlong_as_double(sm[k * %d + %d]));" % (t.name, j, nf, j))
                else:
                    P.append("            sdqlrt::red_add(c.%s_a%d + k, (long long)sm[k * %d + %d]);" % (t.name, j, nf, j))
            P.append("        }")
            P.append("    }")
            P.append("}")
        return STable(t)


class ResultSink:
    """final result set {record -> True}: append rows to SoA output columns (A6 + boundary, gen:871-877)."""

    def __init__(self, q, K):
        self.q, self.K = q, K

    def produce(self, d):
        q, K = self.q, self.K
        key = d.k if isinstance(d, SDictLit) else d
        if isinstance(key, (SScalar, SStr)):
            key = SRec([("_", key)])
        if isinstance(key, SLookup):
            key = key.value()
        items = key.items(K)
        schema, codes = [], []
        for n, x in items:
            if isinstance(x, SLookup):
                x = x.value()
            kind, code = q.result_field(K, x)
            schema.append((n, kind))
            codes.append(code)
        if q.result_schema is None:
            q.result_schema = schema
        elif [s[1] for s in q.result_schema] != [s[1] for s in schema]:
            raise CodegenError("result rows of different shapes")
        pos = K.tmp("pos")
        K.emit("const unsigned long long %s = sdqlrt::append_slot(c.res_count);" % pos)
        K.open_if("%s < (unsigned long long)c.res_cap" % pos)
        for j, code in enumerate(codes):
            K.emit("c.res%d[%s] = %s;" % (j, pos, code))
        K.close()

    def finish(self):
        return "RESULT"


def minimise_key(q, leaves):
    """drop key parts that are functionally determined by the remaining ones."""
    kept = [x for x in leaves if not (isinstance(x, SStr) and x.kind == "const")]

    def closure(parts):
        det = set()
        for p in parts:
            det |= set(p.det)
            if not p.prov or all(isinstance(t, tuple) and t and t[0] == "col" for t in p.prov):
                det |= set(p.prov)
        changed = True
        while changed:
            changed = False
            for tok, deps in q.token_deps.items():
                if tok not in det and deps is not None and all(d in det for d in deps):
                    det.add(tok)
                    changed = True
        return det

    changed = True
    while changed:
        changed = False
        for i, p in enumerate(kept):
            others = kept[:i] + kept[i + 1:]
            if not p.prov:
                continue
            if all(isinstance(t, tuple) and t and t[0] == "col" for t in p.prov):
                continue  # an atom of the scanned row: always by value
            if p.prov <= closure(others):
                kept = others
                changed = True
                break
    return kept


# =============================================================================================
# query compiler
# =============================================================================================
class Query:
    def __init__(self, name, args, schemas):
        self.name, self.args, self.schemas = name, args, schemas
        self.inputs, self.input_idx = [], {}
        self.consts, self.const_idx = [], {}
        self.tables, self.kernels, self.steps = [], [], []
        self.nsc = 0
        self.npart = []
        self.ncnt = 0
        self.ntcount = 0
        self.token_deps = {}
        self.ntok = 0
        self.result_schema = None
        self.result_kind = None
        self.res_cap_expr = "1"

    # -- resources --------------------------------------------------------------------------
    def input(self, arg, col, rep):
        k = (arg, col, rep)
        if k not in self.input_idx:
            self.input_idx[k] = len(self.inputs)
            self.inputs.append(k)
        return self.input_idx[k]

    def const_code(self, arg, col, lit):
        k = ("strcode", arg, col, lit)
        if k not in self.const_idx:
            self.const_idx[k] = len(self.consts)
            self.consts.append(k)
        return "c.k%d" % self.const_idx[k]

    def new_scalar(self):
        self.nsc += 1
        return self.nsc - 1

    def new_partials(self, nf):
        off = "c.part_off%d" % len(self.npart)
        self.npart.append(nf)
        return off

    def new_counter(self):
        self.ncnt += 1
        return self.ncnt - 1

    def new_table(self, kind):
        t = TableDesc(self, "t%d" % len(self.tables), kind)
        t.builder = None
        self.tables.append(t)
        return t

    def new_token(self, table, keyprov):
        self.ntok += 1
        tok = ("lk", table.name, self.ntok)
        self.token_deps[tok] = keyprov
        return tok

    def col_kind(self, arg, col):
        for n, k in self.schemas[arg]:
            if n == col:
                return k
        raise CodegenError("relation '%s' has no column '%s'" % (arg, col))

    # -- column access ------------------------------------------------------------------------
    def col_value(self, K, row, col):
        kind = self.col_kind(row.arg, col)
        on_scan = row.scan and K is not None and K.src == ("rel", row.arg)
        prov = row.prov if not on_scan else frozenset([("col", row.arg, col, "i")])
        if isinstance(kind, tuple):
            s = SStr("ref", arg=row.arg, col=col, row=row.row, scan=on_scan and not row.gather, width=kind[1], prov=prov)
            if row.keycode == ("ref", row.arg, col, row.row) and row.token is not None:
                s.det = frozenset([row.token])
            return s
        rep = "f64" if kind == "float" else "i32"
        if on_scan and not row.gather:
            code = K.scan_col(col, rep)
        else:
            code = "sdqlrt::ldg1(c.in%d + %s)" % (self.input(row.arg, col, rep), row.row)
        if rep == "i32":
            code = "(long long)" + code
        det = E
        if row.token is not None and row.keycode == code:
            det = frozenset([row.token])
        return SScalar("f64" if rep == "f64" else "i64", code, prov, det,
                       ("col", self.input(row.arg, col, rep)) if rep == "i32" else None)

    def str_code(self, K, s):
        if s.kind == "codeval":
            return s.code
        if s.kind == "ref":
            if s.scan:
                return K.scan_col(s.col, "code")
            idx = self.input(s.arg, s.col, "code")
            return "sdqlrt::ldg1_code(c.in%d, %s, c.in%d_w)" % (idx, s.row, idx)
        raise CodegenError("string of kind %s has no dictionary code" % s.kind)

    def str_ptr(self, K, s):
        if s.kind != "ref":
            raise CodegenError("pattern functions need a column string")
        idx = self.input(s.arg, s.col, "bytes")
        if s.scan and K is not None and K.src == ("rel", s.arg):
            if not hasattr(K, "str_scan"):
                K.str_scan = OrderedDict()
            K.str_scan[s.col] = s.width
        if s.scan and K is not None and K.src == ("rel", s.arg) and PIPELINE != "tma" and BYTE_STAGING:
            # the scanned row's bytes come from the warp's shared-memory staging buffer (Kernel.render / stage_rows)
            K.byte_cols[idx] = s.width
            return "(bs%d + (unsigned)(((threadIdx.x & (sdqlrt::kLanes - 1)) << 2) + u) * %du)" % (idx, s.width), s.width
        return "(c.in%d + (long long)(%s) * %d)" % (idx, s.row if not s.scan else "i", s.width), s.width

    def text_candidate(self, K, s, pattern):
        """firstIndex / contains of ``pattern`` in a string of the scanned row: register the pattern with the kernel's warp
        text scan and return the C++ test "this row may contain it" (None: no scan for this string)."""
        if not (TEXTSCAN and PIPELINE != "tma" and not BYTE_STAGING and isinstance(s, SStr) and s.kind == "ref" and s.scan
                and K is not None and K.src == ("rel", s.arg) and len(pattern) >= 4 and "\0" not in pattern):
            return None
        idx = self.input(s.arg, s.col, "bytes")
        if not hasattr(K, "str_scan"):
            K.str_scan = OrderedDict()
        K.str_scan[s.col] = s.width  # the manifest's scanned string bytes (roofline accounting)
        w, pats = K.text_cols.setdefault(idx, [s.width, []])
        if pattern not in pats:
            if len(pats) >= 8:
                return None
            pats.append(pattern)
        return "sdqlrt::text_cand(tm%d, %d, u & 3)" % (idx, pats.index(pattern))

    def text_position(self, K, s, pattern):
        """exact firstIndex of ``pattern`` in a string of the scanned row from the warp's cooperative search (None: not
        available -- no text scan for this string, or the pattern is longer than the 16 characters it handles)"""
        if not TEXTRESOLVE or len(pattern) > 16:
            return None
        cand = self.text_candidate(K, s, pattern)
        if cand is None:
            return None
        idx = self.input(s.arg, s.col, "bytes")
        return "sdqlrt::text_pos(tx%d, %d, u & 3)" % (idx, K.text_cols[idx][1].index(pattern))

    def part_code(self, K, x):
        """by-value key part -> (integer C++ expression, stats)."""
        if isinstance(x, SScalar):
            if x.ctype == "f64":
                return "__double_as_longlong(%s)" % x.code, ("raw",)
            if x.ctype == "bool":
                return "(long long)(%s)" % x.code, ("range", 0, 1)
            return x.code, (x.stats or ("raw",))
        if x.kind in ("ref", "codeval"):
            return "(long long)" + self.str_code(K, x), ("col", self.input(x.arg, x.col, "code"))
        if x.kind == "pack":
            return x.code, ("range", 0, 256 ** x.n - 1)
        raise CodegenError("unsupported key part")

    def result_field(self, K, x):
        if isinstance(x, SScalar):
            if x.ctype == "f64":
                return "f64", "__double_as_longlong(%s)" % x.code
            return ("bool" if x.ctype == "bool" else "i64"), "(long long)(%s)" % x.code
        if x.kind == "ref":
            self.host_strings.add((x.arg, x.col))
            return "str:ref:%s:%s" % (x.arg, x.col), "(long long)(%s)" % (x.row if not x.scan else "i")
        if x.kind == "codeval":
            return "str:code:%s:%s" % (x.arg, x.col), "(long long)(%s)" % x.code
        if x.kind == "pack":
            return "str:pack:%d" % x.n, x.code
        if x.kind == "const":
            return "str:const:" + x.value, "0ll"
        raise CodegenError("unsupported result field")

    # =========================================================================================
    # evaluation
    # =========================================================================================
    def compile(self, root):
        self.host_strings = set()
        env = {}
        for a in self.args:
            env[frontend_dataset(a)] = ("rel", a)
        self.top(root, env)
        return self

    def top(self, e, env):
        """host-level let chain."""
        while True:
            if not isinstance(e, ir.LetExpr):
                raise CodegenError("expected a let chain at the top level")
            if e.varExpr.name == "out":
                return self.finish_result(self.top_value(e.valExpr, env, True), env)
            is_res = (isinstance(e.bodyExpr, ir.LetExpr) and e.bodyExpr.varExpr.name == "out"
                      and isinstance(e.bodyExpr.valExpr, ir.VarExpr) and e.bodyExpr.valExpr.name == e.varExpr.name)
            env = dict(env)
            env[e.varExpr.name] = self.top_value(e.valExpr, env, is_res)
            e = e.bodyExpr

    def top_value(self, e, env, is_res=False):
        if isinstance(e, ir.SumExpr):
            return self.compile_sum(e, env, is_res)
        if isinstance(e, ir.LetExpr):
            env = dict(env)
            env[e.varExpr.name] = self.top_value(e.valExpr, env)
            return self.top_value(e.bodyExpr, env, is_res)
        return self.ev(e, env, None)

    def finish_result(self, v, env):
        if v == "RESULT":
            self.result_kind = "rows"
            return
        K = Kernel(self, "%s_fin" % self.name, ("one",))
        if isinstance(v, SScalar):
            self.result_kind = "f64" if v.ctype == "f64" else "i64"
            self.result_schema = [("_", self.result_kind)]
            K.emit("c.res0[0] = %s; *c.res_count = 1;" %
                   ("__double_as_longlong(%s)" % v.code if v.ctype == "f64" else "(long long)(%s)" % v.code))
            self.res_cap_expr = "1"
        elif isinstance(v, SDictLit):
            self.result_kind = "rows"
            ResultSink(self, K).produce(v)
            self.res_cap_expr = "1"
        elif isinstance(v, STable):
            # a dictionary that is returned as is: rows = key fields ++ value fields
            return self.materialise_table(v.desc)
        else:
            raise CodegenError("unsupported result value %s" % type(v).__name__)
        self.add_kernel(K)

    def materialise_table(self, t):
        K = Kernel(self, "%s_fin" % self.name, ("tbl", t))
        elem = SPair(t.key_at(K, "i"), t.value_at(K, "i", None))
        key = elem.k
        fields = []
        if isinstance(key, (SRec, SRow)):
            fields += key.items(K)
        else:
            fields.append(("key", key))
        if isinstance(elem.v, (SRec, SRow)):
            fields += elem.v.items(K)
        elif isinstance(elem.v, SScalar) and elem.v.code != "true":
            fields.append(("value", elem.v))
        self.result_kind = "rows"
        ResultSink(self, K).produce(SRec(fields))
        self.res_cap_expr = "c.%s.cap" % t.name
        self.add_kernel(K)

    def add_kernel(self, K):
        for t in self.tables:
            if t.builder is None and t.src is not None:
                t.builder = K  # tables are sized/created in order, a kernel owns the tables made since the last one
        self.kernels.append(K)
        self.steps.append(("launch", K))

    # -- sums -----------------------------------------------------------------------------------
    def compile_sum(self, S, env, is_res=False):
        src = self.ev(S.dictExpr, env, None)
        if isinstance(src, tuple) and src[0] == "rel":
            K = Kernel(self, "%s_k%d" % (self.name, len(self.kernels)), src)
            elem = SPair(SRow(self, src[1], "i", scan=True), TRUE)
            cap = "c.n_%s" % src[1]
        elif isinstance(src, STable):
            t = src.desc
            K = Kernel(self, "%s_k%d" % (self.name, len(self.kernels)), ("tbl", t))
            tok = ("it", K.name)
            self.token_deps[tok] = None
            elem = None
            cap = "c.%s.cap" % t.name
        else:
            raise CodegenError("cannot sum over %s" % type(src).__name__)
        K.S, K.is_res = S, is_res
        K.src_cap = cap
        if elem is None:
            t = src.desc
            key = mark_iter(self, K, t)
            elem = key
        env2 = dict(env)
        env2[S.varExpr.name] = elem
        K.root_body, K.root_env, K.row_var = S.bodyExpr, env2, S.varExpr.name
        try:
            self.run_body(S.bodyExpr, env2, K, chain=True)
        except _SplitDone:
            pass
        if K.sink is None:
            raise CodegenError("sum body never produces a value")
        out = K.sink.finish()
        if out == "RESULT":
            self.res_cap_expr = cap
        self.add_kernel(K)
        return out

    def make_sink(self, K, v):
        S = K.S
        if isinstance(v, SDictLit):
            val = v.v.value() if isinstance(v.v, SLookup) else v.v
            is_set = isinstance(val, SScalar) and val.code == "true"
            if K.is_res and is_set:
                return ResultSink(self, K)
            dense = S.dictType.startswith("dense_array")
            if isinstance(val, SVecLit) or isinstance(val, SDictLit):
                return GroupSink(self, K, self.new_table("agg"))
            if S.isAssignmentSum or (dense and not isinstance(val, SVecLit)):
                t = self.new_table("build")
                return BuildSink(self, K, t)
            return GroupSink(self, K, self.new_table("agg"))
        if isinstance(v, (SScalar, SRec)):
            return ReduceSink(self, K, v)
        raise CodegenError("sum body produces unsupported value %s" % type(v).__name__)

    def split_here(self, K):
        """hit compaction: everything emitted so far (scan-row predicates up to and including a selective probe) becomes
        the FILTER phase, which only queues the ids of surviving rows in the warp's shared-memory queue; the complete
        body is then re-evaluated for a queued row id with all 32 lanes busy (a few percent of the rows survive such a
        probe: without the queue the long dependent-load chains behind it run with one active lane per warp)."""
        K.emit("pass_ = true;")
        while K.depth > 0:
            K.close()
        body1, cse1 = K.body, K.cse
        K.body, K.cse, K.depth = [], [{}], 0
        K.scan_var = "iq"
        env = dict(K.root_env)
        env[K.row_var] = SPair(SRow(self, K.src[1], "iq", scan=True, gather=True), TRUE)
        self.run_body(K.root_body, env, K, chain=False)
        K.body2, K.body, K.cse = K.body, body1, cse1
        raise _SplitDone()

    def run_body(self, e, env, K, chain=False):
        if isinstance(e, ir.IfExpr):
            trivial_else = isinstance(e.elseBodyExpr, ir.EmptyDicConsExpr) or (
                isinstance(e.elseBodyExpr, ir.ConstantExpr) and e.elseBodyExpr.value in (None, 0, 0.0, False))
            if not trivial_else:
                raise CodegenError("if/else with a non-zero else branch at statement level is not supported")
            n = 0
            for conj in and_chain(e.condExpr):
                before = K.nprobe_sel
                c = self.ev(conj, env, K)
                K.open_if(as_bool(c))
                n += 1
                if (chain and COMPACT and K.nprobe_sel > before and K.body2 is None and K.src[0] == "rel"
                        and PIPELINE != "tma" and K.sink is None):
                    self.split_here(K)
            self.run_body(e.thenBodyExpr, env, K, chain=chain)
            for _ in range(n):
                K.close()
            return
        if isinstance(e, ir.SumExpr):  # nested sum over the inner dictionary of a lookup (Q12)
            return self.nested_sum(e, env, K)
        if isinstance(e, (ir.EmptyDicConsExpr,)) or (isinstance(e, ir.ConstantExpr) and e.value is None):
            return
        if K.sink is None and K.pred_cols is None:
            K.pred_cols = set(K.scan_cols)  # everything read so far was read by predicates / probes
        v = self.ev(e, env, K)
        if K.sink is None:
            K.sink = self.make_sink(K, v)
            if isinstance(K.sink, (BuildSink, GroupSink)):
                self.bind_table_fns(K, K.sink.t, e, env)
                if K.depth > 0 and K.src[0] in ("rel", "tbl") and COUNT_PASS:
                    K.enable_count()
        K.sink.produce(v)
        if isinstance(K.sink, ReduceSink) and isinstance(e, ir.RecConsExpr):
            for name, fe in e.initialPairs:
                sub = K.sink.sub.get(name)
                if sub is not None and sub.t.key_fn is None:
                    self.bind_table_fns(K, sub.t, fe, env)

    def bind_table_fns(self, K, t, e, env):
        """closures that re-evaluate the key / value expression of a produced {k: v} at a given source index."""
        S = K.S
        q = self
        var = S.varExpr.name
        if not isinstance(e, ir.DicConsExpr):
            return
        kexpr, vexpr = e.initialPairs[0]

        def elem_at(K2, idx, prov):
            if K.src[0] == "rel":
                return SPair(SRow(q, K.src[1], idx, scan=False, prov=prov), TRUE)
            st = K.src[1]
            return SPair(st.key_at(K2, idx, prov), st.value_at(K2, idx, None, prov=prov))

        def key_fn(K2, idx, prov=E):
            env2 = dict(env)
            env2[var] = elem_at(K2, idx, prov)
            return q.ev(kexpr, env2, K2)

        def val_fn(K2, idx, prov=E):
            env2 = dict(env)
            env2[var] = elem_at(K2, idx, prov)
            return q.ev(vexpr, env2, K2)

        t.key_fn, t.val_fn = key_fn, val_fn

    def nested_sum(self, S, env, K):
        d = self.ev(S.dictExpr, env, K)
        if not (isinstance(d, SLookup) and d.table.inner is not None):
            raise CodegenError("nested sum over something that is not a dictionary-valued lookup")
        t = d.table
        n_outer, inner_stats, inner_leaves = t.inner
        if len(inner_stats) != 1 or inner_stats[0][0] != "col":
            raise CodegenError("nested dictionaries need a single dictionary-coded inner key")
        idx = inner_stats[0][1]
        cv = K.tmp("cv")
        # a presence filter on the outer key part rejects the whole inner dictionary at once
        K.open_if("%s_outer_ok && (c.%s.bmod == 0 || sdqlrt::tbl_maybe(c.%s, %s_outer, %s_p0))" % (d.slot, t.name, t.name, d.slot, d.slot))
        K.open_block("for (long long %s = c.%s_mn[%d]; %s < c.%s_mn[%d] + c.%s_rng[%d]; ++%s)" %
                     (cv, t.name, n_outer, cv, t.name, n_outer, t.name, n_outer, cv))
        kk = K.tmp("kk")
        K.emit("unsigned long long %s = %s_outer + (unsigned long long)(%s - c.%s_mn[%d]) * (unsigned long long)c.%s_mul[%d];" %
               (kk, d.slot, cv, t.name, n_outer, t.name, n_outer))
        sl = K.let("int", "sdqlrt::tbl_find(c.%s, %s, %s_outer_ok, %s_p0)" % (t.name, kk, d.slot, d.slot))
        K.open_if("%s >= 0" % sl)
        leaf = inner_leaves[0]
        if isinstance(leaf, SStr):
            kval = SStr("codeval", arg=leaf.arg, col=leaf.col, code=cv)
        else:
            kval = SScalar("i64", cv, stats=inner_stats[0])
        vals = [SScalar(ct, "sdqlrt::ldg1(c.%s_a%d + %s)" % (t.name, j, sl)) for j, (_, ct) in enumerate(t.fields)]
        env2 = dict(env)
        env2[S.varExpr.name] = SPair(kval, vals[0] if (t.scalar_value or t.count_only) else
                                     SRec([(n, v) for (n, _), v in zip(t.fields, vals)]))
        self.run_body(S.bodyExpr, env2, K)
        K.close()
        K.close()
        K.close()

    # -- lookups ----------------------------------------------------------------------------------
    def lookup(self, K, t, keyval):
        if K is None:
            raise CodegenError("dictionary lookup outside of a sum body")
        t.probed = True
        if t.builder is not None and t.builder.counted:
            K.nprobe_sel += 1
        leaves = flatten(K, keyval)
        n_expected = len(t.parts) if t.inner is None else t.inner[0]
        # the probe key has the build key's *full* shape; keep the positions the build kept by value
        if len(leaves) != n_expected:
            full = getattr(t, "full_arity", None)
            if full is not None and len(leaves) == full:
                leaves = [leaves[i] for i in t.kept_pos]
            else:
                raise CodegenError("%s: lookup key has %d parts, table key has %d" % (t.name, len(leaves), n_expected))
        pcs = [self.part_code(K, x) for x in leaves]
        codes = [pc[0] for pc in pcs]
        ckey = ("lookup", t.name, tuple(codes))
        fast1 = (PROBE32 and len(codes) == 1 and t.inner is None and len(t.parts) == 1 and t.parts[0][0] == "col"
                 and pcs[0][1] and pcs[0][1][0] == "col")
        hit = None
        for scope in K.cse:
            if ckey in scope:
                hit = scope[ckey]
        if hit is None:
            kk = K.tmp("lk")
            K.emit("unsigned long long %s = 0; bool %s_ok = true;" % (kk, kk))
            if fast1:
                K.emit("%s_ok = sdqlrt::pack_key1((int)(%s), c.%s_mn[0], c.%s_rng[0], c.%s_sb[0], c.%s_sk[0], %s);" %
                       (kk, codes[0], t.name, t.name, t.name, t.name, kk))
            for j, code in enumerate(codes if not fast1 else ()):
                K.emit("%s_ok &= sdqlrt::pack_part(%s, c.%s_mn[%d], c.%s_rng[%d], c.%s_mul[%d], c.%s_sb[%d], c.%s_sk[%d], %s);" %
                       (kk, code, t.name, j, t.name, j, t.name, j, t.name, j, t.name, j, kk))
                if j == 0:
                    # the packed first part (mul == 1): what a first-part presence filter is indexed by -- passed along instead
                    # of being recovered as key % range (a 64-bit modulo: ~65 instructions, a quarter of q9_k5's)
                    K.emit("const unsigned long long %s_p0 = %s;" % (kk, kk))
            keyprov = frozenset().union(*[x.prov for x in leaves]) if leaves else E
            tok = self.new_token(t, keyprov if all(x.prov for x in leaves) else None)
            if t.inner is not None:
                K.emit("const unsigned long long %s_outer = %s; const bool %s_outer_ok = %s_ok;" % (kk, kk, kk, kk))
                sl = kk
            else:
                sl = K.tmp("sl")
                if fast1:
                    K.emit("const int %s = sdqlrt::tbl_find1(c.%s, (unsigned)%s, %s_ok);" % (sl, t.name, kk, kk))
                else:
                    K.emit("const int %s = sdqlrt::tbl_find(c.%s, %s, %s_ok, %s_p0);" % (sl, t.name, kk, kk, kk))
            hit = (sl, tok)
            K.cse[-1][ckey] = hit
        sl, tok = hit
        lk = SLookup(K, t, sl, tok)
        if t.inner is not None:
            # `d[k] != None` on a dictionary-valued entry: subsumed by iterating the inner dictionary, which
            # yields nothing when no (k, *) entry exists (the only use in the workload is joinProbe, Q12)
            # With a presence filter in front of the table the test is "some (k, *) may exist": it gates the whole inner
            # iteration (and is what the filter phase of a compacted kernel queues rows by); only its positive form is sound
            lk.found = "(%s_outer_ok && sdqlrt::tbl_maybe_outer(c.%s, %s_outer, c.%s_rng[%d], c.%s_mul[%d], %s_p0))" % (
                sl, t.name, sl, t.name, t.inner[0], t.name, t.inner[0], sl)
            lk.found_is_filter = True
        return lk

    # -- expressions ---------------------------------------------------------------------------
    def ev(self, e, env, K):
        m = getattr(self, "ev_" + type(e).__name__)
        return m(e, env, K)

    def ev_ConstantExpr(self, e, env, K):
        v = e.value
        if v is None:
            return SNone()
        if isinstance(v, bool):
            return SScalar("bool", "true" if v else "false")
        if isinstance(v, int):
            return SScalar("i64", "%dll" % v, stats=("range", v, v))
        if isinstance(v, float):
            return SScalar("f64", lit_f64(v))
        return SStr("const", value=v)

    def ev_VarExpr(self, e, env, K):
        if e.name not in env:
            raise CodegenError("unbound variable %s" % e.name)
        return env[e.name]

    def ev_LetExpr(self, e, env, K):
        env = dict(env)
        env[e.varExpr.name] = self.ev(e.valExpr, env, K) if K is not None else self.top_value(e.valExpr, env)
        return self.ev(e.bodyExpr, env, K)

    def ev_SumExpr(self, e, env, K):
        if K is None:
            return self.compile_sum(e, env)
        raise CodegenError("nested sum in expression position")

    def ev_PairAccessExpr(self, e, env, K):
        p = self.ev(e.pairExpr, env, K)
        if not isinstance(p, SPair):
            raise CodegenError("[%d] on a non-pair" % e.index)
        return p.k if e.index == 0 else p.v

    def ev_RecAccessExpr(self, e, env, K):
        r = self.ev(e.recExpr, env, K)
        if isinstance(r, SLookup):
            r = r.value()
        if isinstance(r, (SRec, SRow)):
            return r.field(K, e.name)
        raise CodegenError("field access .%s on %s" % (e.name, type(r).__name__))

    def ev_RecConsExpr(self, e, env, K):
        return SRec([(n, self.ev(x, env, K)) for n, x in e.initialPairs])

    def ev_ConcatExpr(self, e, env, K):
        a, b = self.ev(e.rec1, env, K), self.ev(e.rec2, env, K)
        out = []
        for r in (a, b):
            if isinstance(r, SLookup):
                r = r.value()
            if not isinstance(r, (SRec, SRow)):
                raise CodegenError("concat needs records")
            out += r.items(K)
        return SRec(out)

    def ev_DicConsExpr(self, e, env, K):
        k, v = e.initialPairs[0]
        return SDictLit(self.ev(k, env, K), self.ev(v, env, K))

    def ev_EmptyDicConsExpr(self, e, env, K):
        return SNone()

    def ev_VecConsExpr(self, e, env, K):
        return SVecLit(self.ev(e.exprList[0], env, K))

    def ev_DicLookupExpr(self, e, env, K):
        d = self.ev(e.dicExpr, env, K)
        if not isinstance(d, STable):
            raise CodegenError("lookup in something that is not a dictionary (%s)" % type(d).__name__)
        return self.lookup(K, d.desc, self.ev(e.keyExpr, env, K))

    def ev_IfExpr(self, e, env, K):
        c = self.ev(e.condExpr, env, K)
        a, b = self.ev(e.thenBodyExpr, env, K), self.ev(e.elseBodyExpr, env, K)
        a = a.value() if isinstance(a, SLookup) else a
        b = b.value() if isinstance(b, SLookup) else b
        if isinstance(a, SScalar) and isinstance(b, SScalar):
            ct = "f64" if "f64" in (a.ctype, b.ctype) else a.ctype
            st = ("union", a.stats, b.stats) if (ct == "i64" and a.stats and b.stats) else None
            return SScalar(ct, "(%s ? %s : %s)" % (as_bool(c), cast_to(a, ct), cast_to(b, ct)), a.prov | b.prov | c.prov, E, st)
        raise CodegenError("conditional expression over non-scalars")

    def _arith(self, e, env, K, op):
        a, b = self.ev(e.op1Expr, env, K), self.ev(e.op2Expr, env, K)
        a = a.value() if isinstance(a, SLookup) else a
        b = b.value() if isinstance(b, SLookup) else b
        if not (isinstance(a, SScalar) and isinstance(b, SScalar)):
            raise CodegenError("arithmetic on non-scalars")
        prov = a.prov | b.prov
        if a.ctype == "bool" and b.ctype == "bool" and op in "*+":  # gen:69-70, 78-79
            return SScalar("bool", "(%s %s %s)" % (a.code, "&&" if op == "*" else "||", b.code), prov)
        ct = "f64" if "f64" in (a.ctype, b.ctype) else "i64"
        return SScalar(ct, "(%s %s %s)" % (cast_to(a, ct), op, cast_to(b, ct)), prov)

    def ev_AddExpr(self, e, env, K):
        return self._arith(e, env, K, "+")

    def ev_SubExpr(self, e, env, K):
        return self._arith(e, env, K, "-")

    def ev_MulExpr(self, e, env, K):
        return self._arith(e, env, K, "*")

    def ev_DivExpr(self, e, env, K):
        return self._arith(e, env, K, "/")

    def ev_CompareExpr(self, e, env, K):
        a, b = self.ev(e.leftExpr, env, K), self.ev(e.rightExpr, env, K)
        op = e.compareType
        if isinstance(a, SNone) or isinstance(b, SNone):  # gen:86-91, 147-153
            lk = b if isinstance(a, SNone) else a
            if not isinstance(lk, SLookup):
                raise CodegenError("comparison with None needs a dictionary lookup")
            if lk.found is None:
                raise CodegenError("None test on a nested dictionary lookup")
            if op != CS.NE and getattr(lk, "found_is_filter", False):
                raise CodegenError("== None on a dictionary-valued lookup is not supported")
            return SScalar("bool", lk.found if op == CS.NE else "(!%s)" % lk.found)
        a = a.value() if isinstance(a, SLookup) else a
        b = b.value() if isinstance(b, SLookup) else b
        sym = {CS.EQ: "==", CS.NE: "!=", CS.LT: "<", CS.LTE: "<=", CS.GT: ">", CS.GTE: ">="}[op]
        if isinstance(a, SStr) or isinstance(b, SStr):
            return self.str_compare(K, a, b, sym)
        prov = a.prov | b.prov
        if a.ctype == "bool" or b.ctype == "bool":
            return SScalar("bool", "((bool)(%s) %s (bool)(%s))" % (a.code, sym, b.code), prov)
        ct = "f64" if "f64" in (a.ctype, b.ctype) else "i64"
        return SScalar("bool", "(%s %s %s)" % (cast_to(a, ct), sym, cast_to(b, ct)), prov)

    def str_compare(self, K, a, b, sym):
        if sym not in ("==", "!="):
            raise CodegenError("only == / != on strings")
        if isinstance(a, SStr) and a.kind == "const":
            a, b = b, a
        if not (isinstance(a, SStr) and isinstance(b, SStr)):
            raise CodegenError("string compared with a non-string")
        if b.kind == "const":
            if a.kind == "pack":
                v = 0
                bs = b.value.encode("ascii")[:a.n].ljust(a.n, b"\0")
                for ch in bs:
                    v = (v << 8) | ch
                return SScalar("bool", "(%s %s %dll)" % (a.code, sym, v), a.prov)
            if a.kind == "const":
                return SScalar("bool", "true" if (a.value == b.value) == (sym == "==") else "false")
            return SScalar("bool", "(%s %s %s)" % (self.str_code(K, a), sym, self.const_code(a.arg, a.col, b.value)), a.prov)
        if (a.arg, a.col) == (b.arg, b.col):
            return SScalar("bool", "(%s %s %s)" % (self.str_code(K, a), sym, self.str_code(K, b)), a.prov | b.prov)
        raise CodegenError("comparison of strings from different columns is not supported")

    def ev_ExtFuncExpr(self, e, env, K):
        s = e.symbol
        a = self.ev(e.inp1, env, K)
        a = a.value() if isinstance(a, SLookup) and s != XF.DictSize else a
        if s == XF.ExtractYear:  # gen:623-624
            st = ("year", a.stats[1]) if a.stats and a.stats[0] == "col" else None
            return SScalar("i64", "(%s / 10000)" % a.code, a.prov, E, st)
        if s == XF.DictSize:
            if isinstance(a, SLookup) and a.table.count_only:
                return SScalar("i64", "(%s ? sdqlrt::ldg1(c.%s_a0 + (%s < 0 ? 0 : %s)) : 0ll)" %
                               (a.found, a.table.name, a.slot, a.slot))
            if isinstance(a, SScalar):   # value of a distinct-count table (Q16)
                return a
            raise CodegenError("dictSize of an unsupported value")
        if s == XF.StringContains:
            pat, subj = a, self.ev(e.inp3, env, K)
            ptr, w = self.str_ptr(K, subj)
            fn = "str_find" if STRFIND_W else "str_find_bytes"
            tpos = self.text_position(K, subj, pat.value)
            if tpos:
                return SScalar("bool", "(%s >= 0)" % tpos, subj.prov)
            cand = self.text_candidate(K, subj, pat.value)
            if cand and fn == "str_find":
                fn = "str_find_rare"
            return SScalar("bool", "(%ssdqlrt::%s(%s, %d, %s, %d) >= 0)" % (cand + " && " if cand else "", fn, ptr, w, cstr(pat.value), len(pat.value)), subj.prov)
        b = self.ev(e.inp2, env, K)
        if s in (XF.StartsWith, XF.EndsWith, XF.FirstIndex):
            if not (isinstance(b, SStr) and b.kind == "const"):
                raise CodegenError("pattern must be a string constant")
            ptr, w = self.str_ptr(K, a)
            fn = {XF.StartsWith: "str_starts", XF.EndsWith: "str_ends", XF.FirstIndex: "str_find"}[s]
            if fn == "str_find" and not STRFIND_W:
                fn = "str_find_bytes"
            code = "sdqlrt::%s(%s, %d, %s, %d)" % (fn, ptr, w, cstr(b.value), len(b.value))
            tpos = self.text_position(K, a, b.value) if s == XF.FirstIndex else None
            if tpos:
                return SScalar("i64", "(long long)" + K.let("int", tpos), a.prov)
            cand = self.text_candidate(K, a, b.value) if s == XF.FirstIndex else None
            if cand:
                code = "(%s ? %s : -1)" % (cand, code.replace("sdqlrt::str_find(", "sdqlrt::str_find_rare("))
            if s == XF.FirstIndex:
                return SScalar("i64", "(long long)" + K.let("int", code), a.prov)
            return SScalar("bool", code, a.prov)
        if s == XF.SubStr:
            c3 = self.ev(e.inp3, env, K)
            lo, hi = int(e.inp2.value), int(e.inp3.value)
            ptr, w = self.str_ptr(K, a)
            n = hi - lo + 1
            if n > 7:
                raise CodegenError("substr longer than 7 characters")
            return SStr("pack", n=n, code="sdqlrt::str_pack(%s, %d, %d)" % (ptr, lo, n), prov=a.prov)
        raise CodegenError("unsupported external function %s" % s)


def frontend_dataset(a):
    return "db->" + a + "_dataset"


def and_chain(e):
    """flatten a * b * c (logical and after comp:277-292) into conjuncts, left to right."""
    if isinstance(e, ir.MulExpr):
        return and_chain(e.op1Expr) + and_chain(e.op2Expr)
    return [e]


def as_bool(v):
    if isinstance(v, SLookup):
        v = v.value()
    if not isinstance(v, SScalar):
        raise CodegenError("condition is not a scalar")
    return v.code if v.ctype == "bool" else "(%s != 0)" % v.code


def mark_iter(q, K, t):
    """element of a table-sourced kernel: p[0] = key (re-evaluated at the slot's representative), p[1] = value."""
    tok = ("it", K.name)
    prov = frozenset([tok])
    key = t.key_at(K, "i", prov) if t.key_fn is not None else None
    if t.inner is not None:
        raise CodegenError("iteration over a nested dictionary is only supported through dictSize / inner sums")
    val = t.value_at(K, "i", None, prov=prov)
    # the by-value part of the table key determines the slot (single-part keys)
    if key is not None and getattr(t, "key_shape", None) is not None and len(t.kept_pos) == 1:
        if t.key_shape == "scalar":
            key.det = key.det | frozenset([tok])
        else:
            nm = t.key_shape[t.kept_pos[0]]
            x = key.fields[nm]
            if not callable(x):
                x.det = x.det | frozenset([tok])
    return SPair(key, val)


# =============================================================================================
# module text
# =============================================================================================
def _stats_exprs(st):
    """-> (min expr, range expr) in generated host code for a key-part statistics tuple."""
    if st[0] == "col":
        return "a->cols[%d].min" % st[1], "sdqlhost::col_range(a->cols[%d])" % st[1]
    if st[0] == "year":
        return "(a->cols[%d].min / 10000)" % st[1], "(a->cols[%d].max / 10000 - a->cols[%d].min / 10000 + 1)" % (st[1], st[1])
    if st[0] == "range":
        return "%dll" % st[1], "%dll" % (st[2] - st[1] + 1)
    return "0ll", "-1ll"


def late_domain(t):
    """single-part key whose value range is only known at run time, from int64 aggregate arrays of other tables (and constant
    ranges): -> ([(table name, field)], lo, hi of the constant ranges) or None.  The table is laid out as a hash table of raw
    keys and turned into a direct-indexed one right before its build, when the range turns out small (redomain_table)."""
    if len(t.parts) != 1 or t.parts[0][0] not in ("agg", "union") or t.inner is not None:
        return None
    srcs, lo, hi = [], 0, 0

    def walk(st):
        nonlocal lo, hi
        if st[0] == "agg":
            srcs.append((st[1], st[2]))
            return True
        if st[0] == "range":
            lo, hi = min(lo, st[1]), max(hi, st[2])
            return True
        if st[0] == "union":
            return walk(st[1]) and walk(st[2])
        return False
    return (srcs, lo, hi) if walk(t.parts[0]) and srcs else None


def part_expr(K):
    if K.src[0] == "rel":
        return "part_%s" % K.name
    if K.src[0] == "tbl":
        return "part_%s" % K.src[1].name
    return "false"


def merge_code(q, K):
    """host code after the launch of K: combine this rank's partial outputs with the other ranks' (SURVEY.md 8e).
    Scalars and direct-indexed tables are all-reduced in place through the caller's merge callback; a table keyed by
    the partitioning column stays rank-local (co-partitioned) and marks its consumers as partial instead."""
    L = []
    pe = part_expr(K)
    if pe == "false":
        return L
    L.append("    if (%s) {" % pe)
    off = "(unsigned long long)((char*)(%s) - (char*)a->workspace)"
    sinks = [K.sink] if K.sink is not None else []
    for sk in sinks:
        if isinstance(sk, ReduceSink):
            for f in sk.fields:
                L.append("        if (a->merge(a->merge_ctx, %s, 1, %s)) return sdqlhost::fail(SDQLB200_E_ARG, \"merge callback failed\");" %
                         (off % ("c.sc + %d" % f[2]), "SDQLB200_SUM_F64" if f[1] == "f64" else "SDQLB200_SUM_I64"))
    for t in q.tables:
        if t.builder is not K:
            continue
        cop = " || ".join("((a->cols[%d].flags & SDQLB200_COL_PARTKEY) != 0)" % p[1] for p in t.parts if p[0] == "col") or "false"
        L.append("        part_%s = true;  // iteration over this table is partitioned (by key range or by owner rank)" % t.name)
        L.append("        if (!(%s)) {" % cop)
        if t.kind == "build" and getattr(t, "rep_probed", False):
            # the payload of such an entry is re-evaluated at its representative source row, which only the owner rank
            # holds: a probe on another rank would read row 0 of its own partition instead (silently wrong)
            L.append("            return sdqlhost::fail(SDQLB200_E_ARG, \"%s: table %s is built from a partitioned relation, keyed by a column the relation "
                     "is not partitioned on, and probed for its payload: replicate the relation or partition it on that key\");" % (q.name, t.name))
        L.append("            if (!c.%s.direct) {  // hashed partial dictionary: hash all-to-all + combine + all-gather (SDQLB200_MERGE_TABLE)" % t.name)
        L.append("                sdqlb200_table td; memset(&td, 0, sizeof td);")
        L.append("                td.keys = (uint64_t*)c.%s.keys; td.rep = c.%s.rep; td.cap = c.%s.cap; td.nfields = %d; td.f64_mask = %du;" %
                 (t.name, t.name, t.name, len(t.fields), sum(1 << j for j, (_, ct) in enumerate(t.fields) if ct == "f64")))
        for j in range(len(t.fields)):
            L.append("                td.agg[%d] = c.%s_a%d;" % (j, t.name, j))
        L.append("                if (a->merge(a->merge_ctx, (unsigned long long)(uintptr_t)&td, 0, SDQLB200_MERGE_TABLE)) return sdqlhost::fail(SDQLB200_E_ARG, \"%s: merging hashed table %s across ranks failed\");" % (q.name, t.name))
        nf64 = sum(1 for _, ct in t.fields if ct == "f64")
        words = 1 + nf64 + 2 * (len(t.fields) - nf64)
        L.append("            } else if (c.%s.cap <= sdqlhost::fused_merge_max()) {  // small table: presence + all fields in ONE all-reduce" % t.name)
        L.append("                sdqlrt::TblIO io; memset(&io, 0, sizeof io);")
        L.append("                io.rep = c.%s.rep; io.cap = c.%s.cap; io.nf = %d; io.f64_mask = %du;" %
                 (t.name, t.name, len(t.fields), sum(1 << j for j, (_, ct) in enumerate(t.fields) if ct == "f64")))
        for j in range(len(t.fields)):
            L.append("                io.agg[%d] = (sdqlrt::u64*)c.%s_a%d;" % (j, t.name, j))
        L.append("                const int og = sdqlhost::grid_for(c.%s.cap, 8, sms);" % t.name)
        L.append("                SDQL_LAUNCH(sdqlrt::k_merge_pack, og, sdqlrt::kBlock, 0, st, io, a->rank, mg_%s);" % t.name)
        L.append("                if (a->merge(a->merge_ctx, %s, (unsigned long long)c.%s.cap * %dull, SDQLB200_SUM_F64)) return sdqlhost::fail(SDQLB200_E_ARG, \"merge callback failed\");" %
                 (off % ("mg_%s" % t.name), t.name, words))
        L.append("                SDQL_LAUNCH(sdqlrt::k_merge_unpack, og, sdqlrt::kBlock, 0, st, io, a->rank, mg_%s);" % t.name)
        L.append("            } else {")
        L.append("            int sp_ = 1;  // large direct table: sparse merge of its occupied slots when few are (SDQLB200_MERGE_DIRECT)")
        L.append("            {")
        L.append("                sdqlb200_table td; memset(&td, 0, sizeof td);")
        L.append("                td.keys = nullptr; td.rep = c.%s.rep; td.cap = c.%s.cap; td.nfields = %d; td.f64_mask = %du;" %
                 (t.name, t.name, len(t.fields), sum(1 << j for j, (_, ct) in enumerate(t.fields) if ct == "f64")))
        for j in range(len(t.fields)):
            L.append("                td.agg[%d] = c.%s_a%d;" % (j, t.name, j))
        L.append("                sp_ = a->merge(a->merge_ctx, (unsigned long long)(uintptr_t)&td, 0, SDQLB200_MERGE_DIRECT);")
        L.append("                if (sp_ != 0 && sp_ != 1) return sdqlhost::fail(SDQLB200_E_ARG, \"%s: merging direct table %s across ranks failed\");" % (q.name, t.name))
        L.append("            }")
        L.append("            if (sp_ == 1) {")
        L.append("            const int og = sdqlhost::grid_for(c.%s.cap, 8, sms);" % t.name)
        L.append("            SDQL_LAUNCH(sdqlrt::k_owner_encode, og, sdqlrt::kBlock, 0, st, c.%s.rep, own_%s, c.%s.cap, a->rank);" % (t.name, t.name, t.name))
        L.append("            if (a->merge(a->merge_ctx, %s, (unsigned long long)c.%s.cap, SDQLB200_MIN_I32)) return sdqlhost::fail(SDQLB200_E_ARG, \"merge callback failed\");" % (off % ("own_%s" % t.name), t.name))
        L.append("            SDQL_LAUNCH(sdqlrt::k_owner_decode, og, sdqlrt::kBlock, 0, st, c.%s.rep, own_%s, c.%s.cap, a->rank);" % (t.name, t.name, t.name))
        j = 0
        while j < len(t.fields):  # consecutive aggregate arrays of one type are adjacent in the (zeroed) arena: one call
            k = j
            while k + 1 < len(t.fields) and t.fields[k + 1][1] == t.fields[j][1]:
                k += 1
            ct = t.fields[j][1]
            cnt = "(unsigned long long)((c.%s_a%d + c.%s.cap) - c.%s_a%d)" % (t.name, k, t.name, t.name, j)
            L.append("            if (a->merge(a->merge_ctx, %s, %s, %s)) return sdqlhost::fail(SDQLB200_E_ARG, \"merge callback failed\");" %
                     (off % ("c.%s_a%d" % (t.name, j)), cnt, "SDQLB200_SUM_F64" if ct == "f64" else "SDQLB200_SUM_I64"))
            j = k + 1
        L.append("            }")
        L.append("            }")
        consumers = [K2 for K2 in q.kernels if K2.src == ("tbl", t)]
        if consumers and all(iterates_without_rep(K2) for K2 in consumers):
            # the merged table is complete on every rank and its entries are self-contained: every rank iterates all of them,
            # what is computed from it is not partial any more (no concatenation of result rows, no further merges)
            L.append("            part_%s = false; c.%s_all = 1;" % (t.name, t.name))
        L.append("        }")
    L.append("    }")
    return L


def render_query(q):
    """-> CUDA text of one query: context struct, kernels, host driver."""
    n = q.name
    L = []
    ity = {"i32": "const int*", "f64": "const double*", "code": "const void*", "bytes": "const unsigned char*"}
    L.append("// " + "=" * 100)
    L.append("// %s" % n)
    L.append("// " + "=" * 100)
    L.append("struct %s_ctx {" % n)
    for i, (arg, col, rep) in enumerate(q.inputs):
        L.append("    %s in%d;  // %s.%s (%s)" % (ity[rep], i, arg, col, rep))
        if rep == "code":
            L.append("    int in%d_w;" % i)
    for a in q.args:
        L.append("    long long n_%s;" % a)
    for i, k in enumerate(q.consts):
        L.append("    long long k%d;  // %s" % (i, "/".join(str(x) for x in k)))
    for t in q.tables:
        P = max(1, len(t.parts))
        L.append("    sdqlrt::Tbl %s; long long %s_mn[%d], %s_rng[%d], %s_mul[%d]; int %s_sb[%d], %s_sk[%d];" %
                 (t.name, t.name, P, t.name, P, t.name, P, t.name, P, t.name, P))
        for j, (_, ct) in enumerate(t.fields):
            L.append("    %s* %s_a%d;" % (CT[ct], t.name, j))
        L.append("    int %s_all;  // multi-GPU: the table was merged across ranks and every rank iterates ALL of its entries" % t.name)
    if any(late_domain(t) for t in q.tables):
        L.append("    long long* mm;  // {min, max} of the aggregate values a late key domain comes from")
    for K in q.kernels:
        if K.body2 is not None:
            L.append("    unsigned %s_qo;  // per-warp queues of surviving row ids: offset in dynamic shared memory" % K.name)
        if K.byte_cols:
            L.append("    unsigned %s_bo;  // byte-row staging buffers: offset in dynamic shared memory" % K.name)
        if K.text_cols:
            L.append("    unsigned %s_to;  // text-scan candidate masks: offset in dynamic shared memory" % K.name)
        if K.pipe_mode() == "tma":
            L.append("    unsigned %s_ro; int %s_rs;  // column tile ring: byte offset in dynamic shared memory, stages" % (K.name, K.name))
    L.append("    double* sc; double* part; unsigned* cnt; unsigned long long* tcount;")
    for i in range(len(q.npart)):
        L.append("    long long part_off%d;" % i)
    L.append("    unsigned long long* res_count; long long res_cap;")
    for j in range(len(q.result_schema or [])):
        L.append("    long long* res%d;" % j)
    L.append("};")
    L.append("")
    for K in q.kernels:
        L.append(K.render())
        L.append("")
    # ---- host driver ----
    nres = len(q.result_schema or [])
    L.append("static int %s_run(sdqlb200_args* a) {" % n)
    L.append("    if (a->ncols != %d || a->nargs != %d || a->nconsts != %d)" % (len(q.inputs), len(q.args), len(q.consts)))
    L.append("        return sdqlhost::fail(SDQLB200_E_ARG, \"%s: expected %d inputs / %d args / %d consts\");" %
             (n, len(q.inputs), len(q.args), len(q.consts)))
    L.append("    %s_ctx c; memset(&c, 0, sizeof c);" % n)
    L.append("    sdqlhost::Arena ar(a->workspace, a->workspace_bytes);")
    L.append("    const int sms = sdqlhost_sms();")
    ckind = {"i32": "SDQLB200_I32", "f64": "SDQLB200_F64", "code": "SDQLB200_CODE", "bytes": "SDQLB200_BYTES"}
    for i, (arg, col, rep) in enumerate(q.inputs):
        L.append("    if (a->cols[%d].kind != %s) return sdqlhost::fail(SDQLB200_E_ARG, \"%s: input %d (%s.%s) must be %s\");" %
                 (i, ckind[rep], n, i, arg, col, rep))
        L.append("    c.in%d = (%s)a->cols[%d].data;" % (i, ity[rep], i))
        if rep == "code":
            L.append("    c.in%d_w = a->cols[%d].width;" % (i, i))
    for i, ar_ in enumerate(q.args):
        L.append("    c.n_%s = a->nrows[%d];" % (ar_, i))
    for i in range(len(q.consts)):
        L.append("    c.k%d = a->consts[%d];" % (i, i))
    if IDX32:
        for ar_ in q.args:
            L.append("    if (c.n_%s > 2000000000ll) return sdqlhost::fail(SDQLB200_E_ARG, \"%s: relation %s has more rows than a 32-bit index build supports\");" % (ar_, n, ar_))
    nt = len(q.tables)
    L.append("    sdqlhost::TblRegion tr[%d];" % max(1, nt))
    for ti, t in enumerate(q.tables):
        P = len(t.parts)
        if P == 0:
            raise CodegenError("%s: table %s was never keyed" % (n, t.name))
        t.index = ti
        mns = ", ".join(_stats_exprs(s)[0] for s in t.parts)
        rgs = ", ".join(_stats_exprs(s)[1] for s in t.parts)
        src = "c.n_%s" % t.src[1] if t.src[0] == "rel" else "c.%s.cap" % t.src[1].name
        nf = len(t.fields)
        L.append("    {")
        if t.src[0] == "rel":
            # a table that will be merged across ranks is planned from the relation's GLOBAL row count: all ranks agree on
            # direct vs hashed and on the slots, and a hashed table has room for the union of the ranks' keys
            ai = q.args.index(t.src[1])
            cop = " || ".join("((a->cols[%d].flags & SDQLB200_COL_PARTKEY) != 0)" % p_[1] for p_ in t.parts if p_[0] == "col") or "false"
            L.append("        const long long rows_ = (a->merge && a->nrows_global && ((a->part_mask >> %d) & 1u) && !(%s)) ? (long long)a->nrows_global[%d] : c.n_%s;" %
                     (ai, cop, ai, t.src[1]))
            src = "rows_"
        L.append("        long long mn[%d] = {%s}, rng[%d] = {%s};" % (P, mns, P, rgs))
        L.append("        void* ag[%d] = {nullptr};" % max(1, nf))
        # presence bits in front of tables that are probed and built selectively (predicates in front of the build)
        t.want_bits = bool(t.probed and t.builder is not None and t.builder.counted and BITS_FILTER)
        L.append("        if (!sdqlhost::size_table(&c.%s, %d, mn, rng, %s, c.%s_mn, c.%s_rng, c.%s_mul, ar, &tr[%d], %d, ag, %s, %s))" %
                 (t.name, P, src, t.name, t.name, t.name, ti, nf, "true" if t.want_bits else "false",
                  "true" if t.inner is not None else "false"))
        L.append("            return sdqlhost::fail(SDQLB200_E_ARG, \"%s: key domain of %s does not fit 63 bits\");" % (n, t.name))
        for j, (_, ct) in enumerate(t.fields):
            L.append("        c.%s_a%d = (%s*)ag[%d];" % (t.name, j, CT[ct], j))
        L.append("        if (sdqlhost::debug()) fprintf(stderr, \"[sdqlb200] %s: table %s (%s, %d part(s), %d field(s)): worst case %%s, %%lld slots, domain %%.3Lg, presence bits %%llu%%s\\n\", "
                 "c.%s.direct ? \"direct\" : \"hash\", (long long)c.%s.cap, tr[%d].dom, tr[%d].bdom, tr[%d].bmod ? \" (first part)\" : \"\");" %
                 (n, t.name, t.kind, P, nf, t.name, t.name, ti, ti, ti))
        for j, st_ in enumerate(t.parts):  # strided-dense key columns: the holes of the value range are packed away
            if st_[0] == "col":
                L.append("        c.%s_sb[%d] = sdqlhost::col_sb(a->cols[%d]); c.%s_sk[%d] = sdqlhost::col_sk(a->cols[%d]);" %
                         (t.name, j, st_[1], t.name, j, st_[1]))
        L.append("    }")
    for t in q.tables:
        nf64 = sum(1 for _, ct in t.fields if ct == "f64")
        if late_domain(t):  # may become a direct table of up to 65536 slots (redomain_table): merge buffers for that case
            L.append("    int* own_%s = a->merge ? ar.alloc<int>(c.%s.cap > 65536 ? c.%s.cap : 65536) : nullptr;" % (t.name, t.name, t.name))
            L.append("    double* mg_%s = a->merge ? ar.alloc<double>(65536ll * %d) : nullptr;" % (t.name, 1 + nf64 + 2 * (len(t.fields) - nf64)))
            continue
        L.append("    int* own_%s = a->merge ? ar.alloc<int>(c.%s.cap) : nullptr;" % (t.name, t.name))
        L.append("    double* mg_%s = (a->merge && c.%s.cap <= sdqlhost::fused_merge_max()) ? ar.alloc<double>(c.%s.cap * %d) : nullptr;" %
                 (t.name, t.name, t.name, 1 + nf64 + 2 * (len(t.fields) - nf64)))
    L.append("    const unsigned long long tail_off = ar.used;  // scalars, counters, partials: zeroed before every run")
    L.append("    c.sc = ar.alloc<double>(%d); c.cnt = ar.alloc<unsigned>(%d); c.tcount = ar.alloc<unsigned long long>(%d);" %
             (max(1, q.nsc), max(1, q.ncnt), max(1, q.ntcount)))
    if any(late_domain(t) for t in q.tables):
        L.append("    c.mm = ar.alloc<long long>(2);")
    # launch plans: aggregation tier, shared memory (tier table + column tile ring), resident CTAs per SM, grid
    for K in q.kernels:
        if K.src[0] == "rel":
            R = getattr(K, "R", 4)
            L.append("    const long long w_%s = (c.n_%s + %d) / %d;" % (K.name, K.src[1], R - 1, R))
        elif K.src[0] == "tbl":
            L.append("    const long long w_%s = c.%s.cap;" % (K.name, K.src[1].name))
        else:
            L.append("    const long long w_%s = 1;" % K.name)
        ring = K.pipe_mode() == "tma"
        L.append("    int g_%s = 1, tier_%s = 2; size_t sm_%s = 0;" % (K.name, K.name, K.name))
        plan_start = len(L)
        L.append("    {")
        if K.tiered:
            nf, tn = K.smem_nf, K.smem_tbl
            L.append("        const long long cap = c.%s.cap;" % tn)
            if TIER0_SMEM:
                L.append("        if (c.%s.direct && cap <= %d) { tier_%s = 0; sm_%s = (size_t)%d * sdqlrt::kBlock * 8 + (size_t)%d * sdqlrt::kBlock * 4; }" %
                         (tn, K.rcap, K.name, K.name, K.rcap * nf, K.rcap))
            else:
                L.append("        if (c.%s.direct && cap <= %d) tier_%s = 0;" % (tn, K.rcap, K.name))
            L.append("        else if (c.%s.direct && cap * (%d * 8 + 4) <= 65536) { tier_%s = 1; sm_%s = (size_t)cap * (%d * 8 + 4); }" %
                     (tn, nf, K.name, K.name, nf))
            fn = "(tier_%s == 0 ? (const void*)%s<0> : tier_%s == 1 ? (const void*)%s<1> : (const void*)%s<2>)" % (
                K.name, K.name, K.name, K.name, K.name)
        elif K.templated:
            fn = "(const void*)%s<2>" % K.name
        else:
            fn = "(const void*)%s" % K.name
        if K.body2 is not None:
            L.append("        c.%s_qo = (unsigned)((sm_%s + 15) & ~(size_t)15); sm_%s = c.%s_qo + (sdqlrt::kBlock / sdqlrt::kLanes) * (sdqlrt::kLanes * %d) * 4;" %
                     (K.name, K.name, K.name, K.name, getattr(K, "R", 4) + 1))
        if K.byte_cols:
            # 16 bytes of slack on both sides: the word-wise string search reads whole aligned words around a row
            L.append("        c.%s_bo = (unsigned)((sm_%s + 15) & ~(size_t)15); sm_%s = c.%s_bo + %du;" %
                     (K.name, K.name, K.name, K.name, 32 + 1024 * sum(K.byte_cols.values())))
        if K.text_cols:
            L.append("        c.%s_to = (unsigned)((sm_%s + 15) & ~(size_t)15); sm_%s = c.%s_to + %du;" %
                     (K.name, K.name, K.name, K.name, (1024 + (2048 if TEXTRESOLVE else 0)) * sum(len(pats) for _, pats in K.text_cols.values())))
        if ring:
            widths = " + ".join({"i32": "4", "f64": "8", "code": "(size_t)a->cols[%d].width" % idx}[rep]
                                for (col, rep), (arr, idx) in K.scan_cols.items())
            L.append("        sdqlhost::RingPlan rp;")
            L.append("        if (!sdqlhost::plan_ring(sm_%s, (%s) * (size_t)(sdqlrt::kBlock * %d), &rp))" % (K.name, widths, K.ring_rows()))
            L.append("            return sdqlhost::fail(SDQLB200_E_ARG, \"%s: scan too wide for the shared-memory column ring\");" % K.name)
            L.append("        c.%s_ro = rp.ring_off; c.%s_rs = rp.stages; sm_%s = rp.smem;" % (K.name, K.name, K.name))
            for (col, rep), (arr, idx) in K.scan_cols.items():
                L.append("        if ((size_t)a->cols[%d].data & 15) return sdqlhost::fail(SDQLB200_E_ARG, \"%s: column %s is not 16-byte aligned\");" % (idx, K.name, col))
        L.append("        const int nb = sdqlhost_occupancy(%s, sm_%s);" % (fn, K.name))
        if ring:
            L.append("        g_%s = sdqlhost::grid_for_tiles((c.n_%s + sdqlrt::kBlock * %d - 1) / (sdqlrt::kBlock * %d), nb, sms);" %
                     (K.name, K.src[1], K.ring_rows(), K.ring_rows()))
        else:
            L.append("        g_%s = sdqlhost::grid_for(w_%s, nb < 8 ? nb : 8, sms);" % (K.name, K.name))
        L.append("    }")
        K.plan_lines = L[plan_start:]  # emitted again when a table of the kernel gets its key domain at run time
    L.append("    long long npart = 0;")
    pi = 0
    for K in q.kernels:
        if isinstance(K.sink, ReduceSink) and K.sink.fields:
            L.append("    c.part_off%d = npart; npart += %dll * g_%s;" % (pi, len(K.sink.fields), K.name))
            pi += 1
    L.append("    c.part = ar.alloc<double>(npart);")
    L.append("    c.res_count = ar.alloc<unsigned long long>(1);")
    L.append("    const unsigned long long zero_end = ar.used;")
    L.append("    c.res_cap = %s; if (c.res_cap < 1) c.res_cap = 1;" % q.res_cap_expr)
    for j in range(nres):
        L.append("    c.res%d = ar.alloc<long long>(c.res_cap);" % j)
    L.append("    a->workspace_needed = ar.used;")
    L.append("    if (!ar.ok()) return sdqlhost::fail(SDQLB200_E_WORKSPACE, \"%s: workspace of %%llu bytes needed\", ar.used);" % n)
    L.append("    cudaStream_t st = (cudaStream_t)a->stream;")
    L.append("    SDQL_CUDA(cudaEventRecord(sdqlhost_ev(0), st));")
    L.append("    const unsigned pm = a->merge ? a->part_mask : 0u;  // multi-GPU: which relation arguments are partitioned")
    for t in q.tables:
        L.append("    bool part_%s = false;" % t.name)
    for K in q.kernels:
        if K.src[0] == "rel":
            L.append("    const bool part_%s = (pm >> %d) & 1u;" % (K.name, q.args.index(K.src[1])))
    # cardinality passes: a kernel with predicates in front of a table build first counts the rows that reach the build,
    # when its tables are big enough for their size to matter (or always when the count has to agree across ranks)
    owned = {}
    for t in q.tables:
        owned.setdefault(t.builder, []).append(t)
    for K in q.kernels:
        K.count_ok = bool(K.counted and K.pipe_mode() != "tma" and owned.get(K))
        if K.count_ok:
            tot = " + ".join("tr[%d].len" % t.index for t in owned[K])
            # worth it only when initialising / probing the worst-case tables costs clearly more than scanning the
            # predicate columns a second time
            if K.src[0] == "rel":
                pb = sum({"i32": 4, "f64": 8, "code": 1}[rep] for (_, rep) in (K.pred_cols or ())) + sum(K.byte_cols.values())
                pb += sum(w_ for w_, _ in K.text_cols.values())  # the warp text scan streams the whole string column
                scan = "(unsigned long long)c.n_%s * %dull" % (K.src[1], max(4, pb))
                ai_ = q.args.index(K.src[1])
                scan_g = "(unsigned long long)((a->nrows_global && ((a->part_mask >> %d) & 1u)) ? a->nrows_global[%d] : c.n_%s) * %dull" % (
                    ai_, ai_, K.src[1], max(4, pb))
            else:
                scan = "(unsigned long long)c.%s.cap * 12ull" % K.src[1].name
                scan_g = "(~0ull >> 4)"  # the source table's size may differ between ranks: only the absolute rule below
            L.append("    const bool cnt_%s = (%s) >= sdqlhost::count_min_bytes() && ((%s) >= sdqlhost::count_min_ratio() * (%s) || (%s) >= sdqlhost::count_big_bytes());" %
                     (K.name, tot, tot, scan, tot))
            # tables merged across ranks are planned for the GLOBAL row count (600 M lineitems -> 2^30 slots, whatever the
            # predicates let through): the same rule decides from rank-independent numbers (global rows) whether to count, and
            # the ranks' counts are summed, so every rank re-plans the table alike (Q20 on 2 GPUs: 16.9 ms with the worst-case
            # plan, 2.8 ms counted).  Counting everything big lost: q13_k0's second text scan cost 3 ms for nothing (visit 11)
            L.append("    const bool cntm_%s = (%s) >= sdqlhost::count_min_bytes() && ((%s) >= sdqlhost::count_min_ratio() * (%s) || (%s) >= sdqlhost::count_big_bytes());" %
                     (K.name, tot, tot, scan_g, tot))
            L.append("    const bool late_%s = cnt_%s || (a->merge != nullptr && cntm_%s);  // its tables are initialised right before it" %
                     (K.name, K.name, K.name))
    L.append("    g_trace = (a->flags & SDQLB200_F_TRACE) != 0 || sdqlhost::debug();")
    L.append("    sdqlhost_step(st, nullptr);")
    for t in q.tables:
        K = t.builder
        guard = "if (!late_%s) " % K.name if (K is not None and K.count_ok) else ""
        if late_domain(t) and K is not None:
            continue  # laid out and initialised right before its build (key domain known then)
        L.append("    %sSDQL_CUDA(sdqlhost_init_table(a->workspace, tr[%d], st));" % (guard, t.index))
    L.append("    SDQL_CUDA(cudaMemsetAsync((char*)a->workspace + tail_off, 0, zero_end - tail_off, st));")
    L.append("    sdqlhost_step(st, \"%s:init\");" % n)
    L.append("    int launches = 0; const bool kt = (a->flags & SDQLB200_F_KERNEL_TIMES) != 0;")
    L.append("    if (kt) SDQL_CUDA(cudaEventRecord(sdqlhost_kev(0), st));")
    for K in q.kernels:
        if K.src[0] == "tbl":  # the source table may have been right-sized since the launch plan was made
            L.append("    { const int g2 = sdqlhost::grid_for(c.%s.cap, 8, sms); if (g2 < g_%s) g_%s = g2; }" %
                     (K.src[1].name, K.name, K.name))
        for t in owned.get(K, []):
            ld = late_domain(t)
            if not ld:
                continue
            srcs, lo, hi = ld
            nf = len(t.fields)
            L.append("    {   // key domain of %s: value range of %s, known now that %s complete" %
                     (t.name, ", ".join("%s.a%d" % sj for sj in srcs), "it is" if len(srcs) == 1 else "they are"))
            for tn_, j in srcs:
                L.append("        SDQL_LAUNCH(sdqlrt::k_minmax_i64, sdqlhost::grid_for(c.%s.cap, 8, sms), sdqlrt::kBlock, 0, st, (const long long*)c.%s_a%d, c.%s.cap, c.mm);" %
                         (tn_, tn_, j, tn_))
            L.append("        SDQL_CUDA(cudaGetLastError());")
            L.append("        long long h_mm[2] = {0, 0};")
            L.append("        SDQL_CUDA(cudaMemcpyAsync(h_mm, c.mm, 16, cudaMemcpyDeviceToHost, st));")
            L.append("        SDQL_CUDA(cudaStreamSynchronize(st));")
            L.append("        if (h_mm[0] > %dll) h_mm[0] = %dll;" % (lo, lo))
            L.append("        if (h_mm[1] < %dll) h_mm[1] = %dll;" % (hi, hi))
            L.append("        if (a->merge) {  // every rank must derive the same plan: min / max over the ranks (int32 pair {min, -max})")
            L.append("            const bool fits_ = h_mm[0] >= -2000000000ll && h_mm[1] <= 2000000000ll;")
            L.append("            int h_pr[2] = {fits_ ? (int)h_mm[0] : -2147483647, fits_ ? (int)-h_mm[1] : -2147483647};")
            L.append("            SDQL_CUDA(cudaMemcpyAsync(c.mm, h_pr, 8, cudaMemcpyHostToDevice, st));")
            L.append("            if (a->merge(a->merge_ctx, (unsigned long long)((char*)c.mm - (char*)a->workspace), 2, SDQLB200_MIN_I32)) return sdqlhost::fail(SDQLB200_E_ARG, \"merge callback failed\");")
            L.append("            SDQL_CUDA(cudaMemcpyAsync(h_pr, c.mm, 8, cudaMemcpyDeviceToHost, st));")
            L.append("            SDQL_CUDA(cudaStreamSynchronize(st));")
            L.append("            h_mm[0] = h_pr[0]; h_mm[1] = -(long long)h_pr[1];")
            L.append("        }")
            L.append("        void* ag[%d] = {%s};" % (max(1, nf), ", ".join("c.%s_a%d" % (t.name, j) for j in range(nf)) or "nullptr"))
            L.append("        const bool rd_ = sdqlhost::redomain_table(&c.%s, (char*)a->workspace, &tr[%d], h_mm[0], h_mm[1] - h_mm[0] + 1, c.%s_mn, c.%s_rng, c.%s_mul, ag);" %
                     (t.name, t.index, t.name, t.name, t.name))
            L.append("        if (sdqlhost::debug()) fprintf(stderr, \"[sdqlb200] %s: key domain of %s [%%lld, %%lld] -> %%s, %%lld slots\\n\", h_mm[0], h_mm[1], c.%s.direct ? \"direct\" : \"hash\", (long long)c.%s.cap);" %
                     (K.name, t.name, t.name, t.name))
            for j, (_, ct) in enumerate(t.fields):
                L.append("        c.%s_a%d = (%s*)ag[%d];" % (t.name, j, CT[ct], j))
            L.append("        if (rd_) {  // the launch plan of %s again: aggregation tier, shared memory, grid" % K.name)
            L.append("            tier_%s = 2; sm_%s = 0;" % (K.name, K.name))
            L += ["        " + x for x in K.plan_lines]
            L.append("        }")
            L.append("        SDQL_CUDA(sdqlhost_init_table(a->workspace, tr[%d], st));" % t.index)
            L.append("        sdqlhost_step(st, \"%s:key-domain\");" % K.name)
            L.append("    }")
        bits_tabs = [t for t in owned.get(K, []) if getattr(t, "want_bits", False)]
        if K.count_ok or bits_tabs:
            pe = part_expr(K)
            cops = []
            for t in owned[K]:
                cops.append("(" + (" || ".join("((a->cols[%d].flags & SDQLB200_COL_PARTKEY) != 0)" % p_[1] for p_ in t.parts if p_[0] == "col") or "false") + ")")
            L.append("    // partial tables that are merged across ranks are planned from rank-independent numbers (all ranks must agree on")
            L.append("    // the plan) and get no presence filter (the merge adds keys behind its back)")
            L.append("    const bool merged_%s = a->merge != nullptr && %s && !(%s);" % (K.name, pe, " && ".join(cops)))
            for t in bits_tabs:
                L.append("    if (merged_%s) c.%s.bits = nullptr;" % (K.name, t.name))
        if K.count_ok:
            L.append("    if (late_%s) {" % K.name)
            L.append("        const bool merged_ = merged_%s;" % K.name)
            L.append("        if (merged_ ? cntm_%s : cnt_%s) {" % (K.name, K.name))
            uses_smem = bool(K.byte_cols or K.text_cols or K.body2 is not None)
            if uses_smem:  # same shared-memory layout as the real launch (queues / staging sit behind the tier table)
                L.append("            sdqlhost_occupancy((const void*)%s<3>, sm_%s);  // raises the dynamic shared memory limit" % (K.name, K.name))
            L.append("            SDQL_LAUNCH(%s<3>, g_%s, sdqlrt::kBlock, %s, st, c);" % (K.name, K.name, "sm_%s" % K.name if uses_smem else "0"))
            L.append("            SDQL_CUDA(cudaGetLastError());")
            L.append("            if (merged_ && a->merge(a->merge_ctx, (unsigned long long)((char*)(c.tcount + %d) - (char*)a->workspace), 1, SDQLB200_SUM_I64))" % K.count_slot)
            L.append("                return sdqlhost::fail(SDQLB200_E_ARG, \"merge callback failed\");  // sum over ranks >= keys of the union")
            L.append("            unsigned long long h_cnt = 0;")
            L.append("            SDQL_CUDA(cudaMemcpyAsync(&h_cnt, c.tcount + %d, 8, cudaMemcpyDeviceToHost, st));" % K.count_slot)
            L.append("            SDQL_CUDA(cudaStreamSynchronize(st));")
            L.append("            sdqlhost_step(st, \"%s:count\");" % K.name)
            for t in owned[K]:
                nf = len(t.fields)
                L.append("            {")
                L.append("                void* ag[%d] = {%s};" % (max(1, nf), ", ".join("c.%s_a%d" % (t.name, j) for j in range(nf)) or "nullptr"))
                L.append("                const bool rp_ = sdqlhost::replan_table(&c.%s, (char*)a->workspace, &tr[%d], (long long)h_cnt, ag, merged_ ? 4 : 64);" % (t.name, t.index))
                L.append("                if (sdqlhost::debug()) fprintf(stderr, \"[sdqlb200] %s: %%llu rows reach the build of %s -> %%s, %%lld slots%%s\\n\", h_cnt, c.%s.direct ? \"direct\" : \"hash\", (long long)c.%s.cap, rp_ ? \" (re-planned)\" : \"\");" % (K.name, t.name, t.name, t.name))
                for j, (_, ct) in enumerate(t.fields):
                    L.append("                c.%s_a%d = (%s*)ag[%d];" % (t.name, j, CT[ct], j))
                if getattr(t, "want_bits", False):
                    # the new layout brought the presence filter back: still none for a merged table (the merge adds the
                    # other ranks' keys behind the bitmap's back -- probes for them would miss: Q20 on 2 GPUs lost 28 % of its rows)
                    L.append("                if (merged_) c.%s.bits = nullptr;" % t.name)
                L.append("            }")
            L.append("        }")
            for t in owned[K]:
                L.append("        SDQL_CUDA(sdqlhost_init_table(a->workspace, tr[%d], st));" % t.index)
            L.append("        sdqlhost_step(st, \"%s:table-init\");" % K.name)
            L.append("    }")
        if K.templated:
            if K.tiered:
                L.append("    if (tier_%s == 0) SDQL_LAUNCH(%s<0>, g_%s, sdqlrt::kBlock, sm_%s, st, c);" % (K.name, K.name, K.name, K.name))
                L.append("    else if (tier_%s == 1) SDQL_LAUNCH(%s<1>, g_%s, sdqlrt::kBlock, sm_%s, st, c);" % (K.name, K.name, K.name, K.name))
                L.append("    else SDQL_LAUNCH(%s<2>, g_%s, sdqlrt::kBlock, sm_%s, st, c);" % (K.name, K.name, K.name))
                L.append("    a->tier = tier_%s;" % K.name)
            else:
                L.append("    SDQL_LAUNCH(%s<2>, g_%s, sdqlrt::kBlock, sm_%s, st, c);" % (K.name, K.name, K.name))
        else:
            L.append("    SDQL_LAUNCH(%s, g_%s, sdqlrt::kBlock, sm_%s, st, c);" % (K.name, K.name, K.name))
        L.append("    SDQL_CUDA(cudaGetLastError()); ++launches;")
        L.append("    sdqlhost_step(st, \"%s\");" % K.name)
        for t in bits_tabs:  # presence bits of the finished table (one pass over its slots; never for merged tables)
            L.append("    if (c.%s.bits) { SDQL_LAUNCH(sdqlrt::k_tbl_bits, sdqlhost::grid_for(c.%s.cap, 8, sms), sdqlrt::kBlock, 0, st, c.%s); SDQL_CUDA(cudaGetLastError()); }" %
                     (t.name, t.name, t.name))
        if bits_tabs:
            L.append("    sdqlhost_step(st, \"%s:bits\");" % K.name)
        # the kernel's own time ends here: the cross-GPU merge that follows is charged to the next interval
        L.append("    if (kt && launches < 24) SDQL_CUDA(cudaEventRecord(sdqlhost_kev(launches), st));")
        mc = merge_code(q, K)
        L += mc
        if mc:
            L.append("    sdqlhost_step(st, \"%s:merge\");" % K.name)
    L.append("    SDQL_CUDA(cudaEventRecord(sdqlhost_ev(1), st));")
    L.append("    a->launches = launches;")
    L.append("    a->result_partial = %s ? 1 : 0;" % part_expr(q.kernels[-1]))
    L.append("    long long* rcols[%d] = {%s};" % (max(1, nres), ", ".join("c.res%d" % j for j in range(nres)) or "nullptr"))
    L.append("    return sdqlhost_fetch(a, st, c.res_count, c.res_cap, %d, rcols);" % nres)
    L.append("}")
    L.append("")
    return "\n".join(L)


def manifest_of(q):
    return {
        "name": q.name,
        "args": q.args,
        "schemas": {a: [[c, list(k) if isinstance(k, tuple) else k] for c, k in q.schemas[a]] for a in q.args},
        "inputs": [list(k) for k in q.inputs],
        "consts": [list(k) for k in q.consts],
        "result_kind": q.result_kind,
        "result": [list(s) for s in (q.result_schema or [])],
        "kernels": [{"name": K.name, "source": list(K.src[:1]) + [K.src[1] if K.src[0] == "rel" else (K.src[1].name if K.src[0] == "tbl" else "")],
                     "scan_cols": [[c, r] for (c, r) in K.scan_cols],
                     # fixed-width string columns whose bytes the kernel reads for every scanned row: [column, width]
                     "byte_cols": [[c, w] for (c, w) in getattr(K, "str_scan", {}).items()]} for K in q.kernels],
    }


MODULE_HEAD = r'''// GENERATED by sdqlpy_b200.codegen -- do not edit.  Source workload: %(src)s
#include "sdqlb200_host.h"

#ifndef SDQLB200_EMU
// per host thread: one thread drives one GPU (events, occupancy results and staging buffers belong to its device)
static thread_local cudaEvent_t g_ev[2];
static thread_local bool g_ev_init = false;
static cudaEvent_t sdqlhost_ev(int i) {
    if (!g_ev_init) { cudaEventCreate(&g_ev[0]); cudaEventCreate(&g_ev[1]); g_ev_init = true; }
    return g_ev[i];
}
static thread_local cudaEvent_t g_kev[25];
static thread_local bool g_kev_init = false;
static cudaEvent_t sdqlhost_kev(int i) {
    if (!g_kev_init) { for (int k = 0; k < 25; ++k) cudaEventCreate(&g_kev[k]); g_kev_init = true; }
    return g_kev[i];
}
// resident CTAs per SM of `fn` at `smem` bytes of dynamic shared memory (also raises the kernel's dynamic shared
// memory limit); cached per (kernel, size) so steady-state launches make no driver queries
static int sdqlhost_occupancy(const void* fn, size_t smem) {
    struct Ent { const void* fn; size_t smem; int nb; };
    static thread_local Ent cache[512];
    static thread_local int used = 0;
    for (int i = 0; i < used; ++i)
        if (cache[i].fn == fn && cache[i].smem == smem) return cache[i].nb;
    int nb = 0;
    if (smem > 48 * 1024) cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, sdqlrt::kBlock, smem) != cudaSuccess || nb < 1) nb = 1;
    if (used < 512) cache[used++] = Ent{fn, smem, nb};
    return nb;
}
static int sdqlhost_sms() {
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms < 1) sms = 148; }
    return sms;
}
#endif

// SDQLB200_F_TRACE (or SDQLB200_DEBUG=1): wall-clock time of every step of the host driver on stderr (the stream is
// drained after each step, so the figures are per step and the query as a whole runs slower than normal).
// what == nullptr restarts the clock.
static thread_local bool g_trace = false;
static void sdqlhost_step(cudaStream_t st, const char* what) {
    if (!g_trace) return;
    static thread_local double last = 0;
    cudaStreamSynchronize(st);
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    const double now = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    if (what) fprintf(stderr, "[sdqlb200] step %%-24s %%9.3f ms\n", what, now - last);
    last = now;
}

static thread_local unsigned long long g_init_bytes = 0;  // bytes of table arrays initialised since the last sdqlb200_stats() call
// fill a dictionary's key / representative arrays with 0xFF (free) and zero its aggregate arrays
static cudaError_t sdqlhost_init_table(void* ws, const sdqlhost::TblRegion& r, cudaStream_t st) {
    g_init_bytes += r.ff_len + r.z_len;
    cudaError_t e = cudaMemsetAsync((char*)ws + r.off, 0xFF, r.ff_len, st);
    if (e == cudaSuccess && r.z_len) e = cudaMemsetAsync((char*)ws + r.off + r.ff_len, 0, r.z_len, st);
    return e;
}

// copy the result rows to host buffers (after the query's kernels).  The row counter sits directly in front of the
// result columns in the arena, so a small result (the common case: aggregates) is one D2H copy into a pinned staging
// buffer and one synchronisation; large results copy the counter first, then exactly count rows per column.
static int sdqlhost_fetch(sdqlb200_args* a, cudaStream_t st, unsigned long long* d_count, long long cap, int nf,
                          long long* const* d_cols) {
    unsigned long long cnt = 0;
    const size_t span = nf ? (size_t)((char*)(d_cols[nf - 1] + cap) - (char*)d_count) : 8;
    static thread_local char* stage = nullptr;  // per thread: concurrent sdqlb200_run calls do not share it
    const size_t kStage = 1 << 20;
#ifndef SDQLB200_EMU
    if (!stage && cudaHostAlloc((void**)&stage, kStage, cudaHostAllocDefault) != cudaSuccess) stage = nullptr;
#endif
    const bool small = stage && span <= kStage;
    if (small) SDQL_CUDA(cudaMemcpyAsync(stage, d_count, span, cudaMemcpyDeviceToHost, st));
    else SDQL_CUDA(cudaMemcpyAsync(&cnt, d_count, 8, cudaMemcpyDeviceToHost, st));
    SDQL_CUDA(cudaStreamSynchronize(st));
    if (small) memcpy(&cnt, stage, 8);
    float ms = 0;
    cudaEventElapsedTime(&ms, sdqlhost_ev(0), sdqlhost_ev(1));
    a->device_ms = ms;
    if (a->flags & SDQLB200_F_KERNEL_TIMES)
        for (int k = 0; k < a->launches && k < 24; ++k) cudaEventElapsedTime(&a->kernel_ms[k], sdqlhost_kev(k), sdqlhost_kev(k + 1));
    if ((long long)cnt > cap) return sdqlhost::fail(SDQLB200_E_ARG, "result overflow: %%llu rows > capacity %%lld", cnt, cap);
    a->result.count = (long long)cnt;
    a->result.nfields = nf;
    for (int j = 0; j < 32; ++j) a->result.cols[j] = nullptr;
    if (a->flags & SDQLB200_F_NOFETCH) return SDQLB200_OK;
    for (int j = 0; j < nf; ++j) {
        a->result.cols[j] = (int64_t*)malloc((cnt ? cnt : 1) * 8);
        if (!cnt) continue;
        if (small) memcpy(a->result.cols[j], stage + ((char*)d_cols[j] - (char*)d_count), cnt * 8);
        else SDQL_CUDA(cudaMemcpyAsync(a->result.cols[j], d_cols[j], cnt * 8, cudaMemcpyDeviceToHost, st));
    }
    if (!small) SDQL_CUDA(cudaStreamSynchronize(st));
    return SDQLB200_OK;
}

'''

MODULE_TAIL = r'''
struct sdql_entry { const char* name; int (*fn)(sdqlb200_args*); };
static const sdql_entry g_queries[] = {
%(entries)s
};
static const char g_manifest[] = %(manifest)s;

extern "C" {
int sdqlb200_num_queries(void) { return (int)(sizeof g_queries / sizeof g_queries[0]); }
const char* sdqlb200_query_name(int i) { return (i >= 0 && i < sdqlb200_num_queries()) ? g_queries[i].name : nullptr; }
const char* sdqlb200_manifest(void) { return g_manifest; }
const char* sdqlb200_last_error(void) { return sdqlhost::g_err; }
int sdqlb200_run(const char* query, sdqlb200_args* args) {
    if (!query || !args) return sdqlhost::fail(SDQLB200_E_ARG, "null argument");
    for (int i = 0; i < sdqlb200_num_queries(); ++i)
        if (!strcmp(g_queries[i].name, query)) return g_queries[i].fn(args);
    return sdqlhost::fail(SDQLB200_E_NOQUERY, "no query named %%s in this module", query);
}
int sdqlb200_stats(uint64_t* out, int32_t n) {
    if (!out || n < 0) return sdqlhost::fail(SDQLB200_E_ARG, "stats: bad arguments");
    unsigned long long v[sdqlrt::kStCount + 1];
    memset(v, 0, sizeof v);
    int have = 0;
#ifdef SDQLB200_STATS
    have = 1;
#ifndef SDQLB200_EMU
    SDQL_CUDA(cudaDeviceSynchronize());
    SDQL_CUDA(cudaMemcpyFromSymbol(v, sdqlrt::g_stats, sizeof(unsigned long long) * sdqlrt::kStCount));
    unsigned long long z[sdqlrt::kStCount];
    memset(z, 0, sizeof z);
    SDQL_CUDA(cudaMemcpyToSymbol(sdqlrt::g_stats, z, sizeof z));
#else
    memcpy(v, sdqlrt::g_stats, sizeof(unsigned long long) * sdqlrt::kStCount);
    memset(sdqlrt::g_stats, 0, sizeof(unsigned long long) * sdqlrt::kStCount);
#endif
#endif
    v[sdqlrt::kStCount] = g_init_bytes;
    g_init_bytes = 0;
    for (int i = 0; i < n && i <= sdqlrt::kStCount; ++i) out[i] = v[i];
    return have;
}
void sdqlb200_result_free(sdqlb200_result* r) {
    if (!r) return;
    for (int j = 0; j < 32; ++j) { free(r->cols[j]); r->cols[j] = nullptr; }
    r->count = 0;
}

static int sdql_tbl_io(const sdqlb200_table* t, sdqlrt::TblIO* io) {
    if (!t || !t->keys || t->nfields < 0 || t->nfields > 16 || t->cap < 1 || (t->cap & (t->cap - 1)))
        return sdqlhost::fail(SDQLB200_E_ARG, "bad table descriptor");
    io->keys = (sdqlrt::u64*)t->keys; io->rep = t->rep; io->cap = t->cap; io->nf = t->nfields; io->f64_mask = t->f64_mask;
    for (int j = 0; j < 16; ++j) io->agg[j] = (sdqlrt::u64*)t->agg[j];
    return SDQLB200_OK;
}
int sdqlb200_table_count(const sdqlb200_table* t, int32_t world, uint64_t* d_counts, void* stream) {
    sdqlrt::TblIO io;
    if (int rc = sdql_tbl_io(t, &io)) return rc;
    if (world < 1 || !d_counts) return sdqlhost::fail(SDQLB200_E_ARG, "table_count: bad arguments");
    SDQL_LAUNCH(sdqlrt::k_tbl_count, sdqlhost::grid_for(io.cap, 8, sdqlhost_sms()), sdqlrt::kBlock, 0, stream, io, world,
                (sdqlrt::u64*)d_counts);
    SDQL_CUDA(cudaGetLastError());
    return SDQLB200_OK;
}
int sdqlb200_table_pack(const sdqlb200_table* t, int32_t world, int32_t rank, const int32_t* d_own,
                        const uint64_t* d_offsets, uint64_t* d_cursor, int64_t* d_records, void* stream) {
    sdqlrt::TblIO io;
    if (int rc = sdql_tbl_io(t, &io)) return rc;
    if (world < 1 || !d_offsets || !d_cursor || !d_records) return sdqlhost::fail(SDQLB200_E_ARG, "table_pack: bad arguments");
    SDQL_LAUNCH(sdqlrt::k_tbl_pack, sdqlhost::grid_for(io.cap, 8, sdqlhost_sms()), sdqlrt::kBlock, 0, stream, io, world, rank,
                (const int*)d_own, (const sdqlrt::u64*)d_offsets, (sdqlrt::u64*)d_cursor, (sdqlrt::i64*)d_records);
    SDQL_CUDA(cudaGetLastError());
    return SDQLB200_OK;
}
int sdqlb200_table_absorb(const sdqlb200_table* t, const int64_t* d_records, int64_t n, int32_t mode, int32_t rank,
                          int32_t* d_own, void* stream) {
    sdqlrt::TblIO io;
    if (int rc = sdql_tbl_io(t, &io)) return rc;
    if (n < 0 || (n > 0 && !d_records) || (mode == 0 && !d_own) || (mode == 1 && !io.rep))
        return sdqlhost::fail(SDQLB200_E_ARG, "table_absorb: bad arguments");
    if (n == 0) return SDQLB200_OK;
    SDQL_LAUNCH(sdqlrt::k_tbl_absorb, sdqlhost::grid_for(n, 8, sdqlhost_sms()), sdqlrt::kBlock, 0, stream, io,
                (const sdqlrt::i64*)d_records, (sdqlrt::i64)n, mode, rank, (int*)d_own);
    SDQL_CUDA(cudaGetLastError());
    return SDQLB200_OK;
}
}
'''


def render_module(queries, src_name):
    body = [MODULE_HEAD % {"src": src_name}]
    for q in queries:
        body.append(render_query(q))
    man = json.dumps({"queries": [manifest_of(q) for q in queries]})
    lit = "\n".join("    " + cstr(man[i:i + 100]) for i in range(0, len(man), 100))
    entries = "\n".join('    {"%s", %s_run},' % (q.name, q.name) for q in queries)
    body.append(MODULE_TAIL % {"entries": entries, "manifest": "\n" + lit})
    return "\n".join(body)


def schema_from_in_type(in_type_node_or_obj):
    raise NotImplementedError
