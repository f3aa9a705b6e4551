"""SDQL IR -- the data model the B200 code generator consumes.

Node kinds and field names follow the reference IR (/root/reference/src/sdqlpy/lib/sdql_ir.py:8-98 types,
:123-385 expressions) because the IR is the contract between the Python front end and any back end
("the SDQL IR stays as it is").  The implementation is new: plain dataclass-like nodes, no operator
overloading (the front end builds nodes directly from the Python AST, see frontend.py), no global
star imports.
"""
import enum
import itertools

# ---------------------------------------------------------------------------------------------
# types (ref sdql_ir.py:8-98)
# ---------------------------------------------------------------------------------------------


class Type:
    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self).__name__)

    def __repr__(self):
        return type(self).__name__.replace("Type", "").lower()


class IntType(Type):
    pass


class FloatType(Type):
    pass


class BoolType(Type):
    pass


class NoneType_(Type):
    """type of the literal ``None`` (the reference leaves it untyped, sdql_ir.py:196-197)."""


class StringType(Type):
    def __init__(self, charCount=None):
        self.charCount = charCount

    def __repr__(self):
        return "string(%s)" % self.charCount


class RecordType(Type):
    def __init__(self, pairs):
        self.typesList = list(pairs)
        self.typesDict = dict(pairs)

    def __eq__(self, other):
        return isinstance(other, RecordType) and self.typesList == other.typesList

    __hash__ = Type.__hash__

    def __repr__(self):
        return "<" + ", ".join("%s: %r" % p for p in self.typesList) + ">"


class DictionaryType(Type):
    def __init__(self, fromType=None, toType=None):
        self.fromType, self.toType = fromType, toType

    def __eq__(self, other):
        return isinstance(other, DictionaryType) and self.fromType == other.fromType and self.toType == other.toType

    __hash__ = Type.__hash__

    def __repr__(self):
        return "{%r -> %r}" % (self.fromType, self.toType)


class VectorType(Type):
    def __init__(self, exprTypes=None):
        self.exprTypes = list(exprTypes or [])

    def __repr__(self):
        return "vector%r" % (self.exprTypes,)


class CompareSymbol(enum.Enum):
    EQ, LT, GT, LTE, GTE, NE = range(1, 7)


class ExtFuncSymbol(enum.Enum):
    StringContains, SubStr, ToStr, ExtractYear, StartsWith, EndsWith, DictSize, FirstIndex = range(1, 9)


# ---------------------------------------------------------------------------------------------
# expressions (ref sdql_ir.py:123-385)
# ---------------------------------------------------------------------------------------------
_ids = itertools.count(1)
_names = itertools.count(1)


def fresh_name(prefix="v"):
    return "%s%d" % (prefix, next(_names))


class Expr:
    fields = ()

    def __init__(self):
        self.id = next(_ids)
        self.lineno = None

    def children(self):
        for f in self.fields:
            v = getattr(self, f)
            if isinstance(v, Expr):
                yield v
            elif isinstance(v, (list, tuple)):
                for e in v:
                    if isinstance(e, Expr):
                        yield e
                    elif isinstance(e, tuple):
                        for x in e:
                            if isinstance(x, Expr):
                                yield x

    def __repr__(self):
        return dump(self)


class ConstantExpr(Expr):
    def __init__(self, value):
        super().__init__()
        self.value = value
        if isinstance(value, bool):
            self.type = BoolType()
        elif isinstance(value, int):
            self.type = IntType()
        elif isinstance(value, float):
            self.type = FloatType()
        elif isinstance(value, str):
            self.type = StringType()
        elif value is None:
            self.type = NoneType_()
        else:
            raise TypeError("constant type not supported: %r" % (value,))


class VarExpr(Expr):
    def __init__(self, name):
        super().__init__()
        self.name = name


class LetExpr(Expr):
    fields = ("varExpr", "valExpr", "bodyExpr")

    def __init__(self, varExpr, valExpr, bodyExpr):
        super().__init__()
        self.varExpr, self.valExpr, self.bodyExpr = varExpr, valExpr, bodyExpr


class SumExpr(Expr):
    """sum over a dictionary / relation (ref sdql_ir.py:223-235).  ``isAssignmentSum`` = the body was wrapped in
    ``unique(..)`` (or joinBuild / non-update joinProbe); ``dictType`` is 'dense_array(N)' after ``dense(N, ..)``."""
    fields = ("varExpr", "dictExpr", "bodyExpr")

    def __init__(self, varExpr, dictExpr, bodyExpr, isAssignmentSum=False, dictType="phmap::flat_hash_map"):
        super().__init__()
        self.varExpr, self.dictExpr, self.bodyExpr = varExpr, dictExpr, bodyExpr
        self.isAssignmentSum, self.dictType = isAssignmentSum, dictType
        self.outputExpr = VarExpr(fresh_name())


class DicConsExpr(Expr):
    fields = ("initialPairs",)

    def __init__(self, initialPairs):
        super().__init__()
        self.initialPairs = list(initialPairs)


class EmptyDicConsExpr(Expr):
    pass


class DicLookupExpr(Expr):
    fields = ("dicExpr", "keyExpr")

    def __init__(self, dicExpr, keyExpr):
        super().__init__()
        self.dicExpr, self.keyExpr = dicExpr, keyExpr


class RecConsExpr(Expr):
    fields = ("initialPairs",)

    def __init__(self, initialPairs):
        super().__init__()
        self.initialPairs = list(initialPairs)  # [(name, Expr)]


class VecConsExpr(Expr):
    fields = ("exprList",)

    def __init__(self, exprList):
        super().__init__()
        self.exprList = list(exprList)


class RecAccessExpr(Expr):
    fields = ("recExpr",)

    def __init__(self, recExpr, name):
        super().__init__()
        self.recExpr, self.name = recExpr, name


class IfExpr(Expr):
    fields = ("condExpr", "thenBodyExpr", "elseBodyExpr")

    def __init__(self, condExpr, thenBodyExpr, elseBodyExpr):
        super().__init__()
        self.condExpr, self.thenBodyExpr, self.elseBodyExpr = condExpr, thenBodyExpr, elseBodyExpr


class _BinExpr(Expr):
    fields = ("op1Expr", "op2Expr")

    def __init__(self, op1Expr, op2Expr):
        super().__init__()
        self.op1Expr, self.op2Expr = op1Expr, op2Expr


class AddExpr(_BinExpr):
    pass


class SubExpr(_BinExpr):
    pass


class MulExpr(_BinExpr):
    pass


class DivExpr(_BinExpr):
    pass


class PromoteExpr(Expr):
    fields = ("bodyExpr",)

    def __init__(self, fromType, toType, bodyExpr):
        super().__init__()
        self.fromType, self.toType, self.bodyExpr = fromType, toType, bodyExpr


class CompareExpr(Expr):
    fields = ("leftExpr", "rightExpr")

    def __init__(self, compareType, leftExpr, rightExpr):
        super().__init__()
        self.compareType, self.leftExpr, self.rightExpr = compareType, leftExpr, rightExpr


class PairAccessExpr(Expr):
    fields = ("pairExpr",)

    def __init__(self, pairExpr, index):
        super().__init__()
        self.pairExpr, self.index = pairExpr, index


class ConcatExpr(Expr):
    fields = ("rec1", "rec2")

    def __init__(self, rec1, rec2):
        super().__init__()
        self.rec1, self.rec2 = rec1, rec2


class ExtFuncExpr(Expr):
    fields = ("inp1", "inp2", "inp3")

    def __init__(self, symbol, inp1, inp2=None, inp3=None):
        super().__init__()
        self.symbol, self.inp1, self.inp2, self.inp3 = symbol, inp1, inp2, inp3


def dump(e, depth=0, maxdepth=6):
    """compact one-line-per-node printer (debugging aid)."""
    pad = "  " * depth
    name = type(e).__name__
    extra = ""
    for a in ("name", "value", "index", "compareType", "symbol", "isAssignmentSum", "dictType"):
        if hasattr(e, a):
            extra += " %s=%r" % (a, getattr(e, a))
    s = "%s%s%s\n" % (pad, name, extra)
    if depth < maxdepth:
        for c in e.children():
            s += dump(c, depth + 1, maxdepth)
    return s
