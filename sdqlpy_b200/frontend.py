"""Python AST -> SDQL IR (py3.12-native front end).

Follows the translation rules of the reference's compiler driver
(/root/reference/src/sdqlpy/lib/sdql_compiler.py:14-370) so that the same decorated query functions
produce the same IR shapes:

  X.sum(lambda p: body)                 -> SumExpr(p, X, body)                       (comp:43-82)
  unique(e) inside a sum body           -> that sum is an assignment sum             (comp:154-156)
  dense(N, e)                           -> dictType 'dense_array(N)'                 (comp:158-160, 75-77)
  X.joinBuild(col, filter, cols)        -> JoinPartitionBuilder                      (comp:162-187, ir:424-441)
  X.joinProbe(idx, col, filter, f, upd) -> JoinProbeBuilder                          (comp:189-210, ir:443-454)
  a and b / a or b                      -> a * b / a + b                             (comp:277-292)
  x in s                                -> StringContains(x, -1, s) == True          (comp:234-240)
  -x                                    -> ConstantExpr(-1) * x                      (comp:267-275)
  return e                              -> LetExpr(VarExpr("out"), e, True)          (comp:340-343)
  function args                         -> VarExpr("db->" + arg + "_dataset")       (comp:31-35, 11-12)

Unlike the reference (which prints IR-building Python text and exec()s it, and needs the py3.8
``ast.Index`` node) this builds IR nodes directly.
"""
import ast

from . import ir
from .ir import (AddExpr, CompareExpr, CompareSymbol, ConcatExpr, ConstantExpr, DicConsExpr, DicLookupExpr, DivExpr,
                 EmptyDicConsExpr, ExtFuncExpr, ExtFuncSymbol, IfExpr, LetExpr, MulExpr, PairAccessExpr, RecAccessExpr,
                 RecConsExpr, SubExpr, SumExpr, VarExpr, VecConsExpr)


class FrontendError(Exception):
    pass


def dataset_var(arg):
    return "db->" + arg + "_dataset"


class _RecAlias:
    """lambda parameter of a joinBuild/joinProbe filter: ``p[0]`` means the record itself (comp:314-327)."""

    def __init__(self, rec):
        self.rec = rec


_CMP = {ast.Eq: CompareSymbol.EQ, ast.NotEq: CompareSymbol.NE, ast.Lt: CompareSymbol.LT, ast.LtE: CompareSymbol.LTE,
        ast.Gt: CompareSymbol.GT, ast.GtE: CompareSymbol.GTE}
_EXT1 = {"extractYear": ExtFuncSymbol.ExtractYear, "dictSize": ExtFuncSymbol.DictSize}
_EXT2 = {"startsWith": ExtFuncSymbol.StartsWith, "endsWith": ExtFuncSymbol.EndsWith,
         "firstIndex": ExtFuncSymbol.FirstIndex}


class Translator:
    def __init__(self, func_node, global_consts=None):
        self.fn = func_node
        self.globals = global_consts or {}
        self.sum_frames = []

    # -- statements ---------------------------------------------------------------------------
    def translate(self):
        env = {}
        self.args = [a.arg for a in self.fn.args.args]
        for a in self.args:
            env[a] = VarExpr(dataset_var(a))
        return self._stmts(list(self.fn.body), env)

    def _stmts(self, body, env):
        if not body:
            raise FrontendError("function %s has no return" % self.fn.name)
        st, rest = body[0], body[1:]
        if isinstance(st, ast.Assign):
            if len(st.targets) != 1 or not isinstance(st.targets[0], ast.Name):
                raise FrontendError("only 'name = expr' assignments are supported (line %d)" % st.lineno)
            name = st.targets[0].id
            var = VarExpr(name)
            val = self.tr(st.value, env)
            env2 = dict(env)
            env2[name] = var
            e = LetExpr(var, val, self._stmts(rest, env2))
            e.lineno = st.lineno
            return e
        if isinstance(st, ast.Return):
            e = LetExpr(VarExpr("out"), self.tr(st.value, env), ConstantExpr(True))
            e.lineno = st.lineno
            return e
        if isinstance(st, ast.Expr) and isinstance(st.value, ast.Constant):
            return self._stmts(rest, env)  # docstring
        raise FrontendError("unsupported statement %s (line %d)" % (type(st).__name__, st.lineno))

    # -- expressions --------------------------------------------------------------------------
    def tr(self, n, env):
        m = getattr(self, "tr_" + type(n).__name__, None)
        if m is None:
            raise FrontendError("unsupported syntax %s (line %s)" % (type(n).__name__, getattr(n, "lineno", "?")))
        e = m(n, env)
        if isinstance(e, ir.Expr) and e.lineno is None:
            e.lineno = getattr(n, "lineno", None)
        return e

    def tr_Constant(self, n, env):
        return ConstantExpr(n.value)

    def tr_Name(self, n, env):
        if n.id in env:
            v = env[n.id]
            if isinstance(v, _RecAlias):
                raise FrontendError("filter parameter '%s' must be used as %s[0] (line %d)" % (n.id, n.id, n.lineno))
            return v
        if n.id in self.globals:
            return ConstantExpr(self.globals[n.id])
        raise FrontendError("unknown name '%s' (line %d)" % (n.id, n.lineno))

    def tr_Attribute(self, n, env):
        return RecAccessExpr(self.tr(n.value, env), n.attr)

    def tr_Subscript(self, n, env):
        sl = n.slice
        if isinstance(n.value, ast.Name) and isinstance(env.get(n.value.id), _RecAlias):
            if isinstance(sl, ast.Constant) and sl.value == 0:
                return env[n.value.id].rec
            if isinstance(sl, ast.Constant) and sl.value == 1:
                return ConstantExpr(True)
        if isinstance(sl, ast.Constant) and isinstance(sl.value, int) and not isinstance(sl.value, bool):
            return PairAccessExpr(self.tr(n.value, env), sl.value)
        return DicLookupExpr(self.tr(n.value, env), self.tr(sl, env))

    def tr_IfExp(self, n, env):
        return IfExpr(self.tr(n.test, env), self.tr(n.body, env), self.tr(n.orelse, env))

    def tr_Compare(self, n, env):
        op = n.ops[0]
        if isinstance(op, ast.In):
            return CompareExpr(CompareSymbol.EQ,
                               ExtFuncExpr(ExtFuncSymbol.StringContains, self.tr(n.left, env), ConstantExpr(-1),
                                           self.tr(n.comparators[0], env)), ConstantExpr(True))
        if type(op) not in _CMP or len(n.ops) != 1:
            raise FrontendError("unsupported comparison (line %d)" % n.lineno)
        return CompareExpr(_CMP[type(op)], self.tr(n.left, env), self.tr(n.comparators[0], env))

    def tr_BoolOp(self, n, env):
        cls = MulExpr if isinstance(n.op, ast.And) else AddExpr
        acc = self.tr(n.values[0], env)
        for v in n.values[1:]:
            acc = cls(acc, self.tr(v, env))
        return acc

    def tr_BinOp(self, n, env):
        cls = {ast.Mult: MulExpr, ast.Sub: SubExpr, ast.Add: AddExpr, ast.Div: DivExpr, ast.FloorDiv: DivExpr}.get(type(n.op))
        if cls is None:
            raise FrontendError("unsupported operator (line %d)" % n.lineno)
        return cls(self.tr(n.left, env), self.tr(n.right, env))

    def tr_UnaryOp(self, n, env):
        if isinstance(n.op, ast.USub):
            if isinstance(n.operand, ast.Constant) and isinstance(n.operand.value, (int, float)):
                return ConstantExpr(-n.operand.value)
            return MulExpr(ConstantExpr(-1), self.tr(n.operand, env))
        if isinstance(n.op, ast.Not):
            return CompareExpr(CompareSymbol.EQ, self.tr(n.operand, env), ConstantExpr(False))
        raise FrontendError("unsupported unary operator (line %d)" % n.lineno)

    def tr_Dict(self, n, env):
        if len(n.keys) == 0:
            return EmptyDicConsExpr()
        return DicConsExpr([(self.tr(n.keys[0], env), self.tr(n.values[0], env))])

    def tr_Set(self, n, env):  # vector({x})
        return self.tr(n.elts[0], env)

    def tr_List(self, n, env):
        return self.tr(n.elts[0], env)

    # -- calls --------------------------------------------------------------------------------
    def _lambda(self, lam, env, bindings):
        if not isinstance(lam, ast.Lambda):
            raise FrontendError("expected a lambda (line %d)" % lam.lineno)
        names = [a.arg for a in lam.args.args]
        if len(names) != len(bindings):
            raise FrontendError("lambda takes %d parameters, expected %d (line %d)" % (len(names), len(bindings), lam.lineno))
        env2 = dict(env)
        env2.update(zip(names, bindings))
        return self.tr(lam.body, env2)

    def tr_Call(self, n, env):
        f = n.func
        if isinstance(f, ast.Attribute):
            if f.attr == "sum":
                return self._sum(n, env)
            if f.attr == "concat":
                return ConcatExpr(self.tr(f.value, env), self.tr(n.args[0], env))
            if f.attr == "joinBuild":
                return self._join_build(n, env)
            if f.attr == "joinProbe":
                return self._join_probe(n, env)
            raise FrontendError("unknown method .%s (line %d)" % (f.attr, n.lineno))
        if not isinstance(f, ast.Name):
            raise FrontendError("unsupported call (line %d)" % n.lineno)
        name = f.id
        if name in _EXT1:
            return ExtFuncExpr(_EXT1[name], self.tr(n.args[0], env))
        if name in _EXT2:
            return ExtFuncExpr(_EXT2[name], self.tr(n.args[0], env), self.tr(n.args[1], env))
        if name == "substr":
            return ExtFuncExpr(ExtFuncSymbol.SubStr, self.tr(n.args[0], env), self.tr(n.args[1], env), self.tr(n.args[2], env))
        if name == "sr_dict":
            return self.tr(n.args[0], env) if n.args else EmptyDicConsExpr()
        if name == "record":
            d = n.args[0]
            if not isinstance(d, ast.Dict):
                raise FrontendError("record(...) needs a dict literal (line %d)" % n.lineno)
            return RecConsExpr([(k.value, self.tr(v, env)) for k, v in zip(d.keys, d.values)])
        if name == "vector":
            return VecConsExpr([self.tr(n.args[0], env)])
        if name == "unique":
            if self.sum_frames:
                self.sum_frames[-1]["unique"] = True
            return self.tr(n.args[0], env)
        if name == "dense":
            if self.sum_frames:
                self.sum_frames[-1]["dense"] = n.args[0].value
            return self.tr(n.args[1], env)
        raise FrontendError("unknown function %s (line %d)" % (name, n.lineno))

    def _sum(self, n, env):
        src = self.tr(n.func.value, env)
        var = VarExpr(ir.fresh_name())
        self.sum_frames.append({"unique": False, "dense": None})
        body = self._lambda(n.args[0], env, [var])
        fr = self.sum_frames.pop()
        dict_type = "phmap::flat_hash_map"
        if len(n.args) < 2 and fr["dense"] is not None:
            dict_type = "dense_array(%d)" % fr["dense"]
        return SumExpr(var, src, body, fr["unique"], dict_type)

    def _join_build(self, n, env):
        src = self.tr(n.func.value, env)
        col = n.args[0].value
        out_cols = [e.value for e in n.args[2].elts] if len(n.args) > 2 else []
        var = VarExpr(ir.fresh_name())
        rec = PairAccessExpr(var, 0)
        cond = self._lambda(n.args[1], env, [_RecAlias(rec)])
        fields = [(c, RecAccessExpr(rec, c)) for c in (out_cols or [col])]
        body = IfExpr(cond, DicConsExpr([(RecAccessExpr(rec, col), RecConsExpr(fields))]), EmptyDicConsExpr())
        return SumExpr(var, src, body, True)

    def _join_probe(self, n, env):
        left = self.tr(n.args[0], env)
        right = self.tr(n.func.value, env)
        col = n.args[1].value
        probe_var = VarExpr(ir.fresh_name())
        var = VarExpr(ir.fresh_name())
        rec = PairAccessExpr(var, 0)
        cond = self._lambda(n.args[2], env, [_RecAlias(rec)])
        hit = DicLookupExpr(probe_var, RecAccessExpr(rec, col))
        self.sum_frames.append({"unique": False, "dense": None})
        out = self._lambda(n.args[3], env, [DicLookupExpr(probe_var, RecAccessExpr(rec, col)), rec])
        self.sum_frames.pop()
        assign = (not n.args[4].value) if len(n.args) > 4 else False
        body = IfExpr(cond, IfExpr(CompareExpr(CompareSymbol.NE, hit, ConstantExpr(None)), out, EmptyDicConsExpr()),
                      EmptyDicConsExpr())
        return LetExpr(probe_var, left, SumExpr(var, right, body, assign))


def parse_module(source):
    """-> (dict fn_name -> (FunctionDef, decorator in_type dict node), module-level constant dict)."""
    tree = ast.parse(source)
    funcs, consts = {}, {}
    for st in tree.body:
        if isinstance(st, ast.FunctionDef):
            for d in st.decorator_list:
                if isinstance(d, ast.Call) and isinstance(d.func, ast.Name) and d.func.id == "sdql_compile":
                    funcs[st.name] = (st, d.args[0] if d.args else None)
        elif isinstance(st, ast.Assign) and len(st.targets) == 1 and isinstance(st.targets[0], ast.Name) \
                and isinstance(st.value, ast.Constant) and isinstance(st.value.value, (int, float, str)):
            consts[st.targets[0].id] = st.value.value
    return funcs, consts


def function_to_ir(func_node, global_consts=None):
    t = Translator(func_node, global_consts)
    return t.translate(), t.args
