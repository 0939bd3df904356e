"""Energy files as the reference's users write them: a `.t` front end.

Reference programs hand `Thallo_ProblemDefine` the *file name* of an energy written in Thallo's
DSL, which is ordinary Lua evaluated in an environment that resolves free names through the
library table of API/src/lib.t (`lib.t:584-590`; the file is loaded by `thallo.t:5954-5975`).
No Lua/Terra exists in this image, so this module is a small interpreter for the Lua subset those
files use -- locals and globals, multiple assignment, functions and closures, tables (1-based array
part + keyed part), numeric and generic `for`, `if`/`while`/`repeat`, method calls, the full
operator table -- evaluated against a DSL namespace `L` (thallo_b200.frontend.dsl.SymbolicL for
the product, oracle.npdsl.NumpyL in tests).  Values that are not Lua values (images, index
variables, AD expressions, vectors) are Python objects of that namespace; Lua operators, calls,
indexing and `obj:method(...)` map to the corresponding Python protocol.

    define = tlang.load("examples/image_warping/image_warping.t")   # -> define(L), like energies/*.py

Library names bound (reference lib.t line in brackets): Dims [43], Inputs [578], Unknown/Array/
Sparse/Param/Image [568-576], Residuals [18], UsePreconditioner [76], Stencil [559], All [55],
And/Or/Not [72-74], Select [192], dot [92], Sqrt [96], normalize [100], length [104], gemv [78],
Rotate2D [138], Rotate3D [123], cross [242], AngleAxisRotatePoint [514], SampledImage [144],
Vector, the rigid-transform helpers [196-512: SelectOnAll, Max, matmul, transpose, PoseToMatrix, rigid_trans, ...],
InBounds/InBoundsExpanded (thallo.t:2091-2112), the comparison constructors and unary math
of ad.t:698-836, the scalar type names (thallo.t / precision.t:3-7), plus Lua's own `ipairs pairs
unpack print assert error type tostring tonumber select math table string.format`.
"""
import math
import re
import sys


class LuaError(Exception):
    pass


# ------------------------------------------------------------------ lexer
_KEYWORDS = {"and", "break", "do", "else", "elseif", "end", "false", "for", "function", "if", "in", "local",
             "nil", "not", "or", "repeat", "return", "then", "true", "until", "while"}
_TOKEN = re.compile(r"""
    (?P<ws>\s+)
  | (?P<lcomment>--\[(?P<lc_eq>=*)\[)
  | (?P<comment>--[^\n]*)
  | (?P<number>0[xX][0-9a-fA-F]+|(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?)
  | (?P<name>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<lstring>\[(?P<ls_eq>=*)\[)
  | (?P<string>"(?:\\.|[^"\\\n])*"|'(?:\\.|[^'\\\n])*')
  | (?P<op>\.\.\.|\.\.|==|~=|<=|>=|[-+*/%^\#<>=(){}\[\];:,.])
""", re.X)
_ESC = {"n": "\n", "t": "\t", "r": "\r", "\\": "\\", '"': '"', "'": "'", "0": "\0", "a": "\a", "b": "\b", "f": "\f", "v": "\v", "\n": "\n"}


def _unescape(s):
    out, i = [], 0
    while i < len(s):
        c = s[i]
        if c == "\\" and i + 1 < len(s):
            n = s[i + 1]
            if n.isdigit():
                j = i + 1
                while j < len(s) and j < i + 4 and s[j].isdigit():
                    j += 1
                out.append(chr(int(s[i + 1:j])))
                i = j
                continue
            out.append(_ESC.get(n, n))
            i += 2
        else:
            out.append(c)
            i += 1
    return "".join(out)


def tokenize(text, chunk="?"):
    toks, pos, line = [], 0, 1
    n = len(text)
    while pos < n:
        m = _TOKEN.match(text, pos)
        if not m:
            raise LuaError("%s:%d: unexpected character %r" % (chunk, line, text[pos]))
        kind = m.lastgroup
        if kind in ("lcomment", "lstring"):
            eq = m.group("lc_eq") if kind == "lcomment" else m.group("ls_eq")
            close = "]" + eq + "]"
            end = text.find(close, m.end())
            if end < 0:
                raise LuaError("%s:%d: unfinished long %s" % (chunk, line, "comment" if kind == "lcomment" else "string"))
            body = text[m.end():end]
            if kind == "lstring":
                toks.append(("string", body[1:] if body.startswith("\n") else body, line))
            line += text.count("\n", pos, end + len(close))
            pos = end + len(close)
            continue
        val = m.group(kind)
        if kind == "number":
            if val[:2] in ("0x", "0X"):
                toks.append(("number", int(val, 16), line))
            elif re.fullmatch(r"\d+", val):
                toks.append(("number", int(val), line))
            else:
                toks.append(("number", float(val), line))
        elif kind == "name":
            toks.append(("kw" if val in _KEYWORDS else "name", val, line))
        elif kind == "string":
            toks.append(("string", _unescape(val[1:-1]), line))
        elif kind == "op":
            toks.append(("op", val, line))
        line += val.count("\n")
        pos = m.end()
    toks.append(("eof", None, line))
    return toks


# ------------------------------------------------------------------ parser (AST = nested tuples, first element the node kind)
_BINPRI = {"or": (1, 1), "and": (2, 2), "<": (3, 3), ">": (3, 3), "<=": (3, 3), ">=": (3, 3), "~=": (3, 3), "==": (3, 3),
           "..": (5, 4), "+": (6, 6), "-": (6, 6), "*": (7, 7), "/": (7, 7), "%": (7, 7), "^": (10, 9)}
_UNARY_PRI = 8


class Parser:
    def __init__(self, text, chunk="?"):
        self.t, self.i, self.chunk = tokenize(text, chunk), 0, chunk

    def err(self, msg):
        raise LuaError("%s:%d: %s near %r" % (self.chunk, self.t[self.i][2], msg, self.t[self.i][1]))

    def peek(self, k=0):
        return self.t[self.i + k]

    def check(self, kind, val=None):
        tk = self.t[self.i]
        return tk[0] == kind and (val is None or tk[1] == val)

    def accept(self, kind, val=None):
        if self.check(kind, val):
            self.i += 1
            return True
        return False

    def expect(self, kind, val=None):
        if not self.check(kind, val):
            self.err("expected %s" % (val or kind))
        tk = self.t[self.i]
        self.i += 1
        return tk[1]

    def block_end(self):
        tk = self.t[self.i]
        return tk[0] == "eof" or (tk[0] == "kw" and tk[1] in ("end", "else", "elseif", "until"))

    def block(self):
        stats = []
        while not self.block_end():
            if self.accept("op", ";"):
                continue
            if self.check("kw", "return"):
                line = self.t[self.i][2]
                self.i += 1
                exps = []
                if not self.block_end() and not self.check("op", ";"):
                    exps = self.explist()
                self.accept("op", ";")
                stats.append(("return", exps, line))
                break
            stats.append(self.statement())
        return stats

    def statement(self):
        tk = self.t[self.i]
        line = tk[2]
        if tk[0] == "kw":
            k = tk[1]
            if k == "local":
                self.i += 1
                if self.accept("kw", "function"):
                    name = self.expect("name")
                    return ("localfunction", name, self.funcbody(), line)
                names = [self.expect("name")]
                while self.accept("op", ","):
                    names.append(self.expect("name"))
                exps = self.explist() if self.accept("op", "=") else []
                return ("local", names, exps, line)
            if k == "function":
                self.i += 1
                target = ("name", self.expect("name"), line)
                method = False
                while True:
                    if self.accept("op", "."):
                        target = ("index", target, ("const", self.expect("name")), line)
                    elif self.accept("op", ":"):
                        target = ("index", target, ("const", self.expect("name")), line)
                        method = True
                        break
                    else:
                        break
                return ("assign", [target], [self.funcbody(method)], line)
            if k == "if":
                self.i += 1
                clauses, orelse = [], None
                cond = self.exp()
                self.expect("kw", "then")
                clauses.append((cond, self.block()))
                while True:
                    if self.accept("kw", "elseif"):
                        cond = self.exp()
                        self.expect("kw", "then")
                        clauses.append((cond, self.block()))
                    elif self.accept("kw", "else"):
                        orelse = self.block()
                        self.expect("kw", "end")
                        break
                    else:
                        self.expect("kw", "end")
                        break
                return ("if", clauses, orelse, line)
            if k == "for":
                self.i += 1
                n1 = self.expect("name")
                if self.accept("op", "="):
                    a = self.exp()
                    self.expect("op", ",")
                    b = self.exp()
                    c = self.exp() if self.accept("op", ",") else None
                    self.expect("kw", "do")
                    body = self.block()
                    self.expect("kw", "end")
                    return ("fornum", n1, a, b, c, body, line)
                names = [n1]
                while self.accept("op", ","):
                    names.append(self.expect("name"))
                self.expect("kw", "in")
                exps = self.explist()
                self.expect("kw", "do")
                body = self.block()
                self.expect("kw", "end")
                return ("forin", names, exps, body, line)
            if k == "while":
                self.i += 1
                cond = self.exp()
                self.expect("kw", "do")
                body = self.block()
                self.expect("kw", "end")
                return ("while", cond, body, line)
            if k == "repeat":
                self.i += 1
                body = self.block()
                self.expect("kw", "until")
                return ("repeat", body, self.exp(), line)
            if k == "do":
                self.i += 1
                body = self.block()
                self.expect("kw", "end")
                return ("do", body, line)
            if k == "break":
                self.i += 1
                return ("break", line)
            self.err("unexpected keyword")
        e = self.suffixedexp()
        if self.check("op", "=") or self.check("op", ","):
            targets = [e]
            while self.accept("op", ","):
                targets.append(self.suffixedexp())
            self.expect("op", "=")
            for t in targets:
                if t[0] not in ("name", "index"):
                    self.err("cannot assign to this expression")
            return ("assign", targets, self.explist(), line)
        if e[0] not in ("call", "method"):
            self.err("syntax error (expression is not a statement)")
        return ("callstat", e, line)

    def funcbody(self, method=False):
        line = self.t[self.i][2]
        self.expect("op", "(")
        params, vararg = (["self"] if method else []), False
        if not self.check("op", ")"):
            while True:
                if self.accept("op", "..."):
                    vararg = True
                    break
                params.append(self.expect("name"))
                if not self.accept("op", ","):
                    break
        self.expect("op", ")")
        body = self.block()
        self.expect("kw", "end")
        return ("function", params, vararg, body, line)

    def explist(self):
        exps = [self.exp()]
        while self.accept("op", ","):
            exps.append(self.exp())
        return exps

    def primaryexp(self):
        tk = self.t[self.i]
        if tk[0] == "name":
            self.i += 1
            return ("name", tk[1], tk[2])
        if self.accept("op", "("):
            e = self.exp()
            self.expect("op", ")")
            return ("paren", e)
        self.err("unexpected symbol")

    def suffixedexp(self):
        e = self.primaryexp()
        while True:
            tk = self.t[self.i]
            line = tk[2]
            if tk[0] == "op" and tk[1] == ".":
                self.i += 1
                e = ("index", e, ("const", self.expect("name")), line)
            elif tk[0] == "op" and tk[1] == "[":
                self.i += 1
                k = self.exp()
                self.expect("op", "]")
                e = ("index", e, k, line)
            elif tk[0] == "op" and tk[1] == ":":
                self.i += 1
                name = self.expect("name")
                e = ("method", e, name, self.callargs(), line)
            elif (tk[0] == "op" and tk[1] in ("(", "{")) or tk[0] == "string":
                e = ("call", e, self.callargs(), line)
            else:
                return e

    def callargs(self):
        tk = self.t[self.i]
        if tk[0] == "string":
            self.i += 1
            return [("const", tk[1])]
        if self.check("op", "{"):
            return [self.table()]
        self.expect("op", "(")
        if self.accept("op", ")"):
            return []
        args = self.explist()
        self.expect("op", ")")
        return args

    def table(self):
        line = self.t[self.i][2]
        self.expect("op", "{")
        items = []                                  # ("pos", exp) | ("key", keyexp, exp)
        while not self.check("op", "}"):
            if self.check("name") and self.peek(1)[0] == "op" and self.peek(1)[1] == "=":
                k = self.expect("name")
                self.i += 1
                items.append(("key", ("const", k), self.exp()))
            elif self.check("op", "["):
                self.i += 1
                k = self.exp()
                self.expect("op", "]")
                self.expect("op", "=")
                items.append(("key", k, self.exp()))
            else:
                items.append(("pos", self.exp()))
            if not (self.accept("op", ",") or self.accept("op", ";")):
                break
        self.expect("op", "}")
        return ("table", items, line)

    def simpleexp(self):
        tk = self.t[self.i]
        if tk[0] == "number" or tk[0] == "string":
            self.i += 1
            return ("const", tk[1])
        if tk[0] == "kw":
            if tk[1] == "nil":
                self.i += 1
                return ("const", None)
            if tk[1] == "true":
                self.i += 1
                return ("const", True)
            if tk[1] == "false":
                self.i += 1
                return ("const", False)
            if tk[1] == "function":
                self.i += 1
                return self.funcbody()
        if tk[0] == "op":
            if tk[1] == "...":
                self.i += 1
                return ("vararg",)
            if tk[1] == "{":
                return self.table()
        return self.suffixedexp()

    def exp(self, limit=0):
        tk = self.t[self.i]
        if (tk[0] == "kw" and tk[1] == "not") or (tk[0] == "op" and tk[1] in ("-", "#")):
            self.i += 1
            left = ("unop", tk[1], self.exp(_UNARY_PRI), tk[2])
        else:
            left = self.simpleexp()
        while True:
            tk = self.t[self.i]
            op = tk[1] if tk[0] in ("op", "kw") else None
            pri = _BINPRI.get(op)
            if pri is None or pri[0] <= limit:
                return left
            self.i += 1
            right = self.exp(pri[1])
            left = ("binop", op, left, right, tk[2])

    def chunk_(self):
        b = self.block()
        if not self.check("eof"):
            self.err("unexpected token")
        return b


def parse(text, chunk="?"):
    return Parser(text, chunk).chunk_()


# ------------------------------------------------------------------ values
class LuaTable:
    """Lua table: insertion-ordered keyed part; integer keys 1..n form the array part."""

    def __init__(self):
        self.d = {}

    @staticmethod
    def _key(k):
        if isinstance(k, float) and k.is_integer():
            return int(k)
        return k

    def get(self, k):
        return self.d.get(self._key(k))

    def set(self, k, v):
        k = self._key(k)
        if k is None:
            raise LuaError("table index is nil")
        if v is None:
            self.d.pop(k, None)
        else:
            self.d[k] = v

    def length(self):
        n = 0
        while (n + 1) in self.d:
            n += 1
        return n

    def array(self):
        return [self.d[i] for i in range(1, self.length() + 1)]

    def keyed(self):
        n = self.length()
        return {k: v for k, v in self.d.items() if not (isinstance(k, int) and 1 <= k <= n)}

    def is_array(self):
        return self.length() == len(self.d)


def table_of(seq=(), **kw):
    t = LuaTable()
    for i, v in enumerate(seq):
        t.set(i + 1, v)
    for k, v in kw.items():
        t.set(k, v)
    return t


def to_python(v):
    """Lua value -> Python value for DSL calls: array tables become lists (recursively)."""
    if isinstance(v, LuaTable):
        if v.is_array():
            return [to_python(x) for x in v.array()]
        return {k: to_python(x) for k, x in v.d.items()}
    return v


class Multi(list):
    """Multiple return values of a builtin."""


class _Break(Exception):
    pass


class _Return(Exception):
    def __init__(self, values):
        self.values = values


class Scope:
    __slots__ = ("vars", "parent")

    def __init__(self, parent=None):
        self.vars, self.parent = {}, parent

    def find(self, name):
        s = self
        while s is not None:
            if name in s.vars:
                return s
            s = s.parent
        return None


class LuaFunction:
    def __init__(self, interp, node, scope, name="?"):
        self.interp, self.node, self.scope, self.name = interp, node, scope, name

    def __call__(self, *args):          # callable from Python (DSL callbacks): first return value
        r = self.interp.call(self, list(args))
        return r[0] if r else None


# ------------------------------------------------------------------ interpreter
def _truthy(v):
    return v is not None and v is not False


def _isnum(v):
    return isinstance(v, (int, float)) and not isinstance(v, bool)


class Interpreter:
    def __init__(self, globals_, chunk="?"):
        self.G, self.chunk = globals_, chunk
        self.line = 0

    def error(self, msg):
        raise LuaError("%s:%d: %s" % (self.chunk, self.line, msg))

    # ---- calls
    def call(self, f, args):
        if isinstance(f, LuaFunction):
            _, params, vararg, body, _ = f.node
            sc = Scope(f.scope)
            for i, p in enumerate(params):
                sc.vars[p] = args[i] if i < len(args) else None
            if vararg:
                sc.vars["..."] = list(args[len(params):])
            try:
                self.exec_block(body, sc)
            except _Return as r:
                return r.values
            return []
        if isinstance(f, LuaTable):
            self.error("attempt to call a table value")
        if f is None:
            self.error("attempt to call a nil value")
        if not callable(f):
            self.error("attempt to call a %s value" % type(f).__name__)
        r = f(*args)
        if isinstance(r, Multi):
            return list(r)
        return [r]

    def index(self, obj, key):
        if isinstance(obj, LuaTable):
            return obj.get(key)
        if obj is None:
            self.error("attempt to index a nil value (key %r)" % (key,))
        if isinstance(key, str):
            try:
                return getattr(obj, key)
            except AttributeError:
                if isinstance(obj, dict):
                    return obj.get(key)
                self.error("%s has no field '%s'" % (type(obj).__name__, key))
        if isinstance(obj, (list, tuple)):
            return obj[int(key) - 1] if 1 <= key <= len(obj) else None
        return obj[int(key) if isinstance(key, float) and key.is_integer() else key]

    def setindex(self, obj, key, val):
        if isinstance(obj, LuaTable):
            obj.set(key, val)
        elif isinstance(key, str):
            setattr(obj, key, val)
        else:
            obj[key] = val

    # ---- expressions
    def eval_multi(self, e, sc):
        k = e[0]
        if k == "call":
            self.line = e[3]
            f = self.eval(e[1], sc)
            args = self.eval_list(e[2], sc)
            line = e[3]
            if f is None:
                self.line = line
                what = ("global '%s'" % e[1][1]) if e[1][0] == "name" else ("field '%s'" % e[1][2][1]) if e[1][0] == "index" and e[1][2][0] == "const" else "expression"
                self.error("attempt to call a nil value (%s)" % what)
            try:
                return self.call(f, args)
            except (LuaError, _Return, _Break):
                raise
            except Exception as ex:                 # errors raised by the DSL namespace: add the .t position
                raise LuaError("%s:%d: %s: %s" % (self.chunk, line, type(ex).__name__, ex)) from ex
        if k == "method":
            self.line = e[4]
            obj = self.eval(e[1], sc)
            args = self.eval_list(e[3], sc)
            line = e[4]
            if isinstance(obj, LuaTable):
                f = obj.get(e[2])
                args = [obj] + args
            elif isinstance(obj, str):                      # s:format(...), s:rep(n): methods of the string library
                f = self.G["string"].get(e[2]) if isinstance(self.G.get("string"), LuaTable) else None
                if f is None:
                    self.error("string has no method '%s'" % e[2])
                args = [obj] + args
            else:
                if obj is None:
                    self.error("attempt to call method '%s' of a nil value" % e[2])
                f = getattr(obj, e[2], None)
                if f is None:
                    self.error("%s has no method '%s'" % (type(obj).__name__, e[2]))
            try:
                return self.call(f, args)
            except (LuaError, _Return, _Break):
                raise
            except Exception as ex:
                raise LuaError("%s:%d: %s: %s" % (self.chunk, line, type(ex).__name__, ex)) from ex
        if k == "vararg":
            s = sc.find("...")
            if s is None:
                self.error("cannot use '...' outside a vararg function")
            return list(s.vars["..."])
        return [self.eval(e, sc)]

    def eval_list(self, exps, sc):
        out = []
        for i, e in enumerate(exps):
            if i == len(exps) - 1 and e[0] in ("call", "method", "vararg"):
                out.extend(self.eval_multi(e, sc))
            else:
                out.append(self.eval(e, sc))
        return out

    def eval(self, e, sc):
        k = e[0]
        if k == "const":
            return e[1]
        if k == "name":
            s = sc.find(e[1])
            if s is not None:
                return s.vars[e[1]]
            return self.G.get(e[1])
        if k == "paren":
            return self.eval(e[1], sc)
        if k in ("call", "method", "vararg"):
            r = self.eval_multi(e, sc)
            return r[0] if r else None
        if k == "index":
            self.line = e[3]
            return self.index(self.eval(e[1], sc), self.eval(e[2], sc))
        if k == "function":
            return LuaFunction(self, e, sc)
        if k == "table":
            t = LuaTable()
            n = 0
            items = e[1]
            for i, it in enumerate(items):
                if it[0] == "key":
                    t.set(self.eval(it[1], sc), self.eval(it[2], sc))
                elif i == len(items) - 1 and it[1][0] in ("call", "method", "vararg"):
                    for v in self.eval_multi(it[1], sc):
                        n += 1
                        t.set(n, v)
                else:
                    n += 1
                    t.set(n, self.eval(it[1], sc))
            return t
        if k == "unop":
            self.line = e[3]
            v = self.eval(e[2], sc)
            if e[1] == "not":
                return not _truthy(v)
            if e[1] == "-":
                if v is None:
                    self.error("attempt to perform arithmetic on a nil value")
                return -v
            if isinstance(v, LuaTable):
                return v.length()
            return len(v)
        if k == "binop":
            op = e[1]
            if op == "and":
                a = self.eval(e[2], sc)
                return self.eval(e[3], sc) if _truthy(a) else a
            if op == "or":
                a = self.eval(e[2], sc)
                return a if _truthy(a) else self.eval(e[3], sc)
            a, b = self.eval(e[2], sc), self.eval(e[3], sc)
            self.line = e[4]
            return self.binop(op, a, b)
        self.error("cannot evaluate node %s" % k)

    def binop(self, op, a, b):
        if op == "==":
            return a is b or (type(a) in (int, float, str, bool) and type(b) in (int, float, str, bool) and a == b)
        if op == "~=":
            return not self.binop("==", a, b)
        if op == "..":
            return self.tostring(a) + self.tostring(b)
        if a is None or b is None or isinstance(a, (bool, LuaTable)) or isinstance(b, (bool, LuaTable)):
            self.error("attempt to perform '%s' on a %s and a %s value" % (op, self.typename(a), self.typename(b)))
        try:
            if op == "+":
                return a + b
            if op == "-":
                return a - b
            if op == "*":
                return a * b
            if op == "/":
                if _isnum(a) and _isnum(b):
                    return a / b if b != 0 else (math.nan if a == 0 else math.copysign(math.inf, a))
                return a / b
            if op == "%":
                return a - math.floor(a / b) * b
            if op == "^":
                if _isnum(a) and _isnum(b):
                    return float(a) ** b
                return a ** b
            if op == "<":
                return a < b
            if op == "<=":
                return a <= b
            if op == ">":
                return a > b
            if op == ">=":
                return a >= b
        except TypeError as ex:
            self.error("attempt to perform '%s' on %s and %s (%s)" % (op, type(a).__name__, type(b).__name__, ex))
        self.error("unknown operator %s" % op)

    @staticmethod
    def typename(v):
        if v is None:
            return "nil"
        if isinstance(v, bool):
            return "boolean"
        if _isnum(v):
            return "number"
        if isinstance(v, str):
            return "string"
        if isinstance(v, LuaTable):
            return "table"
        import types
        if isinstance(v, (LuaFunction, types.FunctionType, types.BuiltinFunctionType, types.MethodType)):
            return "function"
        return "userdata"                    # DSL objects: images, index variables, expressions, vectors

    @staticmethod
    def tostring(v):
        if v is None:
            return "nil"
        if v is True:
            return "true"
        if v is False:
            return "false"
        if isinstance(v, float) and v.is_integer() and abs(v) < 1e15:
            return str(int(v))
        return str(v)

    # ---- statements
    def exec_block(self, stats, sc):
        for s in stats:
            self.exec(s, sc)

    def assign(self, target, val, sc):
        if target[0] == "name":
            s = sc.find(target[1])
            if s is not None:
                s.vars[target[1]] = val
            else:
                if isinstance(val, LuaFunction) and val.name == "?":
                    val.name = target[1]
                self.G[target[1]] = val
        else:
            self.line = target[3]
            self.setindex(self.eval(target[1], sc), self.eval(target[2], sc), val)

    def exec(self, s, sc):
        k = s[0]
        if k == "local":
            vals = self.eval_list(s[2], sc)
            for i, n in enumerate(s[1]):
                sc.vars[n] = vals[i] if i < len(vals) else None
        elif k == "assign":
            vals = self.eval_list(s[2], sc)
            for i, t in enumerate(s[1]):
                self.assign(t, vals[i] if i < len(vals) else None, sc)
        elif k == "callstat":
            self.eval_multi(s[1], sc)
        elif k == "localfunction":
            sc.vars[s[1]] = None
            sc.vars[s[1]] = LuaFunction(self, s[2], sc, s[1])
        elif k == "if":
            for cond, body in s[1]:
                if _truthy(self.eval(cond, sc)):
                    self.exec_block(body, Scope(sc))
                    return
            if s[2] is not None:
                self.exec_block(s[2], Scope(sc))
        elif k == "fornum":
            a, b = self.eval(s[2], sc), self.eval(s[3], sc)
            c = self.eval(s[4], sc) if s[4] is not None else 1
            if not (_isnum(a) and _isnum(b) and _isnum(c)) or c == 0:
                self.line = s[6]
                self.error("'for' bounds and step must be numbers (step non-zero)")
            i = a
            try:
                while (c > 0 and i <= b) or (c < 0 and i >= b):
                    inner = Scope(sc)
                    inner.vars[s[1]] = i
                    self.exec_block(s[5], inner)
                    i += c
            except _Break:
                pass
        elif k == "forin":
            vals = self.eval_list(s[2], sc)
            f = vals[0] if vals else None
            state = vals[1] if len(vals) > 1 else None
            ctrl = vals[2] if len(vals) > 2 else None
            try:
                while True:
                    self.line = s[4]
                    r = self.call(f, [state, ctrl])
                    if not r or r[0] is None:
                        break
                    ctrl = r[0]
                    inner = Scope(sc)
                    for i, n in enumerate(s[1]):
                        inner.vars[n] = r[i] if i < len(r) else None
                    self.exec_block(s[3], inner)
            except _Break:
                pass
        elif k == "while":
            try:
                while _truthy(self.eval(s[1], sc)):
                    self.exec_block(s[2], Scope(sc))
            except _Break:
                pass
        elif k == "repeat":
            try:
                while True:
                    inner = Scope(sc)
                    self.exec_block(s[1], inner)
                    if _truthy(self.eval(s[2], inner)):
                        break
            except _Break:
                pass
        elif k == "do":
            self.exec_block(s[1], Scope(sc))
        elif k == "return":
            raise _Return(self.eval_list(s[1], sc))
        elif k == "break":
            raise _Break()
        else:
            self.error("cannot execute node %s" % k)

    def run(self, stats):
        try:
            self.exec_block(stats, Scope())
        except _Return as r:
            return r.values
        except _Break:
            self.error("'break' outside a loop")
        return []


# ------------------------------------------------------------------ Lua standard library subset
def _base_globals(out):
    def ipairs(t):
        seq = t.array() if isinstance(t, LuaTable) else list(t)

        def it(_, i):
            return Multi([i + 1, seq[i]]) if i < len(seq) else Multi([None])
        return Multi([it, t, 0])

    def pairs(t):
        items = list(t.d.items()) if isinstance(t, LuaTable) else list(enumerate(t, 1))
        pos = [0]

        def it(_, __):
            if pos[0] >= len(items):
                return Multi([None])
            pos[0] += 1
            return Multi(list(items[pos[0] - 1]))
        return Multi([it, t, None])

    def unpack(t, i=1, j=None):
        seq = t.array() if isinstance(t, LuaTable) else list(t)
        return Multi(seq[int(i) - 1:(len(seq) if j is None else int(j))])

    def lua_print(*a):
        out.write("\t".join(Interpreter.tostring(x) for x in a) + "\n")

    def lua_assert(v=None, msg="assertion failed!", *rest):
        if not _truthy(v):
            raise LuaError(str(msg))
        return Multi([v, msg] + list(rest)) if rest or msg != "assertion failed!" else v

    def lua_error(msg="error", level=1):
        raise LuaError(Interpreter.tostring(msg))

    def lua_select(n, *a):
        if n == "#":
            return len(a)
        return Multi(list(a[int(n) - 1:]))

    def tonumber(v, base=None):
        try:
            if isinstance(v, str):
                return int(v, int(base)) if base else (int(v) if re.fullmatch(r"\s*-?\d+\s*", v) else float(v))
            return v if _isnum(v) else None
        except ValueError:
            return None

    def tinsert(t, *a):
        if len(a) == 1:
            t.set(t.length() + 1, a[0])
        else:
            pos, v = int(a[0]), a[1]
            for i in range(t.length(), pos - 1, -1):
                t.set(i + 1, t.get(i))
            t.set(pos, v)

    def tremove(t, pos=None):
        n = t.length()
        if n == 0:
            return None
        pos = n if pos is None else int(pos)
        v = t.get(pos)
        for i in range(pos, n):
            t.set(i, t.get(i + 1))
        t.set(n, None)
        return v

    def sformat(fmt, *a):
        fmt = re.sub(r"%(\d*)i", r"%\1d", fmt)
        return fmt % tuple(a)

    m = table_of(pi=math.pi, huge=math.inf, sqrt=math.sqrt, sin=math.sin, cos=math.cos, tan=math.tan, exp=math.exp,
                 log=math.log, abs=abs, floor=lambda x: int(math.floor(x)), ceil=lambda x: int(math.ceil(x)),
                 pow=lambda a, b: float(a) ** b, fmod=math.fmod, atan=math.atan, atan2=math.atan2, acos=math.acos,
                 asin=math.asin, max=lambda *a: max(a), min=lambda *a: min(a))
    return {
        "ipairs": ipairs, "pairs": pairs, "unpack": unpack, "print": lua_print, "assert": lua_assert, "error": lua_error,
        "select": lua_select, "tonumber": tonumber, "tostring": Interpreter.tostring, "type": Interpreter.typename,
        "math": m, "table": table_of(insert=tinsert, remove=tremove, unpack=unpack,
                                     concat=lambda t, sep="": sep.join(Interpreter.tostring(x) for x in t.array())),
        "string": table_of(format=sformat, rep=lambda s, n: s * int(n), len=len, upper=str.upper, lower=str.lower),
    }


# ------------------------------------------------------------------ the DSL library (reference API/src/lib.t)
def _dsl_globals(L, G):
    from energies import _lib

    def py(v):
        return to_python(v)

    def vec(v):
        """Accept a Lua table of components where the library expects a vector."""
        return L.Vector(*v.array()) if isinstance(v, LuaTable) else v

    def Dims(*names):                                            # lib.t:43-49
        d = L.Dims(*names)
        return Multi(d) if isinstance(d, (list, tuple)) else d

    def decl(kind):
        def f(*args):                                            # lib.t:568-571: recorded, created by Inputs
            return ("decl", kind, args)
        return f

    def Inputs(tbl):                                             # lib.t:578-582
        kw = {}
        for name, d in tbl.d.items():
            if not (isinstance(d, tuple) and d and d[0] == "decl"):
                raise LuaError("Inputs{}: entry '%s' is not an Unknown/Array/Sparse/Param declaration" % (name,))
            _, kind, a = d
            if kind in ("Unknown", "Array"):
                if isinstance(a[0], LuaTable):                   # type omitted: (dims, idx) (thallo.t:1612)
                    a = (L.float,) + tuple(a)
                if a[0] is None:
                    raise LuaError("Inputs{}: unknown scalar type for '%s'" % name)
                kw[name] = getattr(L, kind)(a[0], py(a[1]), _pidx(a[2]))
            elif kind == "Sparse":
                kw[name] = L.Sparse(py(a[0]), py(a[1]), _pidx(a[2]))
            else:
                kw[name] = L.Param(a[0], _pidx(a[1]))
        ns = L.Inputs(**kw)
        for name in kw:
            G[name] = getattr(ns, name)

    def _pidx(i):
        return int(i) if _isnum(i) else i

    def Residuals(tbl):                                          # lib.t:18-35
        kw = {}
        for name, v in tbl.d.items():
            if not isinstance(name, str):
                raise LuaError("Residuals{}: residual groups must be named")
            kw[name] = v.array() if isinstance(v, LuaTable) else v
        return L.Residuals(**kw)

    def Stencil(lst):                                            # lib.t:559-566
        rows = [r.array() if isinstance(r, LuaTable) else [r] for r in lst.array()]
        pos = [0]

        def it(*_):
            if pos[0] >= len(rows):
                return Multi([None])
            pos[0] += 1
            return Multi(rows[pos[0] - 1])
        return it

    def All(v):                                                  # lib.t:55-61
        return _lib.All(L, vec(v))

    def dot(a, b):                                               # lib.t:92-94
        return _lib.dot(L, vec(a), vec(b))

    def normalize(v):                                            # lib.t:100-102
        v = vec(v)
        return v / L.sqrt(_lib.dot(L, v, v))

    def length(a, b):                                            # lib.t:104-107
        d = vec(a) - vec(b)
        return L.sqrt(_lib.dot(L, d, d))

    def gemv(m, v):                                              # lib.t:78-90
        return _lib.gemv(L, m.array() if isinstance(m, LuaTable) else list(m), vec(v))

    def Vector(*c):
        return L.Vector(*c)

    def SampledImage(im, dx=None, dy=None):                      # lib.t:144
        return L.SampledImage(im, dx, dy)

    def unary(name):
        f = getattr(L, name)
        mf = getattr(math, name if name != "abs" else "fabs")
        return lambda x: mf(x) if _isnum(x) else f(x)

    env = {
        "Dim": lambda name, idx: L.Dim(name, int(idx)),
        "Dims": Dims, "Inputs": Inputs, "Residuals": Residuals, "Stencil": Stencil,
        "Unknown": decl("Unknown"), "Array": decl("Array"), "Image": decl("Array"), "Sparse": decl("Sparse"),
        "Param": decl("Param"),
        "UsePreconditioner": L.UsePreconditioner,
        "All": All, "dot": dot, "normalize": normalize, "length": length, "gemv": gemv, "Vector": Vector,
        "SampledImage": SampledImage, "Sqrt": unary("sqrt"),
        "Rotate2D": lambda a, v: _lib.Rotate2D(L, a, vec(v)),
        "Rotate3D": lambda a, v: _lib.Rotate3D(L, vec(a), vec(v)),
        "cross": lambda a, b: _lib.cross(L, vec(a), vec(b)),
        "AngleAxisRotatePoint": lambda a, p: _lib.AngleAxisRotatePoint(L, vec(a), vec(p)),
        "SelectOnAll": lambda ps, v, d: _lib.SelectOnAll(L, ps.array() if isinstance(ps, LuaTable) else list(ps), v, d),
        "Max": lambda a, b: _lib.Max(L, a, b),
        "matmul": lambda a, b: _lib.matmul(L, vec(a), vec(b)),
        "transpose": lambda m: _lib.transpose(L, vec(m)),
        "Matrix4": lambda *c: Vector(*c), "Vec4": lambda *c: Vector(*c), "Vec3": lambda v: Vector(v[0], v[1], v[2]),
        "rotationFromMat4": lambda t: _lib.rotationFromMat4(L, vec(t)),
        "translationFromMat4": lambda t: _lib.translationFromMat4(L, vec(t)),
        "RotationMatrixAndTranslationToMat4": lambda r, t: _lib.RotationMatrixAndTranslationToMat4(L, vec(r), vec(t)),
        "Mat4ToRigidTransform": lambda m: _lib.Mat4ToRigidTransform(L, vec(m)),
        "RigidTransformToMat4": lambda m: _lib.RigidTransformToMat4(L, vec(m)),
        "InvertRigidTransform": lambda m: _lib.InvertRigidTransform(L, vec(m)),
        "CameraToDepth": lambda fx, fy, cx, cy, pos: _lib.CameraToDepth(L, fx, fy, cx, cy, vec(pos)),
        "RodriguesSO3Exp": lambda w, A, B: _lib.RodriguesSO3Exp(L, vec(w), A, B),
        "PoseToMatrix": lambda r, t: _lib.PoseToMatrix(L, vec(r), vec(t)),
        "rigid_trans": lambda M, v: _lib.rigid_trans(L, vec(M), vec(v)),
        "Select": L.Select, "InBounds": L.InBounds, "InBoundsExpanded": L.InBoundsExpanded,
        "And": L.And, "Or": L.Or, "Not": L.Not,
        "inf": math.inf,
    }
    for name in ("eq", "neq", "less", "greater", "lesseq", "greatereq"):
        env[name] = getattr(L, name)
    for name in ("sqrt", "sin", "cos", "tan", "exp", "log", "abs"):
        if hasattr(L, name):
            env[name] = unary(name)
    # scalar types: `float` is C float and `thallo_float` the solver's scalar type (precision.t:3-7); the DSL
    # namespaces carry one real type, so both names map to it (like energies/*.py)
    for n, attr in (("", "float"), ("2", "float2"), ("3", "float3"), ("4", "float4"), ("6", "float6"), ("9", "float9")):
        t = getattr(L, attr, None)
        if t is not None:
            env["float" + n] = env["thallo_float" + n] = env["double" + n] = t
    if getattr(L, "float9", None) is not None:
        env["thallo_mat3f"] = env["mat3f"] = L.float9               # 3x3 matrix unknowns are nine packed scalars
    for n in ("uint8", "int"):
        if hasattr(L, n):
            env[n] = getattr(L, n)
    env["ad"] = table_of(Vector=Vector, select=L.Select, less=L.less, greater=L.greater, lesseq=L.lesseq, greatereq=L.greatereq,
                         eq=L.eq, sqrt=env["sqrt"], sin=env["sin"], cos=env["cos"], toexp=lambda x: x)
    env["uchar"] = env.get("uint8")
    env["int32"] = env.get("int")
    return env


def make_define(text, chunk="energy.t", out=None):
    """Compile `.t` energy text into a `define(L)` callable usable wherever energies/*.py's are."""
    ast = parse(text, chunk)

    def define(L):
        G = _base_globals(out or sys.stderr)
        G.update(_dsl_globals(L, G))
        G["_G"] = G
        Interpreter(G, chunk).run(ast)
        res = getattr(L, "residuals", None)
        if res is None:
            raise LuaError("%s: energy did not call Residuals{}" % chunk)
        return res
    define.__name__ = "define_" + re.sub(r"\W", "_", chunk)
    return define


def load(path, out=None):
    with open(path) as f:
        return make_define(f.read(), path, out)
