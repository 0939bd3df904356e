"""Symbolic expression DAG with automatic differentiation.

Plays the role of reference API/src/ad.t: hash-consed `Const / Var / Apply` nodes
(ad.t:137-264), the derivative rule table (ad.t:698-836) and light algebraic
simplification (constant folding, x*0, x*1, x+0), enough that generated code has
no dead arithmetic.  Expressions are typed `real` or `bool`; a bool multiplied by
a real is a select (ad.t:715-733), bool*bool is `and` (ad.t:814).

Not a port: the reference canonicalises n-ary sum/prod/powc with polynomial
simplification and then schedules instructions for register pressure itself;
here nvcc/NVRTC does CSE and scheduling, so the DAG stays binary and the emitter
is a plain topological walk.

What the constructors do normalise (the role of ad.t:841-964 `polysimplify` for the
patterns the operators are made of -- guarded, weighted sums of partial * vector):
  * guards move outward:  select(c,x,0)*y -> select(c,x*y,0);  select(c,x,0)*select(d,y,0) ->
    select(c&&d, x*y, 0);  select(c,x,0) +- select(c,y,0) -> select(c, x+-y, 0);
    select(c, select(d,x,0), 0) -> select(c&&d, x, 0).  Every partial derivative of a bounds- or
    mask-guarded residual carries the guard; J p and partial * (J p) then cost one select per
    residual instead of one per partial.  Values are unchanged wherever the operands are finite.
  * uniform factors (constants, scalar Params) move to the outside-left of products and out of
    guards, and a common uniform factor is taken out of a sum:  w*x + w*y -> w*(x+y).
  * multiplication by -1 becomes subtraction where a sum absorbs it.
Set THALLO_B200_SIMPLIFY=0 to switch these rewrites off (tuning / bisecting).
"""
import math
import os

REAL, BOOL = "real", "bool"
SIMPLIFY = os.environ.get("THALLO_B200_SIMPLIFY", "1") != "0"

_UNARY = ("sqrt", "sin", "cos", "tan", "exp", "log", "abs", "asin", "acos", "atan", "sinh", "cosh", "tanh")
_CMP = ("eq", "neq", "less", "greater", "lesseq", "greatereq")


class Exp:
    __slots__ = ("kind", "op", "args", "value", "key", "type", "id", "const")
    _table = {}
    _next = 0

    @staticmethod
    def _make(kind, op=None, args=(), value=None, key=None, type_=REAL, const=None):
        h = (kind, op, tuple(a.id for a in args), value, key, type_, const)
        e = Exp._table.get(h)
        if e is None:
            e = object.__new__(Exp)
            e.kind, e.op, e.args, e.value, e.key, e.type, e.const = kind, op, tuple(args), value, key, type_, const
            e.id = Exp._next
            Exp._next += 1
            Exp._table[h] = e
        return e

    # ---- python operators (non-scalar operands, e.g. dsl.Vector, take over via NotImplemented)
    def __add__(self, o): return add(self, o) if _scalar(o) else NotImplemented
    def __radd__(self, o): return add(o, self) if _scalar(o) else NotImplemented
    def __sub__(self, o): return sub(self, o) if _scalar(o) else NotImplemented
    def __rsub__(self, o): return sub(o, self) if _scalar(o) else NotImplemented
    def __mul__(self, o): return mul(self, o) if _scalar(o) else NotImplemented
    def __rmul__(self, o): return mul(o, self) if _scalar(o) else NotImplemented
    def __truediv__(self, o): return div(self, o) if _scalar(o) else NotImplemented
    def __rtruediv__(self, o): return div(o, self) if _scalar(o) else NotImplemented
    def __neg__(self): return mul(-1.0, self)
    def __pow__(self, c): return powc(self, c)

    def __getitem__(self, i):
        assert i == 0, "scalar expression indexed as vector"
        return self

    def __call__(self, i):
        return self[i]

    def __len__(self):
        return 1

    def is_const(self, v=None):
        return self.kind == "const" and (v is None or self.value == v)

    def __repr__(self):
        if self.kind == "const":
            return repr(self.value)
        if self.kind == "var":
            return str(self.key)
        return "%s(%s)" % (self.op if self.const is None else "%s[%s]" % (self.op, self.const), ",".join(map(repr, self.args)))


def _scalar(o):
    return isinstance(o, (Exp, int, float, bool))


def const(v):
    if isinstance(v, bool):
        return Exp._make("const", value=bool(v), type_=BOOL)
    return Exp._make("const", value=float(v), type_=REAL)


def var(key, type_=REAL):
    return Exp._make("var", key=key, type_=type_)


def toexp(x):
    if isinstance(x, Exp):
        return x
    if isinstance(x, (int, float)):
        return const(float(x))
    if isinstance(x, bool):
        return const(x)
    raise TypeError("cannot convert %r to an expression" % (x,))


def _apply(op, args, type_=REAL, const_=None):
    return Exp._make("apply", op=op, args=args, type_=type_, const=const_)


def _uniform(e):
    """Same value for every element of a launch: a constant or a scalar Param."""
    return e.kind == "const" or (e.kind == "var" and type(e.key).__name__ == "Param")


def _urank(e):
    return -1 if e.kind == "const" else e.id


def _guard(e):
    """(c, x) if e is select(c, x, 0), else None."""
    if e.kind == "apply" and e.op == "select" and e.args[2].is_const(0.0):
        return e.args[0], e.args[1]
    return None


def _ufactor(e):
    """(w, x) if e is w*x with w uniform, else None."""
    if e.kind == "apply" and e.op == "mul" and _uniform(e.args[0]):
        return e.args
    return None


def _addsub(op, a, b):
    """Rewrites shared by add and sub; None when none applies."""
    if not SIMPLIFY:
        return None
    f = add if op == "add" else sub
    ga, gb = _guard(a), _guard(b)
    if ga and gb and ga[0] is gb[0]:
        return select(ga[0], f(ga[1], gb[1]), 0.0)
    fa, fb = _ufactor(a), _ufactor(b)
    if fa and fb and fa[0] is fb[0]:
        return mul(fa[0], f(fa[1], fb[1]))
    if fb and fb[0].is_const(-1.0):                     # x + (-1)*y -> x - y;  x - (-1)*y -> x + y
        return (sub if op == "add" else add)(a, fb[1])
    if op == "add" and fa and fa[0].is_const(-1.0):     # (-1)*x + y -> y - x
        return sub(b, fa[1])
    return None


def add(a, b):
    a, b = toexp(a), toexp(b)
    if a.type == BOOL: a = select(a, 1.0, 0.0)
    if b.type == BOOL: b = select(b, 1.0, 0.0)
    if a.is_const() and b.is_const():
        return const(a.value + b.value)
    if a.is_const(0.0): return b
    if b.is_const(0.0): return a
    if a.is_const():          # constants to the right, canonical order
        a, b = b, a
    r = _addsub("add", a, b)
    if r is not None:
        return r
    return _apply("add", (a, b))


def sub(a, b):
    a, b = toexp(a), toexp(b)
    if b.type == BOOL: b = select(b, 1.0, 0.0)
    if a.type == BOOL: a = select(a, 1.0, 0.0)
    if a.is_const() and b.is_const():
        return const(a.value - b.value)
    if b.is_const(0.0): return a
    if a.is_const(0.0): return mul(-1.0, b)
    if a is b: return const(0.0)
    r = _addsub("sub", a, b)
    if r is not None:
        return r
    return _apply("sub", (a, b))


def mul(a, b):
    a, b = toexp(a), toexp(b)
    if a.type == BOOL and b.type == BOOL:
        return and_(a, b)
    if a.type == BOOL:
        return select(a, b, 0.0)
    if b.type == BOOL:
        return select(b, a, 0.0)
    if a.is_const() and b.is_const():
        return const(a.value * b.value)
    if a.is_const(0.0) or b.is_const(0.0): return const(0.0)
    if a.is_const(1.0): return b
    if b.is_const(1.0): return a
    if b.is_const():
        a, b = b, a                                   # constant first
    if a.is_const() and b.kind == "apply" and b.op == "mul" and b.args[0].is_const():
        return mul(const(a.value * b.args[0].value), b.args[1])
    if SIMPLIFY:
        if _uniform(b) and not _uniform(a):
            a, b = b, a                               # uniform factor first
        if _uniform(a):
            fb = _ufactor(b)
            if fb and _urank(fb[0]) < _urank(a):      # uniform factors ordered: constants, then Params by id
                return mul(fb[0], mul(a, fb[1]))
        else:
            fa, fb = _ufactor(a), _ufactor(b)
            if fa:                                    # (w*x)*y -> w*(x*y)
                return mul(fa[0], mul(fa[1], b))
            if fb:                                    # x*(w*y) -> w*(x*y)
                return mul(fb[0], mul(a, fb[1]))
            ga, gb = _guard(a), _guard(b)
            if ga and gb:
                return select(and_(ga[0], gb[0]), mul(ga[1], gb[1]), 0.0)
            if ga:
                return select(ga[0], mul(ga[1], b), 0.0)
            if gb:
                return select(gb[0], mul(a, gb[1]), 0.0)
    return _apply("mul", (a, b))


def powc(a, c):
    a = toexp(a)
    c = int(c) if float(c).is_integer() else float(c)
    if a.is_const():
        return const(a.value ** c)
    if c == 0: return const(1.0)
    if c == 1: return a
    if isinstance(c, int):
        if a.kind == "apply" and a.op == "powc" and isinstance(a.const, int):
            return powc(a.args[0], a.const * c)
        return _apply("powc", (a,), const_=c)
    return _apply("pow", (a, const(c)))


def div(a, b):
    """x/y is x * y^-1 (ad.t:714); a constant denominator folds into a constant factor."""
    a, b = toexp(a), toexp(b)
    if b.is_const():
        return mul(a, const(1.0 / b.value))
    return mul(a, powc(b, -1))


def unary(op, a):
    a = toexp(a)
    if a.is_const():
        f = dict(sqrt=math.sqrt, sin=math.sin, cos=math.cos, tan=math.tan, exp=math.exp, log=math.log,
                 abs=abs, asin=math.asin, acos=math.acos, atan=math.atan, sinh=math.sinh,
                 cosh=math.cosh, tanh=math.tanh)[op]
        return const(f(a.value))
    return _apply(op, (a,))


def select(c, a, b):
    c, a, b = toexp(c), toexp(a), toexp(b)
    assert c.type == BOOL, "select condition must be boolean"
    if a.type == BOOL and b.type == BOOL:
        return or_(and_(c, a), and_(not_(c), b))
    if c.is_const():
        return a if c.value else b
    if a is b:
        return a
    # select(c, select(c, x, y), b) -> select(c, x, b)
    if a.kind == "apply" and a.op == "select" and a.args[0] is c:
        a = a.args[1]
    if b.kind == "apply" and b.op == "select" and b.args[0] is c:
        b = b.args[2]
    if SIMPLIFY and b.is_const(0.0):
        if a.is_const(0.0):
            return a
        ga = _guard(a)
        if ga:                                        # select(c, select(d, x, 0), 0) -> select(c && d, x, 0)
            return select(and_(c, ga[0]), ga[1], 0.0)
        fa = _ufactor(a)
        if fa:                                        # select(c, w*x, 0) -> w*select(c, x, 0)
            return mul(fa[0], select(c, fa[1], 0.0))
    return _apply("select", (c, a, b))


def cmp(op, a, b):
    a, b = toexp(a), toexp(b)
    if a.is_const() and b.is_const():
        f = dict(eq=lambda x, y: x == y, neq=lambda x, y: x != y, less=lambda x, y: x < y,
                 greater=lambda x, y: x > y, lesseq=lambda x, y: x <= y, greatereq=lambda x, y: x >= y)[op]
        return const(bool(f(a.value, b.value)))
    return _apply(op, (a, b), BOOL)


def and_(a, b):
    a, b = toexp(a), toexp(b)
    assert a.type == BOOL and b.type == BOOL
    if a.is_const(): return b if a.value else const(False)
    if b.is_const(): return a if b.value else const(False)
    if a is b: return a
    if a.id > b.id: a, b = b, a
    return _apply("and", (a, b), BOOL)


def or_(a, b):
    a, b = toexp(a), toexp(b)
    assert a.type == BOOL and b.type == BOOL
    if a.is_const(): return const(True) if a.value else b
    if b.is_const(): return const(True) if b.value else a
    if a is b: return a
    if a.id > b.id: a, b = b, a
    return _apply("or", (a, b), BOOL)


def not_(a):
    a = toexp(a)
    assert a.type == BOOL
    if a.is_const(): return const(not a.value)
    if a.kind == "apply" and a.op == "not": return a.args[0]
    return _apply("not", (a,), BOOL)


def sample(image_key, dx_key, dy_key, x, y):
    """Bilinear sample of `image_key` at real coordinates (x, y); partials are samples of
    the derivative images (thallo.t:5803-5817)."""
    return _apply("sample", (toexp(x), toexp(y)), const_=(image_key, dx_key, dy_key))


# ---------------------------------------------------------------- differentiation
def derivative(e, x, memo=None):
    """d e / d x for a Var node x."""
    if memo is None:
        memo = {}
    return _d(e, x, memo)


def _d(e, x, memo):
    k = e.id
    if k in memo:
        return memo[k]
    if e is x:
        r = const(1.0)
    elif e.kind in ("const", "var") or e.type == BOOL:
        r = const(0.0)
    else:
        op, a = e.op, e.args
        if op == "add":
            r = add(_d(a[0], x, memo), _d(a[1], x, memo))
        elif op == "sub":
            r = sub(_d(a[0], x, memo), _d(a[1], x, memo))
        elif op == "mul":
            r = add(mul(_d(a[0], x, memo), a[1]), mul(a[0], _d(a[1], x, memo)))
        elif op == "powc":
            c = e.const
            r = mul(mul(float(c), powc(a[0], c - 1)), _d(a[0], x, memo))
        elif op == "pow":
            c = a[1].value
            r = mul(mul(c, _apply("pow", (a[0], const(c - 1.0)))), _d(a[0], x, memo))
        elif op == "select":
            r = select(a[0], _d(a[1], x, memo), _d(a[2], x, memo))
        elif op == "sample":
            ik, dxk, dyk = e.const
            assert dxk is not None and dyk is not None, "image derivatives are not defined for this sampled image"
            gx = _apply("sample", a, const_=(dxk, None, None))
            gy = _apply("sample", a, const_=(dyk, None, None))
            r = add(mul(gx, _d(a[0], x, memo)), mul(gy, _d(a[1], x, memo)))
        else:
            da = _d(a[0], x, memo)
            if da.is_const(0.0):
                r = da
            else:
                u = a[0]
                if op == "sqrt": g = div(1.0, mul(2.0, e))
                elif op == "sin": g = unary("cos", u)
                elif op == "cos": g = mul(-1.0, unary("sin", u))
                elif op == "tan": g = add(1.0, mul(e, e))
                elif op == "exp": g = e
                elif op == "log": g = div(1.0, u)
                elif op == "abs": g = select(cmp("greatereq", u, 0.0), 1.0, -1.0)
                elif op == "asin": g = div(1.0, unary("sqrt", sub(1.0, mul(u, u))))
                elif op == "acos": g = div(-1.0, unary("sqrt", sub(1.0, mul(u, u))))
                elif op == "atan": g = div(1.0, add(mul(u, u), 1.0))
                elif op == "sinh": g = unary("cosh", u)
                elif op == "cosh": g = unary("sinh", u)
                elif op == "tanh": g = div(1.0, mul(unary("cosh", u), unary("cosh", u)))
                else:
                    raise NotImplementedError(op)
                r = mul(g, da)
    memo[k] = r
    return r


def variables(e, pred=None, out=None, seen=None):
    """All Var nodes reachable from e (optionally filtered), in first-visit order."""
    if out is None:
        out, seen = [], set()
    stack = [e]
    while stack:
        n = stack.pop()
        if n.id in seen:
            continue
        seen.add(n.id)
        if n.kind == "var":
            if pred is None or pred(n):
                out.append(n)
        elif n.kind == "apply":
            stack.extend(reversed(n.args))
    return out


def substitute(e, fn, memo=None):
    """Rebuild e with every Var v replaced by fn(v) (fn returns an Exp)."""
    if memo is None:
        memo = {}
    if e.id in memo:
        return memo[e.id]
    if e.kind == "const":
        r = e
    elif e.kind == "var":
        r = fn(e)
    else:
        args = [substitute(a, fn, memo) for a in e.args]
        r = rebuild(e, args)
    memo[e.id] = r
    return r


def rebuild(e, args):
    op = e.op
    if op == "add": return add(*args)
    if op == "sub": return sub(*args)
    if op == "mul": return mul(*args)
    if op == "powc": return powc(args[0], e.const)
    if op == "pow": return _apply("pow", tuple(args))
    if op == "select": return select(*args)
    if op == "and": return and_(*args)
    if op == "or": return or_(*args)
    if op == "not": return not_(*args)
    if op in _CMP: return cmp(op, *args)
    if op == "sample": return _apply("sample", tuple(args), const_=e.const)
    return unary(op, args[0])


def toposort(roots):
    order, seen = [], set()
    for r in roots:
        stack = [(r, False)]
        while stack:
            n, done = stack.pop()
            if done:
                order.append(n)
                continue
            if n.id in seen:
                continue
            seen.add(n.id)
            stack.append((n, True))
            if n.kind == "apply":
                for a in reversed(n.args):
                    if a.id not in seen:
                        stack.append((a, False))
    return order
