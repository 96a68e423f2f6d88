"""A small interpreter for the Fortran subset the reference's hot path is written in.

TEST INFRASTRUCTURE ONLY (like everything under oracle/).  Purpose: the reference cannot be compiled in this image (no
Fortran front-end), and both oracles (oracle/tamc_oracle.c, oracle/pyref.py) are restatements written by the builder.  This
module instead EXECUTES THE REFERENCE'S OWN SOURCE TEXT -- ran2.f, sourceph.f90, inttau2.f90, stokes.f90, gridset.f90,
ch_opt.f90 (init_opt1), the module variable files, and statement ranges of mcpolar.f90 -- read from /root/reference at
fixture-generation time (tests/golden/make_reference_vectors.py; the vectors are committed, the sources are not copied).
Nothing of the reference's algorithm is restated here: this file knows Fortran, not photon transport.

Semantics implemented (what those files use):
  * free form and fixed form (ran2.f: column-6 continuation, labels in columns 1-5), case-insensitive;
  * modules with variables / parameters / contained procedures, `use m` and `use m, only: ...`, host association;
  * INTEGER / REAL / DOUBLE PRECISION / LOGICAL scalars and explicit-shape, assumed-shape and allocatable arrays with
    lower bounds; REAL is 8 bytes (src/Makefile:3 compiles with -freal-4-real-8: every REAL entity and real literal is
    promoted, so `1.e-8`, `6.283185` and `250d-4` are all converted from their decimal text straight to binary64);
  * PARAMETER (both syntaxes), SAVE, DATA with repeat counts; arguments passed by reference (scalars, array elements, whole
    arrays); function result variables; external functions declared by type (`real :: ran2`);
  * IF / ELSE IF / ELSE, one-line IF, DO (counted, labelled, WHILE, endless), EXIT, CYCLE, RETURN, GOTO to a label of the
    procedure's top level, CONTINUE, CALL, assignments to scalars / elements / whole arrays, array constructors;
  * expressions with Fortran's precedence and left-to-right evaluation of equal precedence, integer division truncating
    toward zero, mixed-mode arithmetic converting the integer operand, `**` with integer and real exponents;
  * intrinsics log, exp, sqrt, sin, cos, tan, acos, asin, atan, abs, min, max, int, nint, floor, real, dble, mod, sign, size.
    Transcendentals are the platform libm's (what gfortran's run time calls for binary64).
  * PRINT / WRITE statements are skipped, ERROR STOP raises.
Not implemented (not needed): 32-bit integer overflow (ran2's Schrage arithmetic never overflows by construction), derived
types, pointers, formatted I/O, array sections, WHERE / FORALL, implicit typing.
"""
from __future__ import annotations

import math
import re

import numpy as np


class FortranError(Exception):
    pass


class _Exit(Exception):
    pass


class _Cycle(Exception):
    pass


class _Return(Exception):
    pass


class _Goto(Exception):
    def __init__(self, label):
        self.label = label


# ----------------------------------------------------------------------------------------------------------------------
# storage
# ----------------------------------------------------------------------------------------------------------------------
class Cell:
    """A scalar variable.  t: 'i' integer, 'r' real (binary64), 'l' logical."""
    __slots__ = ("t", "v")

    def __init__(self, t, v=None):
        self.t = t
        self.v = v

    def set(self, x):
        t = self.t
        if t == "r":
            self.v = float(x)
        elif t == "i":
            self.v = int(x) if not isinstance(x, float) else int(math.trunc(x))
        elif t == "c":
            self.v = str(x)
        else:
            self.v = bool(x)


class FArray:
    """An array with per-dimension lower bounds, column-major like Fortran (irrelevant here: indexed access only)."""
    __slots__ = ("t", "a", "lb")

    def __init__(self, t, shape, lb=None):
        self.t = t
        dt = {"r": np.float64, "i": np.int64, "l": np.bool_}[t]
        self.a = np.zeros(shape, dtype=dt)
        self.lb = tuple(lb) if lb is not None else (1,) * len(shape)

    def _ix(self, idx):
        if len(idx) != self.a.ndim:
            raise FortranError(f"rank mismatch: {len(idx)} subscripts for rank {self.a.ndim}")
        out = []
        for k, (i, lo, n) in enumerate(zip(idx, self.lb, self.a.shape)):
            j = i - lo
            if j < 0 or j >= n:
                raise FortranError(f"subscript {i} of dimension {k + 1} outside [{lo}, {lo + n - 1}]")
            out.append(j)
        return tuple(out)

    def get(self, idx):
        v = self.a[self._ix(idx)]
        return float(v) if self.t == "r" else (int(v) if self.t == "i" else bool(v))

    def put(self, idx, x):
        if self.t == "i" and isinstance(x, float):
            x = math.trunc(x)
        self.a[self._ix(idx)] = x

    def view(self, lb):
        """The same storage under other lower bounds (an explicit-shape dummy associated with this array)."""
        v = FArray.__new__(FArray)
        v.t, v.a, v.lb = self.t, self.a, tuple(lb)
        return v

    def _np_index(self, idx):
        """Subscripts with sections -> a numpy index (scalars drop their dimension, as in Fortran)."""
        out = []
        for k, (i, lo, n) in enumerate(zip(idx, self.lb, self.a.shape)):
            if isinstance(i, Sec):
                a = 0 if i.lo is None else i.lo - lo
                b = n if i.hi is None else i.hi - lo + 1
                if a < 0 or b > n:
                    raise FortranError(f"section {i.lo}:{i.hi} of dimension {k + 1} outside [{lo}, {lo + n - 1}]")
                out.append(slice(a, b))
            else:
                j = i - lo
                if j < 0 or j >= n:
                    raise FortranError(f"subscript {i} of dimension {k + 1} outside [{lo}, {lo + n - 1}]")
                out.append(j)
        return tuple(out)

    def section(self, idx):
        if len(idx) != self.a.ndim:
            raise FortranError("rank mismatch in an array section")
        part = self.a[self._np_index(idx)]
        r = FArray.__new__(FArray)
        r.t, r.a, r.lb = self.t, np.array(part, copy=True), (1,) * part.ndim
        return r

    def put_section(self, idx, x):
        if len(idx) != self.a.ndim:
            raise FortranError("rank mismatch in an array section")
        self.a[self._np_index(idx)] = x.a if isinstance(x, FArray) else x

    @staticmethod
    def _wrap(a, like):
        r = FArray.__new__(FArray)
        r.t, r.a, r.lb = "r", np.asarray(a, dtype=np.float64), (1,) * np.ndim(a)
        return r

    # whole-array arithmetic (elementwise, binary64), e.g. mcpolar.f90:174  jmeanGLOBAL = jmeanGLOBAL * (scalar)
    def __mul__(self, o): return FArray._wrap(self.a * (o.a if isinstance(o, FArray) else o), self)
    def __rmul__(self, o): return FArray._wrap(o * self.a, self)
    def __add__(self, o): return FArray._wrap(self.a + (o.a if isinstance(o, FArray) else o), self)
    def __radd__(self, o): return FArray._wrap(o + self.a, self)
    def __sub__(self, o): return FArray._wrap(self.a - (o.a if isinstance(o, FArray) else o), self)
    def __rsub__(self, o): return FArray._wrap(o - self.a, self)
    def __truediv__(self, o): return FArray._wrap(self.a / (o.a if isinstance(o, FArray) else o), self)
    def __rtruediv__(self, o): return FArray._wrap(o / self.a, self)
    def __neg__(self): return FArray._wrap(-self.a, self)

    def fill(self, x):
        if isinstance(x, FArray):
            if x.a.shape != self.a.shape:
                raise FortranError(f"array assignment of shape {x.a.shape} to shape {self.a.shape}")
            self.a[...] = x.a
        elif isinstance(x, list):
            if len(x) != self.a.size:
                raise FortranError("array constructor of the wrong size")
            self.a.reshape(-1, order="F")[...] = x
        else:
            self.a[...] = x


class Sec:
    """A subscript triplet lo:hi (stride 1); None = the bound of the dimension."""
    __slots__ = ("lo", "hi")

    def __init__(self, lo, hi):
        self.lo, self.hi = lo, hi


class ElemRef:
    """An array element passed as an actual argument: behaves like a Cell."""
    __slots__ = ("arr", "idx", "t")

    def __init__(self, arr, idx):
        self.arr, self.idx, self.t = arr, idx, arr.t

    @property
    def v(self):
        return self.arr.get(self.idx)

    def set(self, x):
        self.arr.put(self.idx, x)


# ----------------------------------------------------------------------------------------------------------------------
# source normalisation
# ----------------------------------------------------------------------------------------------------------------------
def _strip_comment(line):
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out)


def logical_lines(text, fixed=False):
    """-> [(first source line number, label or None, statement text lowercased)]"""
    raw = text.splitlines()
    stmts = []
    if fixed:
        cur = None
        for no, line in enumerate(raw, 1):
            if not line.strip() or line[0] in "cC*!":
                continue
            line = _strip_comment(line.rstrip())
            if len(line) > 5 and line[5] not in " 0" and not line[:5].strip():
                cur[2] += " " + line[6:].strip()
                continue
            if cur:
                stmts.append(tuple(cur))
            label = line[:5].strip() or None
            cur = [no, label, line[6:].strip()]
        if cur:
            stmts.append(tuple(cur))
    else:
        cur, start = "", None
        for no, line in enumerate(raw, 1):
            line = _strip_comment(line).strip()
            if not line:
                continue
            if cur and line.startswith("&"):
                line = line[1:].lstrip()
            if start is None:
                start = no
            if line.endswith("&"):
                cur += line[:-1]
                continue
            cur += line
            m = re.match(r"^(\d+)\s+(.*)$", cur)
            label, body = (m.group(1), m.group(2)) if m else (None, cur)
            for part in _split_top(body, ";"):
                if part.strip():
                    stmts.append((start, label, part.strip()))
                    label = None
            cur, start = "", None
    return [(no, lab, _lower_outside_strings(s)) for no, lab, s in stmts]


def _lower_outside_strings(s):
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        else:
            if ch in "'\"":
                q = ch
            out.append(ch.lower())
    return "".join(out)


def _split_top(s, sep=","):
    """Split at `sep` outside parentheses, brackets and strings."""
    parts, depth, q, cur = [], 0, None, []
    i = 0
    while i < len(s):
        ch = s[i]
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch in "([":
            depth += 1
            cur.append(ch)
        elif ch in ")]":
            depth -= 1
            cur.append(ch)
        elif ch == sep and depth == 0:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
        i += 1
    parts.append("".join(cur))
    return parts


# ----------------------------------------------------------------------------------------------------------------------
# expressions
# ----------------------------------------------------------------------------------------------------------------------
_DOTTED = ("eqv", "neqv", "or", "and", "not", "eq", "ne", "lt", "le", "gt", "ge", "true", "false")
_TOKEN = re.compile(
    r"\s*(?:"
    r"(?P<num>(?:\d+\.(?!(?:" + "|".join(_DOTTED) + r")\.)\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?(?:_\w+)?)"
    r"|(?P<dot>\.(?:" + "|".join(_DOTTED) + r")\.)"
    r"|(?P<name>[a-z_]\w*)"
    r"|(?P<str>'[^']*'|\"[^\"]*\")"
    r"|(?P<op>\(/|/\)|\*\*|==|/=|<=|>=|[-+*/<>(),:=\[\]%])"
    r")")


def tokenize(s):
    toks, pos = [], 0
    s = s.rstrip()
    while pos < len(s):
        m = _TOKEN.match(s, pos)
        if not m or m.end() == pos:
            raise FortranError(f"cannot tokenise {s[pos:]!r} in {s!r}")
        pos = m.end()
        kind = m.lastgroup
        toks.append((kind, m.group(kind)))
    return toks


def _number(text):
    text = re.sub(r"_\w+$", "", text)
    if re.fullmatch(r"\d+", text):
        return int(text)
    return float(text.replace("d", "e"))          # every real literal is binary64 (-freal-4-real-8; `d` exponents anyway)


def _idiv(a, b):
    if b == 0:
        raise FortranError("integer division by zero")
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _pow(a, b):
    if isinstance(a, int) and isinstance(b, int):
        if b >= 0:
            return a ** b
        return _idiv(1, a ** (-b))
    if isinstance(b, int):
        # real ** integer: gfortran expands it into multiplications (__builtin_powi: x**2 -> x*x), which is NOT always
        # what libm's pow(x, 2.0) returns (pow is accurate to < 1 ulp, x*x is correctly rounded; they differed in step 22
        # of a stokes chain).  Square-and-multiply as powi does; only **2 occurs in the interpreted files.
        x, n, r = float(a), abs(b), 1.0
        while n:
            if n & 1:
                r = r * x
            n >>= 1
            if n:
                x = x * x
        return r if b >= 0 else 1.0 / r
    # real ** real: GCC folds pow(x, 2.0) into x*x at every optimisation level; other exponents go to libm
    if float(b) == 2.0:
        return float(a) * float(a)
    return math.pow(float(a), float(b))


def _fmod(a, b):
    if isinstance(a, int) and isinstance(b, int):
        return a - _idiv(a, b) * b
    return math.fmod(a, b)


def _sign(a, b):
    m = abs(a)
    return m if b >= 0 else -m


def _maxmin(fn):
    def f(*xs):
        if any(isinstance(x, float) for x in xs):
            xs = [float(x) for x in xs]
        return fn(xs)
    return f


INTRINSICS = {
    "log": lambda x: math.log(x), "exp": math.exp, "sqrt": lambda x: math.sqrt(x), "sin": math.sin, "cos": math.cos,
    "tan": math.tan, "acos": math.acos, "asin": math.asin, "atan": math.atan, "abs": abs,
    "min": _maxmin(min), "max": _maxmin(max), "int": lambda x: int(math.trunc(x)), "nint": lambda x: int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5)),
    "floor": lambda x: int(math.floor(x)), "real": float, "dble": float, "mod": _fmod, "sign": _sign,
    "trim": lambda c: c.rstrip(),
}


class _Parser:
    """Recursive descent; every parse_* returns a closure  f(frame) -> value."""

    def __init__(self, toks, interp, text):
        self.t, self.i, self.interp, self.text = toks, 0, interp, text

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def take(self, value=None):
        k, v = self.peek()
        if value is not None and v != value:
            raise FortranError(f"expected {value!r}, found {v!r} in {self.text!r}")
        self.i += 1
        return k, v

    def at_end(self):
        return self.i >= len(self.t)

    # precedence, lowest first
    def parse_expr(self):
        left = self.parse_or()
        while self.peek()[1] in (".eqv.", ".neqv."):
            op = self.take()[1]
            right = self.parse_or()
            left = (lambda a, b: lambda f: bool(a(f)) == bool(b(f)))(left, right) if op == ".eqv." else \
                   (lambda a, b: lambda f: bool(a(f)) != bool(b(f)))(left, right)
        return left

    def parse_or(self):
        left = self.parse_and()
        while self.peek()[1] == ".or.":
            self.take()
            right = self.parse_and()
            left = (lambda a, b: lambda f: bool(a(f)) | bool(b(f)))(left, right)
        return left

    def parse_and(self):
        left = self.parse_not()
        while self.peek()[1] == ".and.":
            self.take()
            right = self.parse_not()
            left = (lambda a, b: lambda f: bool(a(f)) & bool(b(f)))(left, right)
        return left

    def parse_not(self):
        if self.peek()[1] == ".not.":
            self.take()
            x = self.parse_not()
            return lambda f: not x(f)
        return self.parse_rel()

    _REL = {"==": "eq", ".eq.": "eq", "/=": "ne", ".ne.": "ne", "<": "lt", ".lt.": "lt", "<=": "le", ".le.": "le",
            ">": "gt", ".gt.": "gt", ">=": "ge", ".ge.": "ge"}

    def parse_rel(self):
        left = self.parse_add()
        op = self._REL.get(self.peek()[1])
        if op:
            self.take()
            right = self.parse_add()
            return {"eq": lambda a, b: lambda f: a(f) == b(f), "ne": lambda a, b: lambda f: a(f) != b(f),
                    "lt": lambda a, b: lambda f: a(f) < b(f), "le": lambda a, b: lambda f: a(f) <= b(f),
                    "gt": lambda a, b: lambda f: a(f) > b(f), "ge": lambda a, b: lambda f: a(f) >= b(f)}[op](left, right)
        return left

    def parse_add(self):
        k, v = self.peek()
        if v == "-":
            self.take()
            x = self.parse_mul()
            left = lambda f, x=x: -x(f)
        elif v == "+":
            self.take()
            left = self.parse_mul()
        else:
            left = self.parse_mul()
        while self.peek()[1] in ("+", "-") and self.peek()[0] == "op":
            op = self.take()[1]
            right = self.parse_mul()
            left = (lambda a, b: lambda f: a(f) + b(f))(left, right) if op == "+" else (lambda a, b: lambda f: a(f) - b(f))(left, right)
        return left

    def parse_mul(self):
        left = self.parse_pow()
        while self.peek()[1] in ("*", "/") and self.peek()[0] == "op":
            op = self.take()[1]
            right = self.parse_pow()
            if op == "*":
                left = (lambda a, b: lambda f: a(f) * b(f))(left, right)
            else:
                def div(f, a=left, b=right):
                    x, y = a(f), b(f)
                    if isinstance(x, FArray) or isinstance(y, FArray):
                        return x / y
                    if isinstance(x, int) and isinstance(y, int) and not isinstance(x, bool):
                        return _idiv(x, y)
                    return x / y if y != 0 else (math.copysign(math.inf, x) * math.copysign(1.0, y) if x != 0 else math.nan)
                left = div
        return left

    def parse_pow(self):
        base = self.parse_primary()
        if self.peek()[1] == "**":
            self.take()
            # right associative; a unary minus may follow ** (x**-2) -- not used, but harmless
            if self.peek()[1] == "-":
                self.take()
                e = self.parse_pow()
                return lambda f: _pow(base(f), -e(f))
            e = self.parse_pow()
            return lambda f: _pow(base(f), e(f))
        return base

    def parse_args(self, close=")"):
        args = []
        if self.peek()[1] == close:
            self.take()
            return args
        while True:
            args.append(self.parse_arg())
            k, v = self.take()
            if v == close:
                return args
            if v != ",":
                raise FortranError(f"expected , or {close} in {self.text!r}")

    def parse_arg(self):
        """An actual argument / subscript: returns (kind, payload, value closure); kind var | elem | expr."""
        start = self.i
        k, v = self.peek()
        if v == ":":                                          # a section  :hi  or  :
            self.take()
            hi = None if self.peek()[1] in (",", ")") else self.parse_expr()
            return ("sec", None, (lambda f, hi=hi: Sec(None, None if hi is None else int(hi(f)))))
        if k == "name":
            # a bare name, or name(subscripts), followed directly by , or ) : may be passed by reference
            j = self.i + 1
            if j >= len(self.t) or self.t[j][1] in (",", ")", "/)", "]"):
                self.take()
                return ("var", v, self._name_value(v))
            if self.t[j][1] == "(":
                depth, m = 0, j
                while m < len(self.t):
                    if self.t[m][1] in ("(", "(/"):
                        depth += 1
                    elif self.t[m][1] in (")", "/)"):
                        depth -= 1
                        if depth == 0:
                            break
                    m += 1
                if m + 1 >= len(self.t) or self.t[m + 1][1] in (",", ")", "/)", "]"):
                    self.take()
                    self.take("(")
                    subs = self.parse_args()
                    return ("elem", (v, subs), self._call_or_index(v, subs))
        self.i = start
        e = self.parse_expr()
        if self.peek()[1] == ":":                             # a section  lo:hi  or  lo:
            self.take()
            hi = None if self.peek()[1] in (",", ")") else self.parse_expr()
            return ("sec", None, (lambda f, lo=e, hi=hi: Sec(int(lo(f)), None if hi is None else int(hi(f)))))
        return ("expr", None, e)

    def _name_value(self, name):
        def get(f):
            try:
                x = f[name]
            except KeyError:
                raise FortranError(f"undefined name {name!r} in {self.text!r}") from None
            if isinstance(x, FArray):
                return x
            if x.v is None:
                raise FortranError(f"variable {name!r} used before it was given a value in {self.text!r}")
            return x.v
        return get

    def _call_or_index(self, name, args):
        interp = self.interp
        vals = [a[2] for a in args]
        has_sec = any(a[0] == "sec" for a in args)

        def run(f):
            x = f.get(name)
            if isinstance(x, FArray):
                idx = tuple(v(f) for v in vals)
                if has_sec:
                    return x.section(idx)
                return x.get(idx)
            if name == "size":
                arr = vals[0](f)
                return int(arr.a.size) if len(vals) == 1 else int(arr.a.shape[vals[1](f) - 1])
            if name in INTRINSICS and not interp.has_procedure(name):
                return INTRINSICS[name](*[v(f) for v in vals])
            return interp.call(name, [interp.actual(a, f) for a in args], want_result=True)
        return run

    def parse_primary(self):
        k, v = self.take()
        if k == "num":
            x = _number(v)
            return lambda f: x
        if k == "dot" and v in (".true.", ".false."):
            b = v == ".true."
            return lambda f: b
        if k == "str":
            s = v[1:-1]
            return lambda f: s
        if v == "(":
            e = self.parse_expr()
            self.take(")")
            return e
        if v in ("(/", "["):
            items = self.parse_args("/)" if v == "(/" else "]")
            vals = [a[2] for a in items]
            return lambda f: [x(f) for x in vals]
        if k == "name":
            if self.peek()[1] == "(":
                self.take()
                args = self.parse_args()
                return self._call_or_index(v, args)
            return self._name_value(v)
        raise FortranError(f"unexpected token {v!r} in {self.text!r}")


# ----------------------------------------------------------------------------------------------------------------------
# program units
# ----------------------------------------------------------------------------------------------------------------------
_TYPE = r"(integer|real|double\s+precision|logical|character)"
_DECL = re.compile(r"^" + _TYPE + r"\b\s*(\([^)]*\))?\s*((?:,\s*[a-z]+(?:\s*\([^)]*\))?\s*)*)(::)?\s*(.*)$")


class Procedure:
    def __init__(self, name, kind, args, stmts, file, host, result_type=None):
        self.name, self.kind, self.args, self.stmts, self.file, self.host = name, kind, args, stmts, file, host
        self.result_type = result_type
        self.compiled = None
        self.saved = {}          # SAVE'd locals (and DATA-initialised ones)
        self.calls = 0


class Module:
    def __init__(self, name):
        self.name = name
        self.vars = {}           # name -> Cell | FArray | None (allocatable, not yet allocated)
        self.alloc_types = {}    # allocatable arrays: name -> (type, rank)
        self.uses = []


class Interpreter:
    def __init__(self):
        self.modules = {}
        self.procs = {}
        self._blocks = {}
        self.alias = {}              # a procedure pointer bound by the harness: name -> procedure name
        self.skipped_calls = set()   # CALLs compiled to nothing besides mpi_* (e.g. checkallocate)

    # -- loading ---------------------------------------------------------------------------------------------------
    def load(self, path, fixed=None):
        text = open(path).read()
        if fixed is None:
            fixed = path.lower().endswith((".f", ".for"))
        self.load_text(text, fixed, path)

    def load_text(self, text, fixed=False, file="<text>"):
        lines = logical_lines(text, fixed)
        i = 0
        while i < len(lines):
            no, lab, s = lines[i]
            m = re.match(r"^module\s+(\w+)$", s)
            if m and not s.startswith("module procedure"):
                i = self._load_module(lines, i, m.group(1), file)
                continue
            m = self._proc_header(s)
            if m:
                i = self._load_proc(lines, i, m, file, None)
                continue
            if re.match(r"^program\b", s):
                # a main program: skipped (statement ranges of it are run with run_block)
                while not re.match(r"^end\s*program\b", lines[i][2]):
                    i += 1
            i += 1

    @staticmethod
    def _proc_header(s):
        m = re.match(r"^((?:(?:elemental|pure|recursive)\s+)*)(?:(integer|real|logical|double\s+precision)\s+)?(subroutine|function)\s+(\w+)"
                     r"\s*(?:\(([^)]*)\))?\s*(?:result\s*\(\s*(\w+)\s*\))?$", s)
        if not m:
            return None
        args = [a.strip() for a in (m.group(5) or "").split(",") if a.strip()]
        return m.group(4), m.group(3), args, m.group(2), m.group(6), "elemental" in m.group(1)

    def _load_module(self, lines, i, name, file):
        mod = Module(name)
        self.modules[name] = mod
        i += 1
        frame = {}
        while True:
            no, lab, s = lines[i]
            if re.match(r"^end\s*module\b", s):
                return i + 1
            if s == "contains":
                i += 1
                while not re.match(r"^end\s*module\b", lines[i][2]):
                    hdr = self._proc_header(lines[i][2])
                    if hdr:
                        i = self._load_proc(lines, i, hdr, file, mod)
                    else:
                        i += 1
                return i + 1
            m = re.match(r"^use\s+(\w+)", s)
            if m:
                mod.uses.append(s)
            elif s in ("implicit none", "save") or re.match(r"^(private|public|procedure)\b", s):
                pass
            else:
                self._declare(s, mod.vars, frame_for_eval=mod.vars, module=mod, where=f"{file}:{no}")
            i += 1

    def _load_proc(self, lines, i, hdr, file, host):
        name, kind, args, rtype, rname, elemental = hdr
        body = []
        i += 1
        while True:
            no, lab, s = lines[i]
            if re.match(r"^end\s*(subroutine|function)?(\s+\w+)?$", s) and not re.match(r"^end\s*(if|do|module|program)", s):
                break
            body.append((no, lab, s))
            i += 1
        self.procs[name] = Procedure(name, kind, args, body, file, host, rtype)
        self.procs[name].result_name = rname or name
        self.procs[name].elemental = elemental
        return i + 1

    def has_procedure(self, name):
        return name in self.procs or name in self.alias

    # -- declarations ----------------------------------------------------------------------------------------------
    def _eval(self, text, frame):
        p = _Parser(tokenize(text), self, text)
        e = p.parse_expr()
        if not p.at_end():
            raise FortranError(f"trailing tokens in {text!r}")
        return e(frame)

    def _declare(self, s, scope, frame_for_eval, module=None, where="", dummies=(), proc=None):
        """One specification statement into `scope`.  Returns True when s was one."""
        m = re.match(r"^parameter\s*\((.*)\)$", s)
        if m:
            for item in _split_top(m.group(1)):
                k, v = item.split("=", 1)
                k = k.strip()
                cell = scope[k]
                cell.set(self._eval(v, frame_for_eval))
            return True
        if re.match(r"^save\b", s):
            if proc is not None:
                names = [x.strip() for x in s[4:].split(",") if x.strip()]
                proc_saved = proc.saved
                for nm in names or list(scope):
                    if nm in scope and nm not in proc.args:
                        proc_saved.setdefault(nm, scope[nm])
            return True
        m = re.match(r"^data\s+(.*)$", s)
        if m:
            for item in re.findall(r"(\w+)\s*/([^/]*)/", m.group(1)):
                nm, vals = item
                out = []
                for v in _split_top(vals):
                    if "*" in v:
                        rep, val = v.split("*", 1)
                        out += [self._eval(val, frame_for_eval)] * int(self._eval(rep, frame_for_eval))
                    else:
                        out.append(self._eval(v, frame_for_eval))
                tgt = scope[nm]
                if isinstance(tgt, FArray):
                    tgt.fill(out)
                else:
                    tgt.set(out[0])
                if proc is not None:
                    proc.saved.setdefault(nm, tgt)          # DATA implies SAVE
            return True
        m = _DECL.match(s)
        if not m:
            return False
        tname, _kind, attrs, dcolon, rest = m.groups()
        if not dcolon and not rest:
            return False
        t = {"integer": "i", "real": "r", "logical": "l", "character": "c"}.get(tname, "r")
        attrs = attrs.lower()
        is_param = "parameter" in attrs
        is_alloc = "allocatable" in attrs
        dim_attr = re.search(r"dimension\s*\(([^)]*)\)", attrs)
        for ent in _split_top(rest):
            ent = ent.strip()
            if not ent:
                continue
            init = None
            if "=" in ent and is_param or re.match(r"^\w+(\([^)]*\))?\s*=", ent):
                ent, init = ent.split("=", 1)
                ent = ent.strip()
            mm = re.match(r"^(\w+)\s*(?:\((.*)\))?$", ent)
            if not mm:
                raise FortranError(f"cannot parse entity {ent!r} at {where}")
            nm, dims = mm.group(1), mm.group(2) or (dim_attr.group(1) if dim_attr else None)
            if nm in self.procs and nm not in dummies and dims is None and (proc is None or nm != proc.name):
                continue                                     # `real :: ran2` -- the type of an external function
            if proc is not None and proc.kind == "function" and nm == proc.result_name:
                continue                                     # the function result
            if nm in dummies:
                if dims is not None and isinstance(scope.get(nm), FArray):
                    specs = [d.strip() for d in _split_top(dims)]
                    if all(d != ":" for d in specs):
                        lbs = [int(self._eval(d.split(":")[0], frame_for_eval)) if ":" in d else 1 for d in specs]
                        if len(lbs) == scope[nm].a.ndim and tuple(lbs) != scope[nm].lb:
                            scope[nm] = scope[nm].view(lbs)
                continue                                     # bound to the actual argument
            if dims is not None:
                specs = [d.strip() for d in _split_top(dims)]
                if is_alloc or any(d == ":" for d in specs):
                    if module is not None:
                        module.alloc_types[nm] = (t, len(specs))
                        scope[nm] = None
                    continue
                shape, lbs = [], []
                for d in specs:
                    if ":" in d:
                        lo, hi = d.split(":")
                        lo, hi = int(self._eval(lo, frame_for_eval)), int(self._eval(hi, frame_for_eval))
                    else:
                        lo, hi = 1, int(self._eval(d, frame_for_eval))
                    shape.append(hi - lo + 1)
                    lbs.append(lo)
                scope[nm] = FArray(t, tuple(shape), lbs)
            else:
                scope[nm] = Cell(t)
                if init is not None:
                    scope[nm].set(self._eval(init, frame_for_eval))
        return True

    def _dummy_types(self, proc):
        """name -> type letter for the dummies (and the function result), from the procedure's declarations."""
        types = {}
        for no, lab, s in proc.stmts:
            m = _DECL.match(s)
            if not m or not (m.group(4) or m.group(5)):
                continue
            t = {"integer": "i", "real": "r", "logical": "l"}.get(m.group(1), "r")
            for ent in _split_top(m.group(5)):
                nm = re.match(r"^\s*(\w+)", ent)
                if nm:
                    types[nm.group(1)] = t
        return types

    # -- module access for the harness --------------------------------------------------------------------------------
    def allocate(self, module, name, shape, lb=None):
        mod = self.modules[module]
        t, rank = mod.alloc_types[name]
        if rank != len(shape):
            raise FortranError(f"{name}: rank {rank} expected")
        mod.vars[name] = FArray(t, tuple(shape), lb)
        return mod.vars[name]

    def var(self, module, name):
        return self.modules[module].vars[name]

    def _import(self, use_stmt, frame):
        m = re.match(r"^use\s+(\w+)\s*(?:,\s*only\s*:\s*(.*))?$", use_stmt)
        if not m:
            raise FortranError(f"cannot parse {use_stmt!r}")
        mod = self.modules.get(m.group(1))
        if mod is None:
            raise FortranError(f"module {m.group(1)!r} is not loaded")
        # (no module of the interpreted set re-exports what it uses, so `use` does not chain)
        if m.group(2):
            for item in m.group(2).split(","):
                item = item.strip()
                if not item:
                    continue
                local, _, remote = item.partition("=>")
                local, remote = local.strip(), (remote.strip() or local.strip())
                if remote in mod.vars:
                    if mod.vars[remote] is None:
                        raise FortranError(f"{remote} of module {mod.name} is not allocated")
                    frame[local] = mod.vars[remote]
                elif remote not in self.procs:
                    raise FortranError(f"{remote!r} is not in module {mod.name}")
        else:
            for k, v in mod.vars.items():
                if v is not None:
                    frame[k] = v

    # -- statements ---------------------------------------------------------------------------------------------------
    def _compile_block(self, stmts, pos, enders, ctx):
        """Compile statements from pos until one matching `enders`; returns (list of (label, exec), index of the ender)."""
        out = []
        while pos < len(stmts):
            no, lab, s = stmts[pos]
            for e in enders:
                if re.match(e, s) or (lab is not None and e == "label:" + lab):
                    return out, pos
            fn, pos = self._compile_stmt(stmts, pos, ctx)
            if fn is not None:
                out.append((lab, fn, no))
        if enders:
            raise FortranError(f"block not closed (looking for {enders}) in {ctx['where']}")
        return out, pos

    @staticmethod
    def _run(block, f):
        for lab, fn, no in block:
            fn(f)

    def _compile_stmt(self, stmts, pos, ctx):
        no, lab, s = stmts[pos]
        where = f"{ctx['where']}:{no}"
        try:
            return self._compile_stmt_inner(stmts, pos, ctx, s, lab, where)
        except FortranError as e:
            if "@" not in str(e):
                raise FortranError(f"{e} @ {where}") from None
            raise

    def _compile_stmt_inner(self, stmts, pos, ctx, s, lab, where):
        interp = self
        if s in ("implicit none", "continue", "contains") or re.match(r"^(print|write|format|open|close|read)\b", s) or \
                re.match(r"^(intent|external|intrinsic)\b", s):
            return (lambda f: None), pos + 1
        if re.match(r"^use\s", s) or re.match(r"^(parameter\s*\(|save\b|data\s)", s) or (_DECL.match(s) and ("::" in s or not re.match(r"^\w+\s*(\(|=)", s))):
            return None, pos + 1                             # specification part: handled at call entry
        if re.match(r"^error\s+stop", s) or re.match(r"^stop\b", s):
            def stop(f):
                raise FortranError(f"ERROR STOP at {where}")
            return stop, pos + 1
        if s == "return":
            def ret(f):
                raise _Return()
            return ret, pos + 1
        if s == "exit":
            def ex(f):
                raise _Exit()
            return ex, pos + 1
        if s == "cycle":
            def cy(f):
                raise _Cycle()
            return cy, pos + 1
        m = re.match(r"^go\s*to\s+(\d+)$", s)
        if m:
            label = m.group(1)

            def go(f):
                raise _Goto(label)
            return go, pos + 1
        m = re.match(r"^call\s+(\w+)\s*(?:\((.*)\))?$", s)
        if m and (m.group(1).startswith("mpi_") or m.group(1) in self.skipped_calls):
            return (lambda f: None), pos + 1
        if m:
            name = m.group(1)
            args = []
            if m.group(2) is not None and m.group(2).strip():
                p = _Parser(tokenize(m.group(2) + ")"), self, s)
                args = p.parse_args()

            def call(f):
                interp.call(name, [interp.actual(a, f) for a in args])
            return call, pos + 1
        # IF constructs
        m = re.match(r"^if\s*\(", s)
        if m:
            cond_text, rest = self._paren_split(s[s.index("("):])
            cond = self._expr(cond_text, s)
            if rest.strip() == "then":
                branches, other = [], None
                body, pos = self._compile_block(stmts, pos + 1, (r"^else\s*if\s*\(", r"^else$", r"^end\s*if$"), ctx)
                branches.append((cond, body))
                while True:
                    s2 = stmts[pos][2]
                    if re.match(r"^end\s*if$", s2):
                        break
                    if s2 == "else":
                        other, pos = self._compile_block(stmts, pos + 1, (r"^end\s*if$",), ctx)
                        break
                    c_text, r2 = self._paren_split(s2[s2.index("("):])
                    if r2.strip() != "then":
                        raise FortranError(f"else if without then at {where}")
                    c2 = self._expr(c_text, s2)
                    body, pos = self._compile_block(stmts, pos + 1, (r"^else\s*if\s*\(", r"^else$", r"^end\s*if$"), ctx)
                    branches.append((c2, body))

                def ifc(f):
                    for c, b in branches:
                        if c(f):
                            interp._run(b, f)
                            return
                    if other is not None:
                        interp._run(other, f)
                return ifc, pos + 1
            # one-line IF
            inner, _ = self._compile_stmt([(stmts[pos][0], None, rest.strip())], 0, ctx)

            def if1(f):
                if cond(f):
                    inner(f)
            return if1, pos + 1
        # DO constructs
        m = re.match(r"^do\s+while\s*\(", s)
        if m:
            cond_text, _ = self._paren_split(s[s.index("("):])
            cond = self._expr(cond_text, s)
            body, pos = self._compile_block(stmts, pos + 1, (r"^end\s*do$",), ctx)

            def dow(f):
                while cond(f):
                    try:
                        interp._run(body, f)
                    except _Cycle:
                        continue
                    except _Exit:
                        break
            return dow, pos + 1
        if s == "do":
            body, pos = self._compile_block(stmts, pos + 1, (r"^end\s*do$",), ctx)

            def doe(f):
                while True:
                    try:
                        interp._run(body, f)
                    except _Cycle:
                        continue
                    except _Exit:
                        break
            return doe, pos + 1
        m = re.match(r"^do\s+(?:(\d+)\s+)?(\w+)\s*=\s*(.*)$", s)
        if m:
            label, var, rng = m.groups()
            parts = _split_top(rng)
            lo, hi = self._expr(parts[0], s), self._expr(parts[1], s)
            st = self._expr(parts[2], s) if len(parts) > 2 else (lambda f: 1)
            if label:
                body, pos = self._compile_block(stmts, pos + 1, ("label:" + label,), ctx)
                if stmts[pos][2] != "continue":
                    raise FortranError(f"labelled DO must end in CONTINUE at {where}")
            else:
                body, pos = self._compile_block(stmts, pos + 1, (r"^end\s*do$",), ctx)

            def doc(f):
                a, b, c = int(lo(f)), int(hi(f)), int(st(f))
                trips = max(0, _idiv(b - a + c, c))
                cell = f[var]
                cell.set(a)
                for _ in range(trips):
                    try:
                        interp._run(body, f)
                    except _Cycle:
                        pass
                    except _Exit:
                        break
                    cell.set(cell.v + c)
            return doc, pos + 1
        # assignment
        eq = self._assign_split(s)
        if eq is not None:
            lhs, rhs = eq
            val = self._expr(rhs, s)
            m = re.match(r"^(\w+)\s*\((.*)\)$", lhs)
            if m:
                name = m.group(1)
                p = _Parser(tokenize(m.group(2) + ")"), self, s)
                parsed = p.parse_args()
                subs = [a[2] for a in parsed]
                if any(a[0] == "sec" for a in parsed):
                    def set_sec(f):
                        v = val(f)
                        f[name].put_section(tuple(x(f) for x in subs), v)
                    return set_sec, pos + 1

                def set_elem(f):
                    v = val(f)
                    f[name].put(tuple(x(f) for x in subs), v)
                return set_elem, pos + 1
            name = lhs.strip()
            if not re.fullmatch(r"\w+", name):
                raise FortranError(f"cannot assign to {lhs!r} at {where}")

            def set_var(f):
                v = val(f)
                tgt = f.get(name)
                if tgt is None:
                    raise FortranError(f"assignment to undeclared {name!r} at {where}")
                if isinstance(tgt, FArray):
                    tgt.fill(v)
                else:
                    tgt.set(v)
            return set_var, pos + 1
        raise FortranError(f"statement not understood: {s!r} at {where}")

    @staticmethod
    def _paren_split(s):
        """s starts with '(' : returns (inside of the matching parentheses, the rest)."""
        depth, q = 0, None
        for i, ch in enumerate(s):
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
                if depth == 0:
                    return s[1:i], s[i + 1:]
        raise FortranError(f"unbalanced parentheses in {s!r}")

    @staticmethod
    def _assign_split(s):
        depth, q = 0, None
        for i, ch in enumerate(s):
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch in "([":
                depth += 1
            elif ch in ")]":
                depth -= 1
            elif ch == "=" and depth == 0:
                if s[i + 1:i + 2] == "=" or s[i - 1] in "<>/=":
                    continue
                return s[:i].strip(), s[i + 1:].strip()
        return None

    def _expr(self, text, stmt):
        p = _Parser(tokenize(text), self, stmt)
        e = p.parse_expr()
        if not p.at_end():
            raise FortranError(f"trailing tokens in {text!r} of {stmt!r}")
        return e

    # -- calls --------------------------------------------------------------------------------------------------------
    def actual(self, arg, f):
        """The object an actual argument associates with the dummy: the variable itself, an element reference or a temporary."""
        kind, payload, value = arg
        if kind == "var":
            x = f.get(payload)
            if x is not None:
                return x
        elif kind == "elem":
            name, subs = payload
            x = f.get(name)
            if isinstance(x, FArray):
                return ElemRef(x, tuple(a[2](f) for a in subs))
        v = value(f)
        if isinstance(v, (FArray, Cell, ElemRef)):
            return v
        if isinstance(v, list):
            t = "l" if isinstance(v[0], bool) else ("i" if isinstance(v[0], int) else "r")
            arr = FArray(t, (len(v),))
            arr.fill(v)
            return arr
        return Cell("l" if isinstance(v, bool) else ("i" if isinstance(v, int) else "r"), v)

    def _prepare(self, proc):
        ctx = {"where": proc.file}
        block, _ = self._compile_block(proc.stmts, 0, (), ctx)
        proc.compiled = block
        proc.spec = [(no, s) for no, lab, s in proc.stmts
                     if re.match(r"^use\s", s) or re.match(r"^(parameter\s*\(|save\b|data\s)", s) or
                     (_DECL.match(s) and ("::" in s or not re.match(r"^\w+\s*(\(|=)", s)))]
        proc.types = self._dummy_types(proc)

    def call(self, name, actuals, want_result=False):
        name = self.alias.get(name, name)
        proc = self.procs.get(name)
        if proc is None:
            raise FortranError(f"no procedure {name!r} is loaded")
        if proc.compiled is None:
            self._prepare(proc)
        if len(actuals) != len(proc.args):
            raise FortranError(f"{name}: {len(actuals)} arguments for {len(proc.args)} dummies")
        if proc.elemental and any(isinstance(a, FArray) for a in actuals):
            # an elemental function referenced with array arguments: element by element, in array element order
            shape = next(a.a.shape for a in actuals if isinstance(a, FArray))
            out = FArray("r", shape)
            for idx in np.ndindex(*shape):
                scal = [Cell(a.t, (float if a.t == "r" else int)(a.a[idx])) if isinstance(a, FArray) else a for a in actuals]
                out.a[idx] = self.call(name, scal, want_result=True)
            return out
        proc.calls += 1
        f = {}
        if proc.host is not None:
            for u in proc.host.uses:
                self._import(u, f)
            for k, v in proc.host.vars.items():
                if v is not None:
                    f[k] = v
        for d, a in zip(proc.args, actuals):
            f[d] = a
        result = None
        if proc.kind == "function":
            result = Cell({"integer": "i", "real": "r", "logical": "l"}.get(proc.result_type or "", None) or
                          proc.types.get(proc.result_name, "r"))
            f[proc.result_name] = result
        # specification part, in source order (an F77 PARAMETER may sit between two type statements and size an array of
        # the second).  Locals are fresh and undefined at every call; SAVE'd / DATA-initialised ones are created at the
        # first call and persist.
        first = not proc.saved.get("__initialised__")
        for no, s in proc.spec:
            if s.startswith("use"):
                self._import(s, f)
            elif re.match(r"^(save\b|data\s)", s):
                if first:
                    self._declare(s, f, frame_for_eval=f, proc=proc, where=f"{proc.file}:{no}")
            else:
                self._declare(s, f, frame_for_eval=f, dummies=proc.args, proc=proc, where=f"{proc.file}:{no}")
        proc.saved["__initialised__"] = True
        for k, v in proc.saved.items():
            if not k.startswith("__"):
                f[k] = v
        self._execute(proc.compiled, f)
        if want_result:
            if result.v is None:
                raise FortranError(f"function {name} returned without a value")
            return result.v
        return None

    def _execute(self, block, f):
        """Runs a procedure's top-level statement list; GOTO targets are labels of this list."""
        i = 0
        while i < len(block):
            lab, fn, no = block[i]
            try:
                fn(f)
            except _Return:
                return
            except _Goto as g:
                for j, (lab2, _, _) in enumerate(block):
                    if lab2 == g.label:
                        i = j
                        break
                else:
                    raise FortranError(f"GOTO {g.label}: no such label at the top level of the procedure") from None
                continue
            i += 1

    # -- statement ranges of a main program -------------------------------------------------------------------------------
    def run_block(self, path, first_line, last_line, frame):
        """Executes the statements of source lines first_line..last_line of a free-form file in `frame` (name -> Cell / FArray)."""
        key = (path, first_line, last_line)
        block = self._blocks.get(key)
        if block is None:
            text = open(path).read().splitlines()
            chunk = "\n".join(text[first_line - 1:last_line])
            stmts = logical_lines(chunk, False)
            stmts = [(no + first_line - 1, lab, s) for no, lab, s in stmts]
            block, _ = self._compile_block(stmts, 0, (), {"where": path})
            self._blocks[key] = block
        self._execute(block, frame)
