"""f90py -- executes a subset of modern Fortran by mechanical translation to Python (TEST INFRASTRUCTURE).

Purpose: the reference (HugoMVale/HR-WENO) is Fortran and no Fortran compiler exists in this image, so its
implementation cannot be run to produce golden vectors.  This module reads the reference's own SOURCE TEXT (never
copied into the repository) and translates, statement by statement, the procedures on the finite-volume path into
Python functions that are then executed on IEEE binary64 floats.  Nothing about the algorithm is restated by hand:
operation order, operator precedence and associativity, `sum()` as a sequential accumulation from zero, array sections,
lower bounds, pointer bounds remapping, optional arguments and the control flow all come from the source lines.

Arithmetic model (what a gfortran x86-64 build without -ffast-math / -march=native does): every + - * / is one
correctly rounded binary64 operation (CPython floats and NumPy float64 element-wise operations), no FMA contraction, no
re-association; `x**n` with an integer n is binary exponentiation (libgcc __powidf2), so `x**2 == x*x`; integer/integer
division truncates; array expressions are evaluated element by element.  `with real_kind(4):` translates and executes
the same source as the reference's REAL32 build (src/hrweno_kinds.F90:9-10): every real literal, scalar and array
element is then a binary32 value (np.float32), every operation one correctly rounded binary32 operation, and a binary64
value that reaches a store raises instead of being converted.

Supported subset (enough for src/hrweno_{weno,fluxes,tvdode,grids}.f90 and the `rhs`/`flux`/`ic` procedures of the two
examples): modules and programs with `contains`, derived types with default component values, type extension,
type-bound procedures, procedure-pointer components, generic interfaces naming one module procedure, subroutines and
functions (`result()`, `pure`/`elemental`, typed prefixes), optional and keyword arguments, explicit-shape / assumed-
shape / automatic arrays with arbitrary lower bounds, allocatable and pointer arrays, `associate`, `do`, `do concurrent`,
`if`, `select case`, `exit`, `cycle`, `return`, `error stop`, `allocate`, pointer assignment with bounds remapping,
array constructors, sections with strides, and the intrinsics listed in `INTRINSICS`.  ISO_C_BINDING (f90c.py): interface
bodies with `bind(c, name=)` become calls into a shared library (`Program.clib`) marshalled from their own dummy
declarations, `bind(c)` procedures can be handed to C with `c_funloc`, plus `c_loc`, `c_f_pointer`, `c_associated`,
interoperable derived types, assumed-size dummies, typed and scalar `allocate`, generic interfaces with several specific
procedures (resolved by the number of actual arguments and by their being procedures or data) -- what it takes to execute
fortran/hrweno_b200_shim.f90 under the reference's programs (tests/test_fortran_shim_exec.py).
Anything else raises `NotImplementedError` with the offending line -- it never guesses.
`sum`, `eoshift` and `real**integer` are held to gfortran's own runtime library in tests/test_gfortran_runtime_pins.py.
"""
from __future__ import annotations

import math
import re

import numpy as np

# ------------------------------------------------------------------------------------------------------------------
# runtime
# ------------------------------------------------------------------------------------------------------------------


class FortranStop(RuntimeError):
    """`error stop`"""


# The kind of `real(rk)` (src/hrweno_kinds.F90:9-17 selects it at compile time: -DREAL32 / -DREAL64).  8: scalars are
# CPython floats, arrays float64.  4: scalars are np.float32, arrays float32 -- every NumPy float32 operation is one
# correctly rounded binary32 operation, Python ints and the (few) Python floats that enter are "weak" operands (NEP 50)
# that take the float32 kind exactly as an integer / a default-real operand does in Fortran.  A binary64 value stored
# into a REAL32 run is a translator error and raises (`_store_check`).  Process-global: use `real_kind(4)` as a context.
_KIND = [8]


def rkind():
    return _KIND[0]


def rdtype():
    return np.float32 if _KIND[0] == 4 else np.float64


def rl(x):
    """a real literal / `real(x, rk)` of the current kind"""
    return np.float32(x) if _KIND[0] == 4 else float(x)


class real_kind:
    """`with real_kind(4): ...` -- translate AND execute inside the block (the translation writes real literals for the
    kind, the run-time allocates arrays of it)"""

    def __init__(self, kind):
        if kind not in (4, 8):
            raise NotImplementedError(f"real kind {kind}")
        self.kind = kind

    def __enter__(self):
        self.prev, _KIND[0] = _KIND[0], self.kind
        return self

    def __exit__(self, *exc):
        _KIND[0] = self.prev
        return False


def _store_check(value):
    if _KIND[0] == 4 and (isinstance(value, float) or (isinstance(value, np.ndarray) and value.dtype == np.float64)):
        raise TypeError(f"a binary64 value ({value!r}) reached a store in a REAL32 run")
    return value


class FS:
    """array section lo:hi:step (inclusive bounds, any of them absent)"""

    __slots__ = ("lo", "hi", "step")

    def __init__(self, lo=None, hi=None, step=None):
        self.lo, self.hi, self.step = lo, hi, step


class Ref:
    """a scalar dummy argument with intent(out) / intent(inout): passed by reference"""

    __slots__ = ("v",)

    def __init__(self, v=None):
        self.v = v


def val(x):
    return x.v if isinstance(x, Ref) else x


def as_ref(x):
    return x if isinstance(x, Ref) else Ref(x)


class FArr:
    """Fortran array: column-major NumPy storage plus one lower bound per dimension.  Sections are views."""

    __slots__ = ("a", "lb")

    def __init__(self, a, lb=None):
        if _KIND[0] == 4 and a.dtype == np.float64:
            raise TypeError("a binary64 array appeared in a REAL32 run")
        self.a = a
        self.lb = tuple(lb) if lb is not None else (1,) * a.ndim

    @staticmethod
    def alloc(bounds, dtype=None):
        dtype = rdtype() if dtype is None else dtype
        shape = tuple(max(0, hi - lo + 1) for lo, hi in bounds)
        return FArr(np.zeros(shape, dtype=dtype, order="F"), tuple(lo for lo, _ in bounds))

    @staticmethod
    def wrap(x, lb=None):
        """dummy-argument association: assumed-shape dummies get lower bound 1 (or the declared one)"""
        if x is None:
            return None
        if isinstance(x, FArr):
            return FArr(x.a, lb if lb is not None else (1,) * x.a.ndim)
        a = np.asarray(x)
        if a.dtype != rdtype() and a.dtype.kind == "f":
            if _KIND[0] == 4:  # a silent copy would also swallow what the callee writes into an intent(out) dummy
                raise TypeError("a binary64 array was passed to a procedure in a REAL32 run")
            a = a.astype(rdtype())
        return FArr(a, lb if lb is not None else (1,) * a.ndim)

    @staticmethod
    def from_list(items):
        flat = []
        for it in items:
            if isinstance(it, FArr):
                flat.extend(it.a.ravel(order="F").tolist())
            else:
                flat.append(it)
        return FArr(np.array(flat, dtype=rdtype()), (1,))

    def copy(self):
        return FArr(np.array(self.a, order="F", copy=True), self.lb)

    def rebase(self, lb):
        return FArr(self.a, tuple(lb))

    # -- indexing -------------------------------------------------------------------------------------------------
    def _key(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if len(key) != self.a.ndim:
            raise IndexError(f"rank mismatch: {len(key)} subscripts for rank {self.a.ndim}")
        out, is_section = [], False
        for d, k in enumerate(key):
            lb, n = self.lb[d], self.a.shape[d]
            if isinstance(k, FS):
                is_section = True
                step = 1 if k.step is None else int(k.step)
                lo = (lb if step > 0 else lb + n - 1) if k.lo is None else int(k.lo)
                hi = (lb + n - 1 if step > 0 else lb) if k.hi is None else int(k.hi)
                cnt = max(0, (hi - lo + step) // step)
                if cnt and not (lb <= lo <= lb + n - 1 and lb <= lo + (cnt - 1) * step <= lb + n - 1):
                    raise IndexError(f"section {lo}:{hi}:{step} outside bounds {lb}:{lb + n - 1}")
                start = lo - lb
                stop = start + cnt * step
                out.append(slice(start, stop if stop >= 0 else None, step))
            else:
                k = int(k)
                if not lb <= k <= lb + n - 1:
                    raise IndexError(f"subscript {k} outside bounds {lb}:{lb + n - 1}")
                out.append(k - lb)
        return tuple(out), is_section

    def __getitem__(self, key):
        k, sec = self._key(key)
        r = self.a[k]
        if sec:
            return FArr(r)  # a section has lower bounds 1
        if isinstance(r, np.float32):
            return r  # REAL32: the scalar keeps its kind
        return r.item() if isinstance(r, np.generic) else r

    def __setitem__(self, key, value):
        k, _ = self._key(key)
        self.a[k] = value.a if isinstance(value, FArr) else _store_check(value)

    def assign(self, value):
        if isinstance(value, FArr):
            if value.a.shape != self.a.shape:
                raise ValueError(f"shape mismatch in array assignment: {self.a.shape} <- {value.a.shape}")
            self.a[...] = value.a
        else:
            self.a[...] = _store_check(value)

    # -- element-wise arithmetic (each NumPy float64 operation is one correctly rounded IEEE operation) ---------------
    @staticmethod
    def _u(x):
        return x.a if isinstance(x, FArr) else x

    def __add__(self, o):
        return FArr(self.a + FArr._u(o))

    def __radd__(self, o):
        return FArr(FArr._u(o) + self.a)

    def __sub__(self, o):
        return FArr(self.a - FArr._u(o))

    def __rsub__(self, o):
        return FArr(FArr._u(o) - self.a)

    def __mul__(self, o):
        return FArr(self.a * FArr._u(o))

    def __rmul__(self, o):
        return FArr(FArr._u(o) * self.a)

    def __truediv__(self, o):
        return FArr(self.a / FArr._u(o))

    def __rtruediv__(self, o):
        return FArr(FArr._u(o) / self.a)

    def __neg__(self):
        return FArr(-self.a)

    def __lt__(self, o):
        return FArr(self.a < FArr._u(o))

    def __le__(self, o):
        return FArr(self.a <= FArr._u(o))

    def __gt__(self, o):
        return FArr(self.a > FArr._u(o))

    def __ge__(self, o):
        return FArr(self.a >= FArr._u(o))

    def __pos__(self):
        return self


def assign(cur, value):
    """`lhs = rhs` for a whole variable: arrays are assigned in place (dummy arguments, pointer targets), an unallocated
    allocatable is allocated to the shape of the right-hand side, scalars are rebound"""
    if isinstance(cur, Ref):
        cur.v = _store_check(val(value))
        return cur
    if isinstance(cur, FArr):
        cur.assign(value)
        return cur
    if isinstance(value, FArr):
        return value.copy()
    return _store_check(val(value))


def assign_alloc(cur, value):
    """`lhs = rhs` where lhs is an allocatable whole array (every array component of the reference's types is): F2003
    reallocates the left-hand side when the shapes differ (gfortran's default -frealloc-lhs)"""
    if isinstance(cur, FArr) and isinstance(value, FArr) and cur.a.shape != value.a.shape:
        return value.copy()
    return assign(cur, value)


def fdiv(a, b):
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)) and not isinstance(a, bool):
        q = abs(int(a)) // abs(int(b))
        return q if (a >= 0) == (b >= 0) else -q
    return a / b


def fpow(x, n):
    if isinstance(n, (int, np.integer)):
        n = int(n)
        if isinstance(x, (int, np.integer)):
            return int(x) ** n
        m = abs(n)
        y = x if (m & 1) else 1.0
        while m > 1:
            m >>= 1
            x = x * x
            if m & 1:
                y = y * x
        return 1.0 / y if n < 0 else y
    if isinstance(x, FArr):
        return FArr(np.power(x.a, FArr._u(n)))
    if _KIND[0] == 4:
        return np.power(np.float32(x), np.float32(n))
    return math.pow(x, n)


def frange(lo, hi, step=1):
    lo, hi, step = int(lo), int(hi), int(step)
    return range(lo, hi + (1 if step > 0 else -1), step)


def _seq_sum(x):
    if not isinstance(x, FArr):
        return x
    s = 0 if x.a.dtype.kind in "iu" else 0.0  # sum() accumulates from zero in array element order
    flat = x.a.ravel(order="F")
    for e in (flat if flat.dtype == np.float32 else flat.tolist()):  # float32 elements keep their kind (0.0 is weak)
        s = s + e
    return s


def _size(x, dim=None):
    return int(x.a.size if dim is None else x.a.shape[dim - 1])


def _lbound(x, dim):
    return x.lb[dim - 1]


def _ubound(x, dim):
    return x.lb[dim - 1] + x.a.shape[dim - 1] - 1


def _sign(a, b):
    if isinstance(a, np.float32) or isinstance(b, np.float32):
        return np.copysign(np.abs(np.float32(a)), np.float32(b))
    return math.copysign(abs(a), b) if isinstance(a, float) or isinstance(b, float) else (abs(a) if b >= 0 else -abs(a))


def _eoshift(array, shift, dim=1):
    out = np.zeros_like(array.a)
    ax = dim - 1
    n = array.a.shape[ax]
    src = [slice(None)] * array.a.ndim
    dst = [slice(None)] * array.a.ndim
    if shift >= 0:  # result(i) = array(i + shift)
        src[ax], dst[ax] = slice(shift, n), slice(0, n - shift)
    else:
        src[ax], dst[ax] = slice(0, n + shift), slice(-shift, n)
    out[tuple(dst)] = array.a[tuple(src)]
    return FArr(np.asfortranarray(out), array.lb)


def _minmax(fn):
    def f(*args):
        if any(isinstance(a, FArr) for a in args):
            r = FArr._u(args[0])
            for a in args[1:]:
                r = fn(r, FArr._u(a))
            return FArr(np.asarray(r, dtype=rdtype()))
        r = args[0]
        for a in args[1:]:
            r = a if (fn is np.minimum and a < r) or (fn is np.maximum and a > r) else r
        return r

    return f


def _elementwise(fn):
    return lambda x: FArr(fn(x.a)) if isinstance(x, FArr) else rl(fn(rl(x)))


def _reshape(x, shape):
    shp = [int(v) for v in (shape.a.ravel().tolist() if isinstance(shape, FArr) else shape)]
    return FArr(np.reshape(x.a, shp, order="F"))


INTRINSICS = {
    "reshape": _reshape,
    "product": lambda x: (int(np.prod(x.a)) if x.a.dtype.kind in "iu" else rl(np.prod(x.a))) if isinstance(x, FArr) else x,
    "sum": _seq_sum,
    "size": _size,
    "lbound": _lbound,
    "ubound": _ubound,
    "sign": _sign,
    "eoshift": _eoshift,
    "min": _minmax(np.minimum),
    "max": _minmax(np.maximum),
    "abs": lambda x: FArr(np.abs(x.a)) if isinstance(x, FArr) else abs(x),
    "epsilon": lambda x: rl(np.finfo(rdtype()).eps),
    "any": lambda x: bool(np.any(FArr._u(x))),
    "all": lambda x: bool(np.all(FArr._u(x))),
    "present": lambda x: x is not None,
    "allocated": lambda x: x is not None,
    "associated": lambda x: x is not None,
    "optval": lambda x, default: default if x is None else val(x),  # fortran-lang/stdlib (the reference's only use of it)
    "real": lambda x, kind=None: rl(x),
    "int": lambda x, kind=None: int(x),
    "exp": _elementwise(np.exp),
    "log": _elementwise(np.log),
    "sqrt": _elementwise(np.sqrt),
}


def callm(obj, name, /, *args, **kw):
    """`call obj%name(...)`: a type-bound procedure (passed-object first) or a procedure-pointer component (nopass)"""
    bound = type(obj)._bindings.get(name) if hasattr(type(obj), "_bindings") else None
    if bound is not None:
        return type(obj)._scope[bound](obj, *args, **kw)
    return getattr(obj, name)(*args, **kw)


# ------------------------------------------------------------------------------------------------------------------
# lexical level
# ------------------------------------------------------------------------------------------------------------------


def preprocess(text, defines=None):
    """the C-preprocessor subset Fortran sources use (`gfortran -cpp -DNAME`, src/hrweno_kinds.F90:9-17): #ifdef / #ifndef /
    #if NAME / #elif NAME / #else / #endif, #define NAME [text] / #undef, object-like macro substitution outside character
    literals.  Lines that are dropped (and the directives) become empty lines, so line numbers stay those of the file."""
    macros = dict(defines or {})
    stack, out = [], []  # stack entries: [a branch of this group has been taken, this branch is active]

    def truth(expr):
        expr = expr.strip()
        m = re.match(r"^defined\s*\(?\s*(\w+)\s*\)?$", expr)
        if m:
            return m.group(1) in macros
        if re.match(r"^\w+$", expr):
            return str(macros.get(expr, "0")).strip() not in ("0", "")
        raise NotImplementedError(f"preprocessor expression {expr!r}")

    def substitute(line):
        if not macros:
            return line
        res, q, word = "", None, ""
        for ch in line + "\0":
            if q is None and (ch.isalnum() or ch == "_"):
                word += ch
                continue
            if word:
                res += str(macros[word]) if word in macros and macros[word] is not True else word
                word = ""
            if ch == "\0":
                break
            if q:
                q = None if ch == q else q
            elif ch in "'\"":
                q = ch
            res += ch
        return res

    for raw in text.splitlines():
        st = raw.strip()
        if st.startswith("#"):
            m = re.match(r"^#\s*(ifdef|ifndef|if|elif|else|endif|define|undef)\b\s*(.*)$", st)
            if not m:
                raise NotImplementedError(f"preprocessor directive {st!r}")
            d, arg = m.groups()
            outer = all(b[1] for b in stack)
            if d in ("ifdef", "ifndef", "if"):
                on = outer and ((arg.strip() in macros) == (d == "ifdef") if d != "if" else truth(arg))
                stack.append([on, on])
            elif d in ("elif", "else"):
                outer = all(b[1] for b in stack[:-1])
                on = outer and not stack[-1][0] and (True if d == "else" else truth(arg))
                stack[-1] = [stack[-1][0] or on, on]
            elif d == "endif":
                stack.pop()
            elif outer and d == "define":
                parts = arg.split(None, 1)
                macros[parts[0]] = parts[1].strip() if len(parts) > 1 else "1"
            elif outer and d == "undef":
                macros.pop(arg.strip(), None)
            out.append("")
            continue
        out.append(substitute(raw) if all(b[1] for b in stack) else "")
    if stack:
        raise NotImplementedError("unterminated #if")
    return "\n".join(out)


def logical_lines(text, defines=None):
    """comment-free logical lines (continuations joined) with the 1-based number of their first physical line"""
    if defines is not None or re.search(r"^\s*#", text, re.M):
        text = preprocess(text, defines)
    out, buf, start = [], "", None
    for no, raw in enumerate(text.splitlines(), 1):
        line, q, i = "", None, 0
        while i < len(raw):  # strip the comment, respecting strings
            ch = raw[i]
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch == "!":
                break
            line += ch
            i += 1
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        if buf:
            line = line[1:].lstrip() if line.startswith("&") else line
        else:
            start = no
        if line.endswith("&"):
            buf += line[:-1].rstrip() + " "
            continue
        for stmt in split_top(buf + line, ";"):  # `a = 1; b = 2` is two statements
            if stmt:
                out.append((start, stmt))
        buf = ""
    return out


TOKEN = re.compile(
    r"\s*(?:(?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eEdD][+-]?\d+)?(?:_\w+)?)|(?P<str>\"[^\"]*\"|'[^']*')|"
    r"(?P<dotop>\.(?:and|or|not|true|false|eq|ne|lt|le|gt|ge|eqv|neqv)\.)|(?P<name>[A-Za-z_]\w*)|"
    r"(?P<op>\*\*|=>|==|/=|<=|>=|//|[-+*/<>=(),:%\[\]]))",
    re.I,
)


def tokenize(s):
    toks, pos = [], 0
    s = s.rstrip()
    while pos < len(s):
        m = TOKEN.match(s, pos)
        if not m:
            raise NotImplementedError(f"cannot tokenize {s[pos:]!r}")
        pos = m.end()
        kind = m.lastgroup
        toks.append((kind, m.group(kind)))
    return toks


# ------------------------------------------------------------------------------------------------------------------
# expressions
# ------------------------------------------------------------------------------------------------------------------


class ExprTranslator:
    """Fortran expression (token list) -> Python source.  Precedence (F2018 10.1.2): ** > * / > unary +- > binary +- >
    relational > .not. > .and. > .or.; ** is right-associative, the others left-associative -- Python agrees on all of
    it except that unary minus binds tighter than * in Python, which cannot change a rounded result."""

    def __init__(self, toks, scope):
        self.t, self.i, self.scope = toks, 0, scope

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def take(self, val=None):
        k, v = self.peek()
        if val is not None and (v is None or v.lower() != val):
            raise NotImplementedError(f"expected {val!r}, found {v!r} in {self.t}")
        self.i += 1
        return k, v

    def at(self, val):
        v = self.peek()[1]
        return v is not None and v.lower() == val

    def expr(self):
        left = self.and_()
        while self.at(".or."):
            self.take()
            left = f"({left} or {self.and_()})"
        return left

    def and_(self):
        left = self.not_()
        while self.at(".and."):
            self.take()
            left = f"({left} and {self.not_()})"
        return left

    def not_(self):
        if self.at(".not."):
            self.take()
            return f"(not {self.not_()})"
        return self.rel()

    REL = {"==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">=", ".eq.": "==", ".ne.": "!=", ".lt.": "<",
           ".le.": "<=", ".gt.": ">", ".ge.": ">="}

    def rel(self):
        left = self.concat()
        v = self.peek()[1]
        if v is not None and v.lower() in self.REL:
            self.take()
            return f"({left} {self.REL[v.lower()]} {self.concat()})"
        return left

    def concat(self):
        left = self.add()
        while self.at("//"):  # character concatenation: below + -, above the relational operators
            self.take()
            left = f"({left} + {self.add()})"
        return left

    def add(self):
        if self.at("-") or self.at("+"):
            op = self.take()[1]
            left = f"({op}{self.mul()})"
        else:
            left = self.mul()
        while self.at("+") or self.at("-"):
            op = self.take()[1]
            left = f"({left} {op} {self.mul()})"
        return left

    def mul(self):
        left = self.pow_()
        while self.at("*") or self.at("/"):
            op = self.take()[1]
            right = self.pow_()
            left = f"({left} * {right})" if op == "*" else f"fdiv({left}, {right})"
        return left

    def pow_(self):
        base = self.primary()
        if self.at("**"):
            self.take()
            if self.at("-") or self.at("+"):
                op = self.take()[1]
                return f"fpow({base}, {op}{self.pow_()})"
            return f"fpow({base}, {self.pow_()})"
        return base

    def arglist(self, close):
        """subscripts / actual arguments up to `close`; returns list of python snippets (sections as FS(...))"""
        args = []
        if self.at(close):
            self.take()
            return args
        while True:
            k, v = self.peek()
            nk, nv = self.t[self.i + 1] if self.i + 1 < len(self.t) else (None, None)
            if k == "name" and nv == "=" :  # keyword argument
                self.take()
                self.take()
                args.append(f"{v.lower()}={self.arg_value()}")
            else:
                args.append(self.section_or_expr(close))
            if self.at(","):
                self.take()
                continue
            self.take(close)
            return args

    def arg_value(self):
        """an actual argument: a bare by-reference scalar keeps its Ref"""
        k, v = self.peek()
        nv = self.t[self.i + 1][1] if self.i + 1 < len(self.t) else None
        if k == "name" and v.lower() in self.scope.refs and nv in (",", ")"):
            self.take()
            return v.lower()
        if k == "name" and v.lower() in getattr(self.scope, "byref_tmp", ()) and nv in (",", ")"):
            self.take()
            return self.scope.byref_tmp[v.lower()]
        return self.expr()

    def section_or_expr(self, close):
        lo = hi = step = None
        if not self.at(":"):
            lo = self.arg_value()
            if not self.at(":"):
                return lo
        self.take(":")
        if not (self.at(",") or self.at(close) or self.at(":")):
            hi = self.expr()
        if self.at(":"):
            self.take()
            step = self.expr()
        return f"FS({lo}, {hi}, {step})"

    def primary(self):
        k, v = self.take()
        if k == "num":
            v = re.sub(r"_\w+$", "", v)
            v = re.sub(r"[dD]", "e", v)
            if _KIND[0] == 4 and re.search(r"[.eE]", v):
                return f"rl({v})"  # REAL32: a real literal (default real or _rk) is a binary32 value
            return v
        if k == "str":
            return repr(v[1:-1])
        if k == "dotop":
            return {".true.": "True", ".false.": "False"}[v.lower()]
        if v == "(":
            e = self.expr()
            self.take(")")
            return f"({e})"
        if v == "[":
            if self.at("("):  # implied do: [(expr, var = lo, hi)]
                save = self.i
                try:
                    self.take("(")
                    e = self.expr()
                    self.take(",")
                    var = self.take()[1].lower()
                    self.take("=")
                    lo = self.expr()
                    self.take(",")
                    hi = self.expr()
                    self.take(")")
                    self.take("]")
                    return f"FArr.from_list([{e} for {var} in frange({lo}, {hi})])"
                except NotImplementedError:
                    self.i = save
            items = self.arglist("]")
            return f"FArr.from_list([{', '.join(items)}])"
        if k != "name":
            raise NotImplementedError(f"unexpected token {v!r} in {self.t}")
        return self.designator(v.lower())

    def c_actuals(self, name, args):
        """actual arguments of a bind(c) procedure: a scalar dummy without VALUE that the callee may define receives a
        reference to the actual (a component, a main-program variable, or an existing by-reference dummy)"""
        proto = self.scope.program.cprotos[name]
        out = []
        for pos, a in enumerate(args):
            kw, expr = (a.split("=", 1) if re.match(r"^\w+=[^=]", a) else (None, a))
            dummy = kw if kw else (proto["args"][pos] if pos < len(proto["args"]) else None)
            d = proto["decls"].get(dummy)
            if d and d["dims"] is None and not d["value"] and d["intent"] in ("out", "inout") and not re.match(r"^type\((?!c_ptr|c_funptr)", d["base"]):
                if re.match(r"^[A-Za-z_]\w*(\.\w+)+$", expr) and not expr.endswith(".v"):
                    obj, comp = expr.rsplit(".", 1)
                    expr = f"AttrRef({obj}, {comp!r})"
                elif re.match(r"^[A-Za-z_]\w*$", expr) and expr not in self.scope.refs and expr not in getattr(self.scope, "byref_tmp", {}).values():
                    unit = self.scope.unit
                    if unit.get("kind") == "program" or expr in unit.get("host", ()) and expr not in self.scope.locals:
                        expr = f"GRef(globals(), {expr!r})"
                    else:
                        raise NotImplementedError(f"{name}: local variable {expr} as an intent({d['intent']}) actual inside an expression "
                                                  "(use it in a CALL statement, or make it a component / program variable)")
            out.append(f"{kw}={expr}" if kw else expr)
        return out

    def designator(self, name):
        cur = self.scope.rename(name)
        first = True
        while True:
            if self.at("("):
                self.take()
                is_call = first and self.scope.is_callable(name)  # decided before the arguments are translated
                args = self.arglist(")")
                if is_call and name in getattr(self.scope.program, "cprotos", {}):
                    args = self.c_actuals(name, args)
                if is_call:
                    cur = f"{cur}({', '.join(args)})"
                else:
                    cur = f"{cur}[{', '.join(args)}]"
            elif self.at("%"):
                self.take()
                comp = self.take()[1].lower()
                if self.at("(") and self.scope.is_method(comp):
                    self.take()
                    args = self.arglist(")")
                    cur = f"callm({cur}, {comp!r}{''.join(', ' + a for a in args)})"
                else:
                    cur = f"{cur}.{comp}"
            else:
                return cur
            first = False


# ------------------------------------------------------------------------------------------------------------------
# program units
# ------------------------------------------------------------------------------------------------------------------


class Scope:
    def __init__(self, unit, program):
        self.unit, self.program = unit, program
        self.refs = set()      # scalar dummies passed by reference: read as name.v
        self.result = None     # (fortran name, python name) of the function result
        self.locals = set()
        self.scalar_locals = set()  # non-dummy real / integer / logical scalars (passed by reference in CALLs)
        self.byref_tmp = {}

    def rename(self, name):
        if self.result and name == self.result[0]:
            return self.result[1]
        if name in self.refs:
            return f"{name}.v"
        if name in PY_RESERVED:
            return name + "_"
        return name

    def is_callable(self, name):
        if name in self.locals and name not in self.unit["proc_dummies"]:
            return False
        return (name in INTRINSICS or name in self.program.procs or name in self.program.generics or name in self.unit["proc_dummies"]
                or name in self.program.cprotos or name in C_INTRINSICS)

    def is_method(self, comp):
        return comp in self.program.methods


def PY_RESERVED_SAFE(name):
    return name + "_" if name in PY_RESERVED else name


C_INTRINSICS = {"c_loc", "c_funloc", "c_associated", "transfer"}  # ISO_C_BINDING (+ transfer), supplied by f90c.install

PY_RESERVED = {"lambda", "from", "in", "is", "not", "pass", "def", "class", "global", "with", "as", "del", "try"}

DECL = re.compile(r"^(real|integer|logical|character|type|class|procedure)\b", re.I)
UNIT_HEAD = re.compile(
    r"^(?P<prefix>(?:(?:pure|elemental|impure|recursive|module)\s+|(?:real|integer|logical|type|class)\s*\([^)]*\)\s+|"
    r"(?:logical|integer|real)\s+)*)(?P<kind>subroutine|function)\s+(?P<name>\w+)\s*(?:\((?P<args>[^)]*)\))?\s*"
    r"(?P<bind1>bind\s*\([^)]*\))?\s*(?:result\s*\(\s*(?P<res>\w+)\s*\))?\s*(?P<bind2>bind\s*\([^)]*\))?\s*$",
    re.I,
)


def split_top(s, sep=","):
    """split at top-level separators (not inside parentheses / brackets / strings)"""
    out, depth, cur, q = [], 0, "", None
    for ch in s:
        if q:
            cur += ch
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


class Program:
    """all program units of a set of source files, translated into one Python namespace"""

    def __init__(self, skip_io=False, defines=None):
        self.skip_io = skip_io  # `write` / `print` statements become no-ops (test programs print diagnostics)
        self.defines = defines  # preprocessor symbols ({"REAL32": "1"} = gfortran -cpp -DREAL32)
        self.procs = {}     # name -> unit
        self.generics = {}  # generic interface name -> specific procedure
        self.types = {}     # type name -> dict(parent, comps=[(name, default_src)], bindings={})
        self.methods = set()
        self.params = []    # module-level parameter declarations (file, line, text)
        self.ns = {
            "FArr": FArr, "FS": FS, "Ref": Ref, "val": val, "as_ref": as_ref, "assign": assign, "assign_alloc": assign_alloc, "fdiv": fdiv, "fpow": fpow,
            "frange": frange, "callm": callm, "FortranStop": FortranStop, **INTRINSICS,
        }
        self.sources = {}
        self.cprotos = {}    # bind(c) interface bodies: Fortran name -> prototype (see f90c.py)
        self.absifaces = {}  # abstract interface bodies, same form
        import f90c

        self.interop = f90c.install(self)  # ISO_C_BINDING: kinds, c_loc / c_funloc / c_f_pointer, bind(c) calls

    @property
    def clib(self):
        return self.interop.lib

    @clib.setter
    def clib(self, lib):
        """the shared library (ctypes.CDLL) the program's bind(c) interfaces resolve against"""
        self.interop.lib = lib

    # -- parsing ----------------------------------------------------------------------------------------------------
    def add_source(self, path, skip=()):
        """parse one source file; procedures named in `skip` (I/O, timers) are not translated"""
        self._skip = {x.lower() for x in skip}
        self._host = set()
        text = open(path).read()
        lines = logical_lines(text, self.defines)
        self.sources[path] = lines
        i, n = 0, len(lines)
        stack = []  # enclosing module / program / procedures
        while i < n:
            no, ln = lines[i]
            low = ln.lower()
            m = UNIT_HEAD.match(ln)
            if re.match(r"^program\s+\w+$", low):
                i = self._parse_program(path, lines, i)
            elif re.match(r"^module\s+\w+$", low) and not low.startswith("module procedure"):
                stack.append(("container", low.split()[1]))
                i += 1
            elif re.match(r"^(abstract\s+)?interface\b", low):
                gen = low.split()[1] if len(low.split()) > 1 and not low.startswith("abstract") else None
                i += 1
                while not lines[i][1].lower().startswith("end interface"):
                    mm = re.match(r"^module\s+procedure\s*(?:::)?\s*(.*)$", lines[i][1], re.I)
                    if gen and mm:  # one specific: an alias; several: resolved by the kind of the actual arguments at the call
                        self.generics.setdefault(gen, []).extend(x.strip().lower() for x in mm.group(1).split(","))
                        i += 1
                    elif (mh := UNIT_HEAD.match(lines[i][1])) and not gen:
                        i = self._parse_interface_body(path, lines, i, mh, low.startswith("abstract"))
                    else:
                        i += 1
                i += 1
            elif re.match(r"^type\s*(,[^:]*)?(::)?\s*\w+$", low) and not low.startswith("type("):
                i = self._parse_type(lines, i)
            elif m:
                i = self._parse_unit(path, lines, i, m)
            elif low.startswith("end ") or low == "end" or low == "contains":
                i += 1
            else:
                if DECL.match(ln) and "::" in ln and ("parameter" in low.split("::")[0] or (stack and "=" in ln.split("::", 1)[1])):
                    self.params.append((path, no, ln))  # named constants, and module variables with an initial value
                    if "parameter" not in low.split("::")[0]:  # module variables: the module's procedures may define them
                        self._host = set(self._host) | {nm for nm, _, _, _ in self._decl_entities(ln)}
                i += 1
        return self

    def _parse_type(self, lines, i):
        head = lines[i][1]
        name = re.split(r"::|\s", head.strip())[-1].lower()
        mext = re.search(r"extends\s*\(\s*(\w+)\s*\)", head, re.I)
        td = {"parent": mext.group(1).lower() if mext else None, "comps": [], "bindings": {}, "cspec": {},
              "bindc": bool(re.search(r"bind\s*\(\s*c\s*\)", head, re.I))}
        i += 1
        in_contains = False
        while not re.match(r"^end\s*type", lines[i][1], re.I):
            ln = lines[i][1]
            if ln.lower() == "contains":
                in_contains = True
            elif in_contains:
                mm = re.match(r"^procedure[^:]*::\s*(.*)$", ln, re.I)
                if mm:
                    for ent in split_top(mm.group(1)):
                        if "=>" in ent:
                            b, p = [x.strip().lower() for x in ent.split("=>")]
                        else:
                            b = p = ent.strip().lower()
                        td["bindings"][b] = p
                        self.methods.add(b)
            else:
                spec, ents = ln.split("::", 1)
                for ent in split_top(ents):
                    mm = re.match(r"^(\w+)\s*(\([^)]*\))?\s*(?:(=>|=)\s*(.*))?$", ent)
                    cname, default = mm.group(1).lower(), mm.group(4)
                    if mm.group(3) == "=>" or default is None:
                        default = None
                    td["comps"].append((cname, default))
                    td["cspec"][cname] = (split_top(spec)[0].strip(), mm.group(2)[1:-1] if mm.group(2) else None)
                    if spec.lower().startswith("procedure"):
                        self.methods.add(cname)
            i += 1
        self.types[name] = td
        return i + 1

    def _parse_program(self, path, lines, i):
        """the main program: translated into `main_<name>()`, whose variables are module-level in the namespace so that
        the internal procedures after `contains` see them by host association"""
        name = "main_" + lines[i][1].split()[1].lower()
        unit = {"name": name, "kind": "program", "args": [], "res": None, "prefix": "", "decls": [], "body": [], "path": path,
                "proc_dummies": set()}
        i += 1
        while True:
            no, ln = lines[i]
            low = ln.lower()
            if low == "contains" or re.match(r"^end\s*program", low) or low == "end":
                break
            if low.startswith(("use ", "use,", "implicit ")):
                pass
            elif not unit["body"] and DECL.match(ln) and "::" in ln:
                unit["decls"].append((no, ln))
            else:
                unit["body"].append((no, ln))
            i += 1
        self.procs[name] = unit
        # internal procedures after `contains` reach the program's variables by host association
        self._host = {nm for _, ln in unit["decls"] for nm, _, _, _ in self._decl_entities(ln)}
        return i + (0 if lines[i][1].lower() == "contains" else 1)

    def _parse_unit(self, path, lines, i, m):
        name = m.group("name").lower()
        args = [a.strip().lower() for a in (m.group("args") or "").split(",") if a.strip()]
        unit = {"name": name, "kind": m.group("kind").lower(), "args": args, "res": (m.group("res") or name).lower(),
                "prefix": (m.group("prefix") or "").lower(), "decls": [], "body": [], "path": path, "proc_dummies": set(),
                "host": set(getattr(self, "_host", ())), "bindc": bool(m.group("bind1") or m.group("bind2"))}
        i += 1
        depth = 0
        while True:
            no, ln = lines[i]
            low = ln.lower()
            if depth == 0 and re.match(rf"^end\s*({unit['kind']})?(\s+{name})?$", low):
                break
            if depth == 0 and low == "contains":
                raise NotImplementedError(f"{path}:{no}: internal procedures inside a procedure")
            is_decl = not unit["body"] and (DECL.match(ln) and "::" in ln or low.startswith(("use ", "implicit ", "import ")))
            if is_decl:
                if DECL.match(ln):
                    unit["decls"].append((no, ln))
            else:
                unit["body"].append((no, ln))
            i += 1
        if name not in self._skip:
            self.procs[name] = unit
        return i + 1

    def _parse_interface_body(self, path, lines, i, m, abstract):
        """one interface body: the dummy declarations are the prototype (nothing is executed).  bind(c, name="x") bodies
        become callables into the attached library at build time; abstract ones are kept for reference."""
        name = m.group("name").lower()
        bind = m.group("bind1") or m.group("bind2")
        args = [a.strip().lower() for a in (m.group("args") or "").split(",") if a.strip()]
        proto = {"name": name, "kind": m.group("kind").lower(), "args": args, "res": (m.group("res") or name).lower(), "decls": {},
                 "cname": None, "path": path, "line": lines[i][0], "prefix": (m.group("prefix") or "").lower()}
        if bind:
            mn = re.search(r"name\s*=\s*[\"']([^\"']*)[\"']", bind)
            proto["cname"] = mn.group(1) if mn else name
        i += 1
        while not re.match(rf"^end\s*({proto['kind']})?(\s+{name})?$", lines[i][1].lower()):
            ln = lines[i][1]
            if DECL.match(ln) and "::" in ln:
                proto["decls"].update(self._cdecls(ln))
            i += 1
        missing = [a for a in args if a not in proto["decls"]]
        if missing:
            raise NotImplementedError(f"{path}:{proto['line']}: interface body {name}: dummies without a declaration: {missing}")
        if abstract or not bind:
            self.absifaces[name] = proto
        else:
            self.cprotos[name] = proto
        return i + 1

    def _cdecls(self, ln):
        """declaration line -> {name: dict(base, dims, value, intent)} (what interoperability needs to know)"""
        out = {}
        for nm, dims, _, info in self._decl_entities(ln):
            out[nm] = {"base": re.sub(r"\s", "", info["base"]), "dims": dims, "value": info["value"], "intent": info["intent"]}
        return out

    # -- code generation ----------------------------------------------------------------------------------------------
    def ex(self, src, scope):
        tr = ExprTranslator(tokenize(src), scope)
        out = tr.expr()
        if tr.i != len(tr.t):
            raise NotImplementedError(f"trailing tokens in expression {src!r}")
        return out

    def _decl_entities(self, ln):
        spec, ents = ln.split("::", 1)
        attrs = [a.strip() for a in split_top(spec)]
        base = attrs[0].lower()
        info = {"base": base, "intent": None, "optional": False, "dims": None, "parameter": False, "pointer": False,
                "allocatable": False, "value": False}
        for a in attrs[1:]:
            al = a.lower()
            if al.startswith("intent"):
                info["intent"] = re.sub(r"\s", "", al)[7:-1]
            elif al == "optional":
                info["optional"] = True
            elif al.startswith("dimension"):
                info["dims"] = a[a.index("(") + 1 : a.rindex(")")]
            elif al == "parameter":
                info["parameter"] = True
            elif al == "pointer":
                info["pointer"] = True
            elif al == "allocatable":
                info["allocatable"] = True
            elif al == "value":
                info["value"] = True
        out = []
        for ent in split_top(ents):
            mm = re.match(r"^(\w+)\s*(?:\((.*?)\))?\s*(?:=\s*(.*))?$", ent, re.S)
            if not mm:
                raise NotImplementedError(f"declaration entity {ent!r}")
            # the dims group must be balanced: re-split by hand for nested parentheses
            nm = mm.group(1).lower()
            rest = ent[len(mm.group(1)) :].strip()
            dims, init = None, None
            if rest.startswith("("):
                depth = 0
                for j, ch in enumerate(rest):
                    depth += ch == "("
                    depth -= ch == ")"
                    if depth == 0:
                        dims, rest = rest[1:j], rest[j + 1 :].strip()
                        break
            if rest.startswith("="):
                init = rest[1:].strip()
            out.append((nm, dims if dims is not None else info["dims"], init, info))
        return out

    def _bounds(self, dims, scope):
        b = []
        for d in split_top(dims):
            if d.strip() == "*":  # assumed size: lower bound 1, the extent is the actual's
                b.append(("1", None))
            elif ":" in d:
                lo, hi = d.split(":", 1)
                b.append((self.ex(lo, scope) if lo.strip() else None, self.ex(hi, scope) if hi.strip() else None))
            else:
                b.append(("1", self.ex(d, scope)))
        return b

    def gen_unit(self, unit):
        scope = Scope(unit, self)
        py = []
        emit = lambda ind, s: py.append("    " * ind + s)  # noqa: E731
        is_fn = unit["kind"] == "function"
        args = list(unit["args"])
        decls = {}
        for no, ln in unit["decls"]:
            for nm, dims, init, info in self._decl_entities(ln):
                decls[nm] = (dims, init, info, no)
                scope.locals.add(nm)
                if info["base"].startswith("procedure") and nm in args:
                    unit["proc_dummies"].add(nm)
        if unit.get("bindc"):
            unit["cdecls"] = {}
            for no, ln in unit["decls"]:
                unit["cdecls"].update(self._cdecls(ln))
        for nm, (dims, init, info, no) in decls.items():
            if nm not in args and dims is None and not info["parameter"] and info["base"].startswith(("real", "integer", "logical")):
                scope.scalar_locals.add(nm)
        res = unit["res"] if is_fn else None
        if is_fn:
            scope.result = (res, "res_")
            scope.locals.add(res)
        for a in args:  # scalar dummies with intent(out|inout) are references
            if a in decls:
                dims, _, info, _ = decls[a]
                if dims is None and info["intent"] in ("out", "inout") and info["base"].startswith(("real", "integer", "logical")):
                    scope.refs.add(a)
        sig = ", ".join(scope.rename(a).replace(".v", "") + ("=None" if a in decls and decls[a][2]["optional"] else "") for a in args)
        # optional dummies must come last for Python: the reference's procedures already satisfy that or are called by keyword
        emit(0, f"def {unit['name']}({sig}):")
        emit(1, f"# {unit['path']}:{unit['decls'][0][0] if unit['decls'] else unit['body'][0][0]}")
        if unit["kind"] == "program" and decls:
            emit(1, "global " + ", ".join(scope.rename(nm) for nm in decls))
        host = sorted(unit.get("host", set()) - set(decls) - set(args) - {unit["name"], res or ""})
        if host:  # host-associated variables may be defined here: they live at module level of the namespace
            emit(1, "global " + ", ".join(PY_RESERVED_SAFE(nm) for nm in host))
        for a in args:
            if a not in decls:
                continue
            dims, _, info, _ = decls[a]
            if a in scope.refs:
                emit(1, f"{a} = as_ref({a})")
            elif dims is not None:
                lbs = []
                for lo, hi in self._bounds(dims, scope):
                    lbs.append(lo if lo is not None else "1")
                emit(1, f"{a} = FArr.wrap({a}, ({', '.join(lbs)},))")
            elif info["base"].startswith(("real", "integer", "logical")):
                emit(1, f"{a} = val({a})")
        if is_fn:
            rb = decls.get(res, (None, None, {"base": unit["prefix"]}, 0))[2]["base"]
            tm = re.search(r"type\s*\(\s*(\w+)\s*\)", rb + " " + unit["prefix"])
            emit(1, f"res_ = new_{tm.group(1).lower()}()" if tm else "res_ = None")
        for nm, (dims, init, info, no) in decls.items():
            if nm in args or nm == res:
                continue
            if info["base"].startswith(("type", "class")):
                tm = re.search(r"\(\s*(\w+)\s*\)", info["base"])
                if dims is not None:  # an array of objects
                    (lo, hi), = self._bounds(dims, scope)
                    emit(1, f"{scope.rename(nm)} = FArr(np.array([new_{tm.group(1).lower()}() for _ in frange({lo}, {hi})], dtype=object), ({lo},))")
                else:
                    emit(1, f"{scope.rename(nm)} = new_{tm.group(1).lower()}()")
            elif dims is not None and not info["pointer"] and not info["allocatable"] and ":" not in [d.strip() for d in split_top(dims)]:
                bs = ", ".join(f"({lo}, {hi})" for lo, hi in self._bounds(dims, scope))
                dt = "np.int64" if info["base"].startswith("integer") else "None"
                emit(1, f"{scope.rename(nm)} = FArr.alloc([{bs}], {dt})")
                if init is not None:
                    emit(1, f"{scope.rename(nm)}.assign({self.ex(init, scope)})")
            else:
                emit(1, f"{scope.rename(nm)} = {self.ex(init, scope) if init is not None else 'None'}")
        self._gen_body(unit, scope, emit)
        emit(1, "return res_" if is_fn else "return None")
        return "\n".join(py)

    def _entity_type(self, designator, unit):
        """declared type specifier (lower case, no blanks) of `name` or of the last component of `a%b%c`"""
        last = designator.split("%")[-1].strip()
        last = re.sub(r"\(.*\)$", "", last)
        if "%" not in designator:
            for _, ln in unit["decls"]:
                for nm, _, _, info in self._decl_entities(ln):
                    if nm == last:
                        return re.sub(r"\s", "", info["base"])
        found = {re.sub(r"\s", "", td["cspec"][last][0].lower()) for td in self.types.values() if last in td["cspec"]}
        if len(found) != 1:
            raise NotImplementedError(f"cannot determine the declared type of {designator!r}: {sorted(found)}")
        return found.pop()

    def _assignment(self, ln, scope):
        """split `lhs = rhs` / `lhs => rhs` at the top-level operator"""
        depth, q = 0, None
        for j, ch in enumerate(ln):
            if q:
                q = None if ch == q else q
                continue
            if ch in "'\"":
                q = ch
            elif ch in "([":
                depth += 1
            elif ch in ")]":
                depth -= 1
            elif ch == "=" and depth == 0:
                if ln[j + 1 : j + 2] == ">":
                    return ln[:j].strip(), "=>", ln[j + 2 :].strip()
                if ln[j + 1 : j + 2] == "=" or ln[j - 1] in "<>/=":
                    continue
                return ln[:j].strip(), "=", ln[j + 1 :].strip()
        return None

    def _gen_stmt(self, no, ln, scope, emit, ind, unit):
        low = ln.lower()
        ret = "return res_" if unit["kind"] == "function" else "return None"
        if low == "return":
            return emit(ind, ret)
        if low == "exit":
            return emit(ind, "break")
        if low == "cycle":
            return emit(ind, "continue")
        if self.skip_io and re.match(r"^(write\s*\(|print\b)", low):
            return emit(ind, "pass")
        if low.startswith("error stop"):
            arg = ln[10:].strip()
            return emit(ind, f"raise FortranStop({self.ex(arg, scope) if arg else repr('error stop')})")
        if m := re.match(r"^call\s+c_f_pointer\s*\((.*)\)$", ln, re.I):  # the pointer is defined by the call
            a = split_top(m.group(1))
            what = self._entity_type(a[1].strip().lower(), unit)
            kind = "obj" if what.startswith(("type", "class")) else ("char" if what.startswith("character") else what.split("(")[0])
            shape = self.ex(a[2], scope) if len(a) > 2 else "None"
            return emit(ind, f"{self.ex(a[1], scope)} = c_f_pointer_({self.ex(a[0], scope)}, {kind!r}, {shape})")
        if low.startswith("call "):
            target = ln[5:].strip()
            if "(" not in target:
                target += "()"
            # local scalar variables given as bare actual arguments are passed by reference (the callee may define them)
            byref = {}
            inner = target[target.index("(") + 1 : target.rindex(")")] if "(" in target else ""
            for a in split_top(inner):
                a = a.strip().lower()
                if a in scope.scalar_locals and a not in scope.refs:
                    byref[a] = f"{a}__ref"
                    emit(ind, f"{a}__ref = Ref({scope.rename(a)})")
            scope.byref_tmp = byref
            flat, depth = "", 0  # the designator with its parenthesised groups removed: a%b(..)%c(..) -> a%b%c
            for ch in target:
                depth += ch == "("
                if depth == 0:
                    flat += ch
                depth -= ch == ")"
            tr = ExprTranslator(tokenize(target), CallScope(scope) if "%" not in flat else scope)
            emit(ind, tr.expr())
            for a, tmp in byref.items():
                emit(ind, f"{scope.rename(a)} = {tmp}.v")
            scope.byref_tmp = {}
            return None
        if low.startswith("allocate"):
            inner = ln[ln.index("(") + 1 : ln.rindex(")")]
            tspec = None
            if "::" in inner:  # allocate (character(n) :: msg)
                tspec, inner = [x.strip() for x in inner.split("::", 1)]
                if not tspec.lower().startswith("character"):
                    raise NotImplementedError(f"{unit['path']}:{no}: typed allocation of {tspec}")
            for ent in split_top(inner):
                if tspec:
                    emit(ind, f"{self.ex(ent, scope)} = ''")
                    continue
                if not ent.rstrip().endswith(")"):  # a scalar pointer / allocatable of derived type
                    what = self._entity_type(ent.strip().lower(), unit)
                    tm = re.match(r"^(?:type|class)\s*\(\s*(\w+)\s*\)", what)
                    if not tm:
                        raise NotImplementedError(f"{unit['path']}:{no}: allocate of scalar {ent} ({what})")
                    emit(ind, f"{self.ex(ent, scope)} = new_{tm.group(1)}()")
                    continue
                # the allocation shape is the last parenthesised group
                depth, k = 0, len(ent) - 1
                while k >= 0:
                    depth += ent[k] == ")"
                    depth -= ent[k] == "("
                    if depth == 0:
                        break
                    k -= 1
                target, dims = ent[:k].strip(), ent[k + 1 : -1]
                bs = ", ".join(f"({lo}, {hi})" for lo, hi in self._bounds(dims, scope))
                emit(ind, f"{self.ex(target, scope)} = FArr.alloc([{bs}])")
            return None
        if low.startswith(("deallocate", "nullify")):
            inner = ln[ln.index("(") + 1 : ln.rindex(")")]
            for ent in split_top(inner):
                emit(ind, f"{self.ex(ent, scope)} = None")
            return None
        asg = self._assignment(ln, scope)
        if asg is None:
            raise NotImplementedError(f"{unit['path']}:{no}: unsupported statement {ln!r}")
        lhs, op, rhs = asg
        r = self.ex(rhs, scope)
        if op == "=>":
            mm = re.match(r"^([\w%]+)\s*\((.*)\)$", lhs)
            if mm and ":" in mm.group(2):  # bounds remapping  p(lb:) => target
                lbs = [self.ex(d.split(":")[0], scope) for d in split_top(mm.group(2))]
                return emit(ind, f"{self.ex(mm.group(1), scope)} = ({r}).rebase(({', '.join(lbs)},))")
            return emit(ind, f"{self.ex(lhs, scope)} = {r}")
        l = self.ex(lhs, scope)
        if l.endswith("]"):  # element or section
            return emit(ind, f"{l} = {r}")
        if l.endswith(".v"):
            return emit(ind, f"{l} = val({r})")
        if "." in l:
            obj, comp = l.rsplit(".", 1)
            return emit(ind, f"{l} = assign_alloc(getattr({obj}, {comp!r}, None), {r})")
        return emit(ind, f"{l} = assign({l}, {r})")

    def _gen_body(self, unit, scope, emit):
        ind = 1
        sel = []  # select-case stack: [tmp name, first-case flag]
        for no, ln in unit["body"]:
            low = ln.lower()
            try:
                if m := re.match(r"^if\s*\((.*)\)\s*then$", ln, re.I):
                    emit(ind, f"if {self.ex(m.group(1), scope)}:")
                    ind += 1
                elif m := re.match(r"^else\s*if\s*\((.*)\)\s*then$", ln, re.I):
                    emit(ind - 1, f"elif {self.ex(m.group(1), scope)}:")
                elif low == "else":
                    emit(ind - 1, "else:")
                elif re.match(r"^end\s*if$", low):
                    emit(ind, "pass")
                    ind -= 1
                elif re.match(r"^end\s*do$", low):
                    emit(ind, "pass")
                    ind -= 1 + self._do_stack.pop()  # a multi-index `do concurrent` opened several nested loops
                elif low.startswith("if") and (m := self._one_line_if(ln)):
                    emit(ind, f"if {self.ex(m[0], scope)}:")
                    self._gen_stmt(no, m[1], scope, emit, ind + 1, unit)
                elif m := re.match(r"^do\s+while\s*\((.*)\)$", ln, re.I):
                    emit(ind, f"while {self.ex(m.group(1), scope)}:")
                    self._do_stack.append(0)
                    ind += 1
                elif low == "do":
                    emit(ind, "while True:")
                    self._do_stack.append(0)
                    ind += 1
                elif m := re.match(r"^do\s+concurrent\s*\((.*)\)$", ln, re.I):
                    specs = split_top(m.group(1))
                    for s in specs:
                        var, rng = s.split("=", 1)
                        parts = rng.split(":")
                        a = ", ".join(self.ex(p, scope) for p in parts)
                        emit(ind, f"for {var.strip().lower()} in frange({a}):")
                        ind += 1
                    self._do_stack.append(len(specs) - 1)
                elif m := re.match(r"^do\s+(\w+)\s*=\s*(.*)$", ln, re.I):
                    a = ", ".join(self.ex(p, scope) for p in split_top(m.group(2)))
                    emit(ind, f"for {m.group(1).lower()} in frange({a}):")
                    self._do_stack.append(0)
                    ind += 1
                elif m := re.match(r"^select\s+case\s*\((.*)\)$", ln, re.I):
                    tmp = f"sel_{len(sel)}_{no}"
                    emit(ind, f"{tmp} = {self.ex(m.group(1), scope)}")
                    sel.append([tmp, True])
                    ind += 1
                elif m := re.match(r"^case\s*\((.*)\)$", ln, re.I):
                    tmp, first = sel[-1]
                    emit(ind - 1, f"{'if' if first else 'elif'} {tmp} == {self.ex(m.group(1), scope)}:")
                    emit(ind, "pass")
                    sel[-1][1] = False
                elif low == "case default":
                    emit(ind - 1, "else:" if not sel[-1][1] else "if True:")
                    emit(ind, "pass")
                elif re.match(r"^end\s*select$", low):
                    sel.pop()
                    ind -= 1
                elif m := re.match(r"^associate\s*\((.*)\)$", ln, re.I):
                    for s in split_top(m.group(1)):
                        a, e = s.split("=>")
                        scope.locals.add(a.strip().lower())
                        emit(ind, f"{a.strip().lower()} = {self.ex(e.strip(), scope)}")
                elif re.match(r"^end\s*associate$", low):
                    pass
                else:
                    self._gen_stmt(no, ln, scope, emit, ind, unit)
            except NotImplementedError:
                raise
            except Exception as e:  # pragma: no cover - translator bug: point at the line
                raise NotImplementedError(f"{unit['path']}:{no}: {ln!r}: {type(e).__name__}: {e}") from e

    _do_stack = []

    @staticmethod
    def _one_line_if(ln):
        depth = 0
        start = ln.index("(")
        for j in range(start, len(ln)):
            depth += ln[j] == "("
            depth -= ln[j] == ")"
            if depth == 0:
                rest = ln[j + 1 :].strip()
                return (ln[start + 1 : j], rest) if rest and rest.lower() != "then" else None
        return None

    def build(self):
        """translate everything and exec it into the namespace; returns the namespace"""
        ns = self.ns
        ns["np"] = np
        ns.setdefault("rk", rkind())  # the kind parameter of hrweno_kinds; only ever used as a kind argument
        ns["rl"] = rl
        dummy = Scope({"proc_dummies": set()}, self)
        # derived types -> Python classes with the declared default component values
        for tname, td in self.types.items():
            comps, bindings, t = [], {}, td
            chain = []
            while t is not None:
                chain.append(t)
                t = self.types.get(t["parent"]) if t["parent"] else None
            for t in reversed(chain):
                comps.extend(t["comps"])
                bindings.update(t["bindings"])
            cls = type(tname, (), {"_bindings": bindings, "_scope": ns, "_comps": comps})

            cspec = {}
            for t in reversed(chain):
                cspec.update(t.get("cspec", {}))

            def make(cls=cls, comps=comps, cspec=cspec):
                def new():
                    o = cls()
                    for cname, default in comps:
                        v = eval(self.ex(default, dummy), ns) if default is not None else None
                        spec, cdims = cspec.get(cname, ("", None))
                        spec = spec.lower()
                        if v is not None and not isinstance(v, FArr) and cdims is not None and ":" not in cdims:
                            arr = FArr.alloc([(int(eval(lo, ns)), int(eval(hi, ns))) for lo, hi in self._bounds(cdims, dummy)],
                                             np.int64 if isinstance(v, int) and not isinstance(v, bool) else None)
                            arr.assign(v)  # a scalar default of an explicit-shape array component is broadcast
                            v = arr
                        if isinstance(v, FArr) and re.match(r"^(integer|type\s*\(\s*c_(fun)?ptr)", spec):
                            v = FArr(v.a.astype(np.int64), v.lb)  # integer / c_ptr array components hold integers
                        setattr(o, cname, v)
                    return o

                return new

            ns[f"new_{tname}"] = make()
        # module-level named constants
        for path, no, ln in self.params:
            for nm, dims, init, info in self._decl_entities(ln):
                if dims is not None:
                    m = re.match(r"^reshape\s*\(\s*\[(.*)\]\s*,\s*\[(.*?)\]\s*(?:,\s*order\s*=\s*\[(.*?)\])?\s*\)$", init, re.I | re.S)
                    bnds = [(int(eval(lo, ns)), int(eval(hi, ns))) for lo, hi in self._bounds(dims, dummy)]
                    arr = FArr.alloc(bnds)
                    if m:
                        vals = [eval(self.ex(v, dummy), ns) for v in split_top(m.group(1))]
                        order = [int(x) for x in split_top(m.group(3))] if m.group(3) else list(range(1, len(bnds) + 1))
                        shape = [int(x) for x in split_top(m.group(2))]
                        perm_shape = [shape[o - 1] for o in order]  # element order varies fastest along order(1)
                        tmp = np.array(vals, dtype=rdtype()).reshape(perm_shape, order="F")
                        arr.a[...] = np.transpose(tmp, np.argsort([o - 1 for o in order]))
                    else:
                        v = eval(self.ex(init, dummy), ns)
                        arr.assign(v)
                    ns[nm] = arr
                else:
                    ns[nm] = eval(self.ex(init, dummy), ns)
        code = {}
        for name, unit in self.procs.items():
            Program._do_stack = []
            src = self.gen_unit(unit)
            code[name] = src
            exec(compile(src, f"<f90py:{unit['path']}:{name}>", "exec"), ns)
            ns[name]._f90unit = unit
            if "elemental" in unit["prefix"]:
                ns[name] = _elemental(ns[name])
        for name, proto in self.cprotos.items():
            ns[name] = self.interop.cfunc(proto)
        for gen, specs in self.generics.items():
            specs = [sp for sp in specs if sp in ns]
            if len(specs) == 1:
                ns[gen] = ns[specs[0]]
            elif specs:
                ns[gen] = _generic(gen, [(ns[sp], self.procs[sp]) for sp in specs])
        self.code = code
        return ns


class CallScope:
    """scope view used for the target of `call name(args)`: the leading name is a procedure even if nothing declares it"""

    def __init__(self, scope):
        self._s = scope
        self.refs = scope.refs
        self.byref_tmp = scope.byref_tmp
        self.program, self.unit, self.locals = scope.program, scope.unit, scope.locals
        self._first = True

    def rename(self, name):
        return self._s.rename(name)

    def is_callable(self, name):
        if self._first:  # the designator translator asks once for the leading name, then for names inside the arguments
            self._first = False
            return True
        return self._s.is_callable(name)

    def is_method(self, comp):
        return self._s.is_method(comp)


def _generic(gen, specs):
    """a generic interface with several specific procedures: the specific whose dummies agree with the actual arguments in
    number and in being a procedure or a data object (what distinguishes rktvd(fu, neq, order) from rktvd(fv, neq, order))"""

    def call(*args, **kw):
        hits = []
        for fn, unit in specs:
            names = unit["args"]
            if len(args) + len(kw) > len(names) or any(k not in names[len(args):] for k in kw):
                continue
            actual = {**dict(zip(names, args)), **kw}
            if all((nm in unit["proc_dummies"]) == callable(a) for nm, a in actual.items()):
                hits.append(fn)
        if len(hits) != 1:
            raise TypeError(f"generic {gen}: {len(hits)} specific procedures match the actual arguments")
        return hits[0](*args, **kw)

    return call


def _elemental(fn):
    def f(x, *rest):
        if isinstance(x, FArr):
            out = np.empty_like(x.a)
            flat_in, flat_out = x.a.ravel(order="K"), out.ravel(order="K")
            for i in range(flat_in.size):
                flat_out[i] = fn(flat_in[i] if flat_in.dtype == np.float32 else float(flat_in[i]), *rest)
            return FArr(out.reshape(x.a.shape), x.lb)
        return fn(x, *rest)

    return f
