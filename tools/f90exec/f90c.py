"""f90c -- ISO_C_BINDING for the f90py translator (TEST INFRASTRUCTURE).

f90py executes Fortran source by translation to Python.  This module gives the translated code what a Fortran processor
gives a program that says `use, intrinsic :: iso_c_binding`: the kind constants, `c_ptr` / `c_funptr` values, `c_loc`,
`c_funloc`, `c_f_pointer`, `c_associated`, interoperable derived types, and -- the point of it -- procedure interfaces
with `bind(c, name=...)`: every interface body of the source becomes a callable that marshals its actual arguments the
way the Fortran standard's interoperability rules (F2018 18.3) say a companion C processor receives them, and calls the
symbol of that name in a shared library through ctypes.  The prototypes come from the Fortran interface bodies ALONE
(never from the C header), so executing fortran/hrweno_b200_shim.f90 through this module exercises exactly what a
compiler would have bound: a `value` attribute that is missing, an argument in the wrong position or a wrong kind reaches
the C library as the wrong bytes here as it would there.

Marshalling rules (F2018 18.3.6):
  * scalar dummy with VALUE                   -> passed by value (c_int, c_int64_t, c_double, c_float, c_ptr, c_funptr)
  * scalar dummy without VALUE                -> pointer to the scalar; intent(out|inout) results are stored back
  * assumed-size array dummy  x(*)            -> pointer to the first element; a non-contiguous actual is copied in and,
                                                 unless intent(in), copied out (what a compiler does for such a dummy)
  * type(t) with bind(c), without VALUE       -> pointer to a struct laid out from the type's component declarations
  * function result                            -> the C function's return value
A `c_ptr` is a Python int (0 = c_null_ptr).  `c_loc(array)` is the address of the array's first element (the array stays
referenced); `c_loc(object)` is a key under which the object is kept, `c_f_pointer` with that key gives the object back.
`c_funloc(proc)` of a translated `bind(c)` procedure is a ctypes callback built from that procedure's own dummy
declarations; an exception raised inside a callback (e.g. `error stop` in a user's integrand) is re-raised when the C
call that triggered it returns.
"""
from __future__ import annotations

import ctypes as C
import re

import numpy as np

KINDS = {"c_int": 4, "c_int32_t": 4, "c_int64_t": 8, "c_long_long": 8, "c_size_t": 8, "c_double": 8, "c_float": 4, "c_char": 1, "c_bool": 1}

_INT = {"c_int": C.c_int, "c_int32_t": C.c_int32, "c_int64_t": C.c_int64, "c_long_long": C.c_longlong, "c_size_t": C.c_size_t}
_REAL = {"c_double": C.c_double, "c_float": C.c_float}
_NP = {C.c_double: np.float64, C.c_float: np.float32, C.c_int: np.int32, C.c_int32: np.int32, C.c_int64: np.int64, C.c_longlong: np.int64}

HUGE = 1 << 40  # extent given to an assumed-size dummy inside a callback (never touched beyond what the callee indexes)


class CInteropError(TypeError):
    """an actual argument that a Fortran compiler would have rejected (kind / type / rank mismatch)"""


def ctype_of(spec, ns=None):
    """`integer(c_int)`, `real(c_double)`, `type(c_ptr)`, `type(c_funptr)`, `character(kind=c_char)` -> ctypes scalar type.
    A kind given by a named constant of the program (`real(crk)` with `integer, parameter :: crk = c_float`) is looked up
    in `ns`, the program's namespace."""
    s = re.sub(r"\s", "", spec.lower())
    m = re.match(r"^(integer|real|type|character|logical)\((?:kind=)?(\w+)\)$", s)
    if not m:
        raise NotImplementedError(f"not an interoperable type specifier: {spec!r}")
    base, kind = m.groups()
    if base in ("integer", "real") and kind not in _INT and kind not in _REAL and ns is not None and isinstance(ns.get(kind), int):
        value = ns[kind]
        table = {("real", 4): C.c_float, ("real", 8): C.c_double, ("integer", 4): C.c_int, ("integer", 8): C.c_int64}
        if (base, value) in table:
            return table[(base, value)]
    if base == "integer" and kind in _INT:
        return _INT[kind]
    if base == "real" and kind in _REAL:
        return _REAL[kind]
    if base == "type" and kind in ("c_ptr", "c_funptr"):
        return C.c_void_p
    if base == "character" and kind == "c_char":
        return C.c_char
    raise NotImplementedError(f"not an interoperable type specifier: {spec!r}")


class AttrRef:
    """a component `obj%comp` as an actual argument of a by-reference dummy"""

    __slots__ = ("obj", "name")

    def __init__(self, obj, name):
        self.obj, self.name = obj, name

    @property
    def v(self):
        return getattr(self.obj, self.name)

    @v.setter
    def v(self, value):
        setattr(self.obj, self.name, value)


class GRef:
    """a main-program (host-associated) scalar variable as an actual argument of a by-reference dummy"""

    __slots__ = ("ns", "name")

    def __init__(self, ns, name):
        self.ns, self.name = ns, name

    @property
    def v(self):
        return self.ns[self.name]

    @v.setter
    def v(self, value):
        self.ns[self.name] = value


class CCharArr:
    """`character(kind=c_char), pointer :: p(:)` after c_f_pointer: the bytes at an address, read on demand"""

    def __init__(self, addr, n):
        self.addr, self.n = addr, n

    def __getitem__(self, key):
        from f90py import FS

        if isinstance(key, FS):
            lo = 1 if key.lo is None else int(key.lo)
            hi = self.n if key.hi is None else int(key.hi)
            return [self[i] for i in range(lo, hi + 1)]
        i = int(key)
        if not 1 <= i <= self.n:
            raise IndexError(f"subscript {i} outside bounds 1:{self.n}")
        return C.string_at(self.addr + i - 1, 1).decode("latin-1")


class Interop:
    """per-Program state: the library, kept-alive targets, callbacks, pending callback exceptions"""

    def __init__(self, program):
        self.program = program
        self.lib = None
        self.objects = {}    # c_loc(object) key -> object
        self.keep = {}       # address -> array (c_loc targets stay referenced)
        self.callbacks = {}  # id(function) -> (ctypes callback, address)
        self.pending = None  # exception raised inside a callback
        self.structs = {}    # type name -> ctypes.Structure subclass
        self.calls = []      # names of the C symbols called, in order (tests look at it)

    # -- iso_c_binding procedures ---------------------------------------------------------------------------------------
    def c_loc(self, x):
        from f90py import FArr

        if isinstance(x, FArr):
            a = x.a
            if a.size and not (a.flags["F_CONTIGUOUS"] or a.flags["C_CONTIGUOUS"]):
                raise CInteropError("c_loc of a non-contiguous array")
            addr = a.ctypes.data
            self.keep[addr] = a
            return addr
        if x is None:
            raise CInteropError("c_loc of a disassociated pointer / unallocated variable")
        key = id(x)
        self.objects[key] = x
        return key

    def c_funloc(self, f):
        if isinstance(f, int):
            return f
        if f is None:
            return 0
        if id(f) in self.callbacks:
            return self.callbacks[id(f)][1]
        unit = getattr(f, "_f90unit", None)
        if unit is None or not unit.get("bindc"):
            raise CInteropError("c_funloc of a procedure that is not bind(c)")
        cb = self._callback(f, unit)
        addr = C.cast(cb, C.c_void_p).value
        self.callbacks[id(f)] = (cb, addr, f)
        return addr

    def c_f_pointer(self, cptr, what, shape=None):
        """what: 'obj' | 'char' | 'real' | 'integer' (the declared type of the Fortran pointer)"""
        from f90py import FArr

        cptr = int(cptr or 0)
        if cptr == 0:
            raise CInteropError("c_f_pointer of c_null_ptr")
        if what == "obj":
            if cptr not in self.objects:
                raise CInteropError("c_f_pointer: this address was not produced by c_loc of an object")
            return self.objects[cptr]
        dims = [int(v) for v in (shape.a.ravel().tolist() if isinstance(shape, FArr) else (shape or []))]
        if what == "char":
            return CCharArr(cptr, dims[0])
        ct = C.c_double if what == "real" else C.c_int
        n = int(np.prod(dims)) if dims else 1
        arr = np.ctypeslib.as_array(C.cast(cptr, C.POINTER(ct)), shape=(n,))
        return FArr(arr.reshape(dims, order="F")) if dims else arr[0].item()

    @staticmethod
    def c_associated(p, q=None):
        return int(p or 0) != 0 and (q is None or int(p) == int(q or 0))

    @staticmethod
    def transfer(source, mold):
        if isinstance(mold, str) or mold is None:
            return "".join(source)
        raise NotImplementedError("transfer() other than character array -> character scalar")

    # -- interoperable derived types -----------------------------------------------------------------------------------------
    def struct_type(self, tname):
        if tname in self.structs:
            return self.structs[tname]
        td = self.program.types[tname]
        if not td.get("bindc"):
            raise CInteropError(f"type({tname}) is not bind(c)")
        fields = []
        for cname, _ in td["comps"]:
            spec, dims = td["cspec"][cname]
            ct = ctype_of(spec, self.program.ns)
            if dims is not None:
                ct = ct * int(eval(dims, self.program.ns))
            fields.append((cname, ct))
        cls = type(tname, (C.Structure,), {"_fields_": fields})
        self.structs[tname] = cls
        return cls

    def to_struct(self, tname, obj):
        from f90py import FArr

        cls = self.struct_type(tname)
        s = cls()
        for cname, ct in cls._fields_:
            v = getattr(obj, cname)
            if isinstance(v, FArr):
                vals = v.a.ravel(order="F").tolist()
                base = ct._type_
                setattr(s, cname, ct(*[self._scalar(base, x) for x in vals]))
            else:
                setattr(s, cname, self._scalar(ct, v))
        return s

    @staticmethod
    def _scalar(ct, x):
        if hasattr(x, "v") and not isinstance(x, (int, float)):
            x = x.v
        if ct is C.c_void_p:
            return int(x or 0) or None
        if ct in (C.c_double, C.c_float):
            if isinstance(x, (bool, str)) or x is None:
                raise CInteropError(f"{x!r} passed where a real is expected")
            if ct is C.c_float and isinstance(x, float) and np.float32(x) != x:
                raise CInteropError("a binary64 value passed to a real(c_float) dummy")
            if isinstance(x, (int, np.integer)):
                raise CInteropError(f"integer {x!r} passed where a real is expected")
            return float(x)
        if isinstance(x, (float, np.floating)):
            if float(x) != int(x):
                raise CInteropError(f"{x!r} passed where an integer is expected")
            x = int(x)
        if not isinstance(x, (int, np.integer)) or isinstance(x, bool):
            raise CInteropError(f"{x!r} passed where an integer is expected")
        return int(x)

    # -- bind(c) interfaces --------------------------------------------------------------------------------------------------
    def cfunc(self, proto):
        return CFunc(self, proto)

    def _arg_types(self, names, decls):
        """[(name, kind, ctype, intent)] with kind in value | ref | array | struct"""
        out = []
        for a in names:
            if a not in decls:
                raise NotImplementedError(f"dummy argument {a} has no declaration")
            d = decls[a]
            base = d["base"]
            if d["dims"] is not None:
                if d["dims"].strip() != "*":
                    raise NotImplementedError(f"dummy {a}({d['dims']}): only assumed-size arrays are interoperable here")
                out.append((a, "array", ctype_of(base, self.program.ns), d["intent"]))
                continue
            m = re.match(r"^type\s*\(\s*(\w+)\s*\)$", base)
            if m and m.group(1) not in ("c_ptr", "c_funptr"):
                if d["value"]:
                    raise NotImplementedError("struct by value")
                out.append((a, "struct", m.group(1), d["intent"]))
                continue
            out.append((a, "value" if d["value"] else "ref", ctype_of(base, self.program.ns), d["intent"]))
        return out

    def _callback(self, f, unit):
        """ctypes callback for a translated bind(c) procedure: C arguments -> the values the translated body expects"""
        from f90py import FArr

        decls = unit["cdecls"]
        sig = self._arg_types(unit["args"], decls)
        if unit["kind"] == "function":
            res = ctype_of(decls[unit["res"]]["base"], self.program.ns)
        else:
            res = None
        ctypes_args = []
        for _, kind, ct, _ in sig:
            ctypes_args.append(ct if kind == "value" else (C.c_void_p if kind == "struct" else C.POINTER(ct)))
        proto = C.CFUNCTYPE(res, *ctypes_args)

        def thunk(*cargs):
            if self.pending is not None:
                return 0 if res is not None else None
            try:
                pyargs, refs = [], []
                for (name, kind, ct, intent), ca in zip(sig, cargs):
                    if kind == "value":
                        pyargs.append(int(ca or 0) if ct is C.c_void_p else (np.float32(ca) if ct is C.c_float else ca))
                    elif kind == "array":
                        arr = np.ctypeslib.as_array(ca, shape=(HUGE,))
                        pyargs.append(FArr(arr))
                    elif kind == "ref":
                        from f90py import Ref

                        r = Ref(np.float32(ca[0]) if ct is C.c_float else ca[0])
                        refs.append((r, ca))
                        pyargs.append(r)
                    else:
                        raise NotImplementedError("struct argument of a callback")
                out = f(*pyargs)
                for r, ca in refs:
                    ca[0] = r.v
                return out if res is not None else None
            except BaseException as e:  # noqa: BLE001 -- ctypes would print and swallow it
                self.pending = e
                return 0 if res is not None else None

        return proto(thunk)


class CFunc:
    """one `bind(c, name=...)` interface body, callable from translated code"""

    def __init__(self, interop, proto):
        self.io, self.proto = interop, proto
        self.name = proto["name"]
        self.sig = interop._arg_types(proto["args"], proto["decls"])
        self.res = ctype_of(proto["decls"][proto["res"]]["base"], interop.program.ns) if proto["kind"] == "function" else None
        self._fn = None

    def _bind(self):
        if self.io.lib is None:
            raise RuntimeError(f"{self.name}: no library is attached to this program (Program.clib)")
        try:
            fn = getattr(self.io.lib, self.proto["cname"])
        except AttributeError:
            raise RuntimeError(f"the library does not export {self.proto['cname']} (bind(c, name=) of {self.name})") from None
        argtypes = []
        for _, kind, ct, _ in self.sig:
            if kind == "value":
                argtypes.append(ct)
            elif kind == "struct":
                argtypes.append(C.c_void_p)
            else:
                argtypes.append(C.c_void_p)
        fn.argtypes, fn.restype = argtypes, self.res
        self._fn = fn

    def __call__(self, *args, **kw):
        from f90py import FArr, Ref

        if self._fn is None:
            self._bind()
        names = [s[0] for s in self.sig]
        actual = dict(zip(names, args))
        for k, v in kw.items():
            if k not in names or k in actual:
                raise CInteropError(f"{self.name}: bad keyword argument {k}")
            actual[k] = v
        if len(args) > len(names) or set(actual) != set(names):
            raise CInteropError(f"{self.name}: expected arguments {names}, got {sorted(actual)}")
        cargs, after, keep = [], [], []
        for name, kind, ct, intent in self.sig:
            x = actual[name]
            if kind == "value":
                cargs.append(self.io._scalar(ct, x))
            elif kind == "ref":
                holder = ct()
                if intent != "out":
                    v = self.io._scalar(ct, x.v if isinstance(x, (Ref, AttrRef, GRef)) else x)
                    holder = ct(v) if v is not None else ct()
                if intent in ("out", "inout"):
                    if not isinstance(x, (Ref, AttrRef, GRef)):
                        raise CInteropError(f"{self.name}: argument {name} is intent({intent}) but the actual is not definable")
                    after.append((x, holder, ct))
                keep.append(holder)
                cargs.append(C.addressof(holder))
            elif kind == "array":
                if isinstance(x, FArr):
                    a = x.a
                elif isinstance(x, np.ndarray):
                    a = x
                else:
                    raise CInteropError(f"{self.name}: argument {name} must be an array, got {type(x).__name__}")
                want = _NP.get(ct)
                if want is None or a.dtype != want:
                    raise CInteropError(f"{self.name}: argument {name} is {a.dtype}, the dummy is {ct.__name__}")
                if a.size and not (a.flags["F_CONTIGUOUS"] or (a.ndim == 1 and a.flags["C_CONTIGUOUS"])):
                    tmp = np.array(a, order="F", copy=True)  # copy-in
                    if intent != "in":
                        after.append((a, tmp, None))          # copy-out
                    a = tmp
                if intent != "in" and not a.flags["WRITEABLE"]:
                    raise CInteropError(f"{self.name}: argument {name} is not definable")
                keep.append(a)
                cargs.append(a.ctypes.data)
            else:  # struct by reference
                s = self.io.to_struct(ct, x)
                keep.append(s)
                cargs.append(C.addressof(s))
        self.io.calls.append(self.proto["cname"])
        out = self._fn(*cargs)
        if self.io.pending is not None:
            e, self.io.pending = self.io.pending, None
            raise e
        for target, holder, ct in after:
            if ct is None:
                target[...] = holder
            else:
                v = holder.value
                target.v = int(v or 0) if ct is C.c_void_p else (np.float32(v) if ct is C.c_float else v)
        if self.res is C.c_void_p:
            return int(out or 0)
        return np.float32(out) if self.res is C.c_float else out


def install(program):
    """names of iso_c_binding in the program's namespace; returns the Interop state"""
    io = Interop(program)
    ns = program.ns
    ns.update(KINDS)
    ns.update(
        c_null_ptr=0, c_null_funptr=0, c_null_char="\0", c_loc=io.c_loc, c_funloc=io.c_funloc, c_associated=io.c_associated,
        c_f_pointer_=io.c_f_pointer, transfer=io.transfer, new_c_ptr=lambda: 0, new_c_funptr=lambda: 0, AttrRef=AttrRef, GRef=GRef,
    )
    return io
