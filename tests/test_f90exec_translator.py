"""CPU-only: the Fortran-subset translator (tools/f90exec/f90py.py) on small snippets whose results are known by the
language rules, so that the fixtures it produces from the reference's source (tests/golden/ref_exec_*.npz) can be trusted.
The snippets are written here; none comes from the reference."""
import os
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools", "f90exec"))
import f90py  # noqa: E402
from f90py import FArr, Ref, callm  # noqa: E402


def build(tmp_path, src):
    p = tmp_path / "snippet.f90"
    p.write_text(textwrap.dedent(src))
    return f90py.Program().add_source(str(p)).build()


def test_expression_rules(tmp_path):
    ns = build(tmp_path, """
        module m
        contains
           pure real(rk) function f(a, b, c)
              real(rk), intent(in) :: a, b, c
              f = a - b*c**2/3 + (-a**2)    ! ** before * /, left to right, unary minus below **
           end function
           pure integer function idiv(i, j)
              integer, intent(in) :: i, j
              idiv = i/j + (-i)/j            ! integer division truncates toward zero
           end function
           pure real(rk) function mixed(x, n)
              real(rk), intent(in) :: x
              integer, intent(in) :: n
              mixed = 13.0_rk/12*x + n/2 + x**(-2) + 2**n   ! real/int is real; int/int is int; real**int by squaring
           end function
           pure logical function rel(a, b)
              real(rk), intent(in) :: a, b
              rel = .not. (a >= b) .and. a /= b .or. a == b
           end function
        end module
    """)
    a, b, c = 0.7, 1.3, -2.9
    assert ns["f"](a, b, c) == a - ((b * (c * c)) / 3) + (-(a * a))
    assert ns["idiv"](7, 2) == 3 + (-3) and ns["idiv"](-7, 2) == -3 + 3
    x = 1.1
    assert ns["mixed"](x, 5) == ((((13.0 / 12) * x) + 2) + 1.0 / (x * x)) + 32
    assert ns["rel"](1.0, 2.0) is True and ns["rel"](2.0, 1.0) is False and ns["rel"](1.0, 1.0) is True


def test_sum_is_sequential_from_zero_and_sections_are_inclusive(tmp_path):
    ns = build(tmp_path, """
        module m
        contains
           subroutine s(v, out)
              real(rk), intent(in) :: v(:)
              real(rk), intent(out) :: out(:)
              real(rk) :: w(-2:3)
              w = 0.0_rk
              w(-2:0) = v(1:3)              ! lower bound -2
              w(1:) = v(6:4:-1)             ! negative stride
              out(1) = sum(w)
              out(2) = sum(w(-2:3:2)*v(1:3))
              out(3) = w(-2) + w(3)
              out(4) = size(w) + lbound(w, 1) + ubound(w, 1)
           end subroutine
        end module
    """)
    v = np.array([1e16, 1.0, -1e16, 3.0, 5.0, 7.0])
    out = np.zeros(4)
    ns["s"](v, out)
    w = [1e16, 1.0, -1e16, 7.0, 5.0, 3.0]
    acc = 0.0
    for e in w:
        acc = acc + e
    assert out[0] == acc == 15.0  # (1e16 + 1) rounds to 1e16: a pairwise or reversed sum would give 16
    assert out[1] == ((0.0 + w[0] * v[0]) + w[2] * v[1]) + w[4] * v[2]
    assert out[2] == 1e16 + 3.0 and out[3] == 6 - 2 + 3


def test_types_methods_optional_and_reference_arguments(tmp_path):
    ns = build(tmp_path, """
        module m
           type :: base
              integer :: n = 4
              real(rk) :: scale = 2.0_rk
              real(rk), allocatable :: a(:)
           contains
              procedure, pass(self) :: fill => base_fill
           end type
           type, extends(base) :: child
              real(rk), allocatable :: h(:, :)
           contains
              procedure, pass(self) :: step => child_step
           end type
           interface child
              module procedure :: child_init
           end interface
        contains
           type(child) function child_init(n, scale) result(self)
              integer, intent(in) :: n
              real(rk), intent(in), optional :: scale
              self%n = n
              if (present(scale)) self%scale = scale
              allocate (self%a(0:n - 1), self%h(n, 2))
              call self%fill(1.0_rk)
           end function
           subroutine base_fill(self, x0)
              class(base), intent(inout) :: self
              real(rk), intent(in) :: x0
              integer :: i
              do concurrent(i=0:self%n - 1)
                 self%a(i) = x0 + self%scale*i
              end do
           end subroutine
           subroutine child_step(self, t, dt, mode)
              class(child), intent(inout) :: self
              real(rk), intent(inout) :: t
              real(rk), intent(in) :: dt
              integer, intent(in), optional :: mode
              integer :: i
              do
                 t = t + dt
                 self%h = eoshift(self%h, shift=-1, dim=2)
                 self%h(:, 1) = self%a
                 select case (optval(mode, 1))
                 case (1)
                    self%a = self%a/2
                 case default
                    self%a = -self%a
                 end select
                 if (t > 1.0_rk) exit
              end do
           end subroutine
        end module
    """)
    c = ns["child"](3, scale=0.5)
    assert c.a.lb == (0,) and np.array_equal(c.a.a, [1.0, 1.5, 2.0]) and c.h.a.shape == (3, 2)
    t = Ref(0.0)
    callm(c, "step", t, 0.4)
    assert t.v == 0.4 + 0.4 + 0.4 and np.array_equal(c.a.a, np.array([1.0, 1.5, 2.0]) / 8)
    assert np.array_equal(c.h.a[:, 0], np.array([1.0, 1.5, 2.0]) / 4) and np.array_equal(c.h.a[:, 1], np.array([1.0, 1.5, 2.0]) / 2)
    callm(c, "step", t, 0.4, mode=2)
    assert np.array_equal(c.a.a, -np.array([1.0, 1.5, 2.0]) / 8)
    d = ns["child"](2)
    assert d.scale == 2.0 and np.array_equal(d.a.a, [1.0, 3.0])


def test_pointer_remapping_cycle_and_multi_index_concurrent(tmp_path):
    ns = build(tmp_path, """
        module m
        contains
           subroutine s(x, out)
              real(rk), intent(in) :: x(0:)
              real(rk), intent(out) :: out(:, :)
              real(rk), target :: ext(-1:size(x))
              real(rk), dimension(:), pointer :: xl
              integer :: i, j, l
              real(rk) :: p
              ext(0:size(x) - 1) = x
              ext(-1) = 2*ext(0) - ext(1)
              ext(size(x)) = 2*ext(size(x) - 1) - ext(size(x) - 2)
              xl(0:) => ext(lbound(ext, 1):ubound(ext, 1) - 1)    ! xl(i) = ext(i-1)
              do concurrent(i=1:2, j=1:3)
                 p = 1.0_rk
                 do l = 0, 3
                    if (l == j) cycle
                    p = p*(xl(l) - xl(j))
                 end do
                 out(i, j) = p*i
              end do
           end subroutine
        end module
    """)
    x = np.array([0.0, 1.0, 3.0])
    out = np.zeros((2, 3), order="F")
    ns["s"](FArr(x, (0,)), out)
    xl = [-1.0, 0.0, 1.0, 3.0]
    for i in (1, 2):
        for j in (1, 2, 3):
            p = 1.0
            for l in range(4):
                if l != j:
                    p = p * (xl[l] - xl[j])
            assert out[i - 1, j - 1] == p * i


def test_unsupported_statements_fail_loudly(tmp_path):
    with pytest.raises(NotImplementedError):
        build(tmp_path, """
            module m
            contains
               subroutine s(x)
                  real(rk), intent(inout) :: x
                  write (*, *) x
               end subroutine
            end module
        """)


def test_fortran_shim_parses_as_far_as_the_translator_can_tell():
    """fortran/hrweno_b200_shim.f90 cannot be compiled in this image (no Fortran compiler).  A weak substitute, not a
    compile: every block construct of the file is balanced, and every procedure body is inside the subset the translator
    accepts (statement and expression syntax).  The strong substitute is tests/test_fortran_shim_exec.py, which EXECUTES it."""
    import re

    path = os.path.join(ROOT, "fortran", "hrweno_b200_shim.f90")
    lines = f90py.logical_lines(open(path).read())
    openers = [(r"^module\s+\w+$", "module"), (r"^(abstract\s+)?interface\b", "interface"), (r"^type\s*(,[^:]*)?(::)?\s*\w+$", "type"),
               (r"^(\w[\w\s\(\),=\*:]*\s)?(subroutine|function)\s+\w+", "proc"), (r"^if\s*\(.*\)\s*then$", "if"), (r"^do\b", "do"),
               (r"^select\s+case", "select"), (r"^associate\b", "associate")]
    stack = []
    for no, ln in lines:
        low = ln.lower()
        m = re.match(r"^end\s*(module|interface|type|subroutine|function|if|do|select|associate)?", low)
        if low.startswith("end") and (low == "end" or m.group(1)):
            kind = {"subroutine": "proc", "function": "proc"}.get(m.group(1), m.group(1))
            assert stack, f"line {no}: `{ln}` closes nothing"
            top = stack.pop()
            assert kind is None or top[0] == kind, f"line {no}: `{ln}` closes {top}"
            continue
        for rx, kind in openers:
            if re.match(rx, low) and not low.startswith(("module procedure", "procedure")) and not (kind == "type" and low.startswith("type(")):
                stack.append((kind, no, ln[:60]))
                break
    assert not stack, f"unclosed constructs: {stack}"
    P = f90py.Program(skip_io=True)
    P.add_source(path)
    assert {"weno", "rktvd", "mstvd"} <= set(P.generics) and {"weno", "tvdode", "rktvd", "mstvd", "hrweno_fv_desc"} <= set(P.types)
    failed = []
    for name, unit in P.procs.items():
        f90py.Program._do_stack = []
        try:
            compile(P.gen_unit(unit), "<shim>", "exec")
        except NotImplementedError as e:
            failed.append((name, str(e)))
    assert failed == []
    assert {"hrweno_weno_create", "hrweno_ode_integrate", "hrweno_fv_create", "hrweno_mgpu_create"} <= set(P.cprotos)


def test_main_program_with_host_association_and_object_arrays(tmp_path):
    ns = build(tmp_path, """
        module m
           type :: box
              real(rk) :: w = 1.5_rk
           contains
              procedure, pass(self) :: grow => box_grow
           end type
        contains
           subroutine box_grow(self, f)
              class(box), intent(inout) :: self
              real(rk), intent(in) :: f
              self%w = self%w*f
           end subroutine
        end module
        program p
           use m
           implicit none
           integer, parameter :: n(2) = [3, 2], k = 2
           real(rk) :: u(product(n)), t, dt
           type(box) :: b(2)
           integer :: i
           call b(2)%grow(f=2.0_rk)
           t = 0.0_rk; dt = 0.25_rk
           do i = 1, size(u)
              u(i) = i*b(2)%w
           end do
           do i = 0, 3
              call advance(t, dt)
              call output(2)
           end do
        contains
           subroutine advance(tt, h)
              real(rk), intent(inout) :: tt
              real(rk), intent(in) :: h
              tt = tt + h*k
              u = u + n(1)      ! host-associated array, parameter array and named constant
           end subroutine
        end program
    """)
    seen = []
    ns["output"] = lambda action: seen.append((ns["t"], ns["u"].a.copy()))
    ns["main_p"]()
    assert [s[0] for s in seen] == [0.5, 1.0, 1.5, 2.0]
    assert np.array_equal(seen[-1][1], np.arange(1, 7) * 3.0 + 12.0)
    assert ns["b"][1].w == 1.5 and ns["b"][2].w == 3.0


# ---- rk = real32 (the reference's -DREAL32 build) ---------------------------------------------------------------------
F = np.float32


def test_real32_every_operation_rounds_to_binary32(tmp_path):
    """literals (with or without _rk), mixed real/integer arithmetic, integer powers, `sum()` and array expressions are
    binary32 operations; the expected values are written operation by operation with np.float32"""
    with f90py.real_kind(4):
        ns = build(tmp_path, """
            module m
            contains
               pure real(rk) function f(a, b, n)
                  real(rk), intent(in) :: a, b
                  integer, intent(in) :: n
                  f = 13.0_rk/12*a + 0.1*b - a**3/n + 1e-6_rk
               end function
               subroutine s(v, out)
                  real(rk), intent(in) :: v(:)
                  real(rk), intent(out) :: out(:)
                  real(rk) :: w(0:2), acc
                  integer :: i
                  w = v(1:3)*2.5_rk + v(4:6)/3
                  out(1) = sum(w)
                  acc = 0
                  do i = 0, 2
                     acc = acc + w(i)**2
                  end do
                  out(2) = acc
                  out(3) = sqrt(abs(w(0))) + max(w(1), w(2), 0.25_rk) + sign(1.0_rk, -w(0))
                  out(4) = real(size(w), rk)/7 + epsilon(1.0_rk)
               end subroutine
            end module
        """)
        a, b = F(0.7), F(1.3)
        r = ns["f"](a, b, 3)
        assert isinstance(r, np.float32)
        assert r == ((F(13.0) / F(12)) * a + F(0.1) * b) - ((a * a) * a) / F(3) + F(1e-6)
        assert float(r) != (13.0 / 12) * 0.7 + 0.1 * 1.3 - 0.7**3 / 3 + 1e-6  # not the binary64 value rounded late
        v, out = np.array([0.3, -1.7, 2.9, 0.11, 5.3, -0.77], dtype=F), np.zeros(4, dtype=F)
        ns["s"](v, out)
        w = v[:3] * F(2.5) + v[3:] / F(3)
        assert w.dtype == F and out.dtype == F
        assert out[0] == ((F(0) + w[0]) + w[1]) + w[2]
        assert out[1] == ((F(0) + w[0] * w[0]) + w[1] * w[1]) + w[2] * w[2]
        assert out[2] == (np.sqrt(np.abs(w[0])) + max(w[1], w[2], F(0.25))) + np.copysign(F(1), -w[0])
        assert out[3] == F(3) / F(7) + np.finfo(F).eps
    assert f90py.rkind() == 8  # the context restores binary64


def test_real32_a_binary64_value_cannot_slip_in(tmp_path):
    with f90py.real_kind(4):
        ns = build(tmp_path, """
            module m
            contains
               subroutine s(x, v)
                  real(rk), intent(in) :: x
                  real(rk), intent(inout) :: v(:)
                  v(1) = x
               end subroutine
            end module
        """)
        v = np.zeros(2, dtype=F)
        ns["s"](F(0.1), v)
        assert v[0] == F(0.1)
        with pytest.raises(TypeError, match="binary64"):
            ns["s"](0.1, v)  # a Python float (binary64) reaching a store
        with pytest.raises(TypeError, match="binary64"):
            ns["s"](F(0.1), np.zeros(2))  # a float64 array as an actual argument
        with pytest.raises(TypeError, match="binary64"):
            FArr(np.zeros(3))


def test_preprocessor_subset_and_module_variables(tmp_path):
    """`gfortran -cpp -DNAME` as src/hrweno_kinds.F90 uses it (#ifdef / #elif / #else / #endif, object-like #define), and
    module variables with an initial value that the module's procedures define"""
    src = """
        module counters
           implicit none
        #ifdef WIDE
           integer, parameter :: width = 8
        #define LABEL "wide"
        #elif NARROW
           integer, parameter :: width = 4
        #define LABEL "narrow"
        #else
           integer, parameter :: width = 0
        #define LABEL "none"
        #endif
           integer :: ncalls = 0
           real(rk) :: seen(3) = -1.0_rk
        contains
           subroutine touch(x)
              real(rk), intent(in) :: x
              ncalls = ncalls + 1
              if (ncalls <= 3) seen(ncalls) = x*width
           end subroutine
           function label() result(s)
              character(:), allocatable :: s
              s = LABEL // "LABEL"
           end function
        #ifndef WIDE
           integer function only_without_wide()
              only_without_wide = 1
           end function
        #endif
        end module
    """
    p = tmp_path / "c.f90"
    p.write_text(textwrap.dedent(src))
    for defines, width, label in (({"WIDE": "1"}, 8, "wide"), ({"NARROW": "1"}, 4, "narrow"), ({}, 0, "none")):
        ns = f90py.Program(defines=defines).add_source(str(p)).build()
        assert ns["width"] == width and ns["label"]() == label + "LABEL"
        assert ("only_without_wide" in ns) == ("WIDE" not in defines)
        for x in (1.5, 2.5):
            ns["touch"](x)
        assert ns["ncalls"] == 2 and list(ns["seen"].a) == [1.5 * width, 2.5 * width, -1.0]
    with pytest.raises(NotImplementedError):
        f90py.preprocess("#if A && B\nx\n#endif\n", {"A": "1"})
    with pytest.raises(NotImplementedError):
        f90py.preprocess("#ifdef A\nx\n", {})
