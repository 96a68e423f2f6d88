"""Semantics of oracle/f90interp.py on small Fortran texts with known answers: what the reference's hot-path files rely on
(the interpreter runs THEIR text to make tests/golden/reference_interp.json.gz; here it is checked on its own)."""
import math

import pytest

from oracle.f90interp import Cell, FArray, FortranError, Interpreter, logical_lines, tokenize

SRC = """
module consts
   implicit none
   integer, parameter :: n = 4, m = n*2
   real,    parameter :: PI = 3.141592, third = 1./3
   real :: shared
   real, allocatable :: grid(:,:)
end module consts

module work
   implicit none
contains
   integer function idiv(a, b)
      integer, intent(IN) :: a, b
      idiv = a / b
   end function idiv

   real function mixed(i, x)
      ! left to right, integer operand converted: ((i-1)*2.)*x/3  -- and integer division inside parentheses
      integer :: i
      real    :: x
      mixed = (i-1) * 2. * x/3 + (7/2)
   end function mixed

   subroutine bump(a, k, flag)
      use consts, only : shared
      real,    intent(INOUT) :: a
      integer, intent(INOUT) :: k
      logical, intent(INOUT) :: flag
      a = a + 1.5
      k = k + 1
      flag = .not. flag
      shared = shared + a
   end subroutine bump

   subroutine fill(v, total)
      real, intent(INOUT) :: v(:)
      real, intent(OUT)   :: total
      integer :: i
      total = 0.
      do i = size(v), 1, -1
         v(i) = i * 1.d-1
         if (i == 2) cycle
         total = total + v(i)
      end do
   end subroutine fill

   integer function collatz(start)
      integer, intent(IN) :: start
      integer :: x
      x = start
      collatz = 0
      do
         if (x == 1) exit
         if (mod(x, 2) == 0) then
            x = x / 2
         else if (x > 0) then
            x = 3*x + 1
         else
            error stop 1
         end if
         collatz = collatz + 1
      end do
   end function collatz

   subroutine jumps(x, path)
      real,    intent(IN)  :: x
      integer, intent(OUT) :: path
      path = 0
      if (x.gt.1.) then
         path = 1
         goto 100
      else
         if (x.eq.-1.) then
            goto 100
         end if
      end if
      path = path + 10
      100 continue
      path = path + 100
   end subroutine jumps

   subroutine flags(d)
      logical, intent(INOUT) :: d(:)
      d = (/.FALSE., .TRUE., .FALSE./)
      if(.not.d(1) .and. d(2) .or. d(3)) d(3) = .true.
   end subroutine flags

   real function powers(x)
      real :: x
      powers = x**2 + x**2. - 2.*x**2
   end function powers

   subroutine elem(a, i)
      use consts
      integer :: i
      real :: a(:)
      call bump1(a(i))
      grid(0, i) = a(i) * PI
   end subroutine elem

   subroutine bump1(z)
      real, intent(INOUT) :: z
      z = z * 2.
   end subroutine bump1

   subroutine undefined_use(y)
      real :: y, q
      y = q + 1.
   end subroutine undefined_use
end module work
"""

FIXED = """\
      FUNCTION counter(reset)
      INTEGER reset,NT
      REAL counter,SCALE
      PARAMETER (NT=3,
     *           SCALE=0.5)
      INTEGER calls,iv(NT),j
      SAVE calls,iv
      DATA calls/0/, iv/NT*7/
C     a comment line
      if (reset.ne.0) calls=0
      calls=calls+1
      do 11 j=NT,1,-1
        iv(j)=iv(j)+j
11    continue
      counter=calls*10+iv(1)*SCALE
      return
      END
"""


@pytest.fixture()
def it():
    x = Interpreter()
    x.load_text(SRC)
    return x


def test_parameters_and_module_variables(it):
    assert it.var("consts", "n").v == 4 and it.var("consts", "m").v == 8
    assert it.var("consts", "pi").v == 3.141592 and it.var("consts", "third").v == 1.0 / 3   # literal -> binary64 directly
    assert it.modules["consts"].vars["grid"] is None and it.modules["consts"].alloc_types["grid"] == ("r", 2)


def test_integer_division_truncates_toward_zero(it):
    for a, b, q in ((7, 2, 3), (-7, 2, -3), (7, -2, -3), (-7, -2, 3), (1, 53668, 0)):
        assert it.call("idiv", [Cell("i", a), Cell("i", b)], want_result=True) == q


def test_mixed_mode_and_left_to_right(it):
    got = it.call("mixed", [Cell("i", 4), Cell("r", 0.7)], want_result=True)
    assert got == ((3 * 2.0) * 0.7) / 3 + 3 and isinstance(got, float)


def test_arguments_are_passed_by_reference(it):
    a, k, flag = Cell("r", 1.0), Cell("i", 5), Cell("l", False)
    it.var("consts", "shared").set(10.0)
    it.call("bump", [a, k, flag])
    assert (a.v, k.v, flag.v) == (2.5, 6, True) and it.var("consts", "shared").v == 12.5


def test_counted_loop_with_negative_step_cycle_and_assumed_shape(it):
    v, total = FArray("r", (5,)), Cell("r")
    it.call("fill", [v, total])
    assert list(v.a) == [1 * 1e-1, 2 * 1e-1, 3 * 1e-1, 4 * 1e-1, 5 * 1e-1]
    assert total.v == ((0.0 + 5 * 1e-1) + 4 * 1e-1 + 3 * 1e-1) + 1 * 1e-1          # accumulation order: i = 5, 4, 3, 1


def test_endless_loop_exit_else_if_and_function_result(it):
    assert it.call("collatz", [Cell("i", 27)], want_result=True) == 111
    with pytest.raises(FortranError, match="ERROR STOP"):
        it.call("collatz", [Cell("i", -5)], want_result=True)


def test_goto_to_a_top_level_label(it):
    for x, want in ((2.0, 101), (-1.0, 100), (0.5, 110)):
        p = Cell("i")
        it.call("jumps", [Cell("r", x), p])
        assert p.v == want


def test_logical_arrays_constructors_and_precedence(it):
    d = FArray("l", (3,))
    it.call("flags", [d])
    assert list(d.a) == [False, True, True]                  # (.not. d1 .and. d2) .or. d3


def test_integer_power_is_a_product_not_pow(it):
    """x**2 is x*x (gfortran's powi expansion) and pow(x, 2.) is folded to x*x too: the three terms cancel exactly for every x,
    which libm's pow -- accurate to under an ulp, not correctly rounded -- does not guarantee."""
    for x in (0.9, 1.0 / 3, math.pi, 0.8264510225504637, 123456.789):
        assert it.call("powers", [Cell("r", x)], want_result=True) == 0.0


def test_array_element_as_actual_argument_and_lower_bounds(it):
    g = it.allocate("consts", "grid", (2, 6), (0, 0))
    a = FArray("r", (3,))
    a.fill([1.0, 2.0, 3.0])
    it.call("elem", [a, Cell("i", 2)])
    assert list(a.a) == [1.0, 4.0, 3.0] and g.get((0, 2)) == 4.0 * 3.141592
    with pytest.raises(FortranError, match="subscript"):
        g.get((2, 0))


def test_use_of_an_undefined_local_is_an_error(it):
    with pytest.raises(FortranError, match="before it was given a value"):
        it.call("undefined_use", [Cell("r", 0.0)])


def test_fixed_form_parameter_save_data_labelled_do():
    x = Interpreter()
    x.load_text(FIXED, fixed=True)
    r = Cell("i", 0)
    assert x.call("counter", [r], want_result=True) == 1 * 10 + (7 + 1) * 0.5      # iv = 7,7,7 -> 8,9,10
    assert x.call("counter", [r], want_result=True) == 2 * 10 + (8 + 1) * 0.5      # SAVE: calls and iv persist
    r.set(1)
    assert x.call("counter", [r], want_result=True) == 1 * 10 + (9 + 1) * 0.5


def test_tokens_that_look_alike():
    kinds = lambda s: [v for _, v in tokenize(s)]
    assert kinds("if(abs(bmu).gt.1.) then") == ["if", "(", "abs", "(", "bmu", ")", ".gt.", "1.", ")", "then"]
    assert kinds("x.eq.-1.") == ["x", ".eq.", "-", "1."]
    assert kinds("1.e-8*(2.*.5/n)") == ["1.e-8", "*", "(", "2.", "*", ".5", "/", "n", ")"]
    assert kinds("tflag.eqv..false.") == ["tflag", ".eqv.", ".false."]
    assert kinds("d=(/.true.,.false./)") == ["d", "=", "(/", ".true.", ",", ".false.", "/)"]
    assert kinds("250d-4") == ["250d-4"]
    ll = logical_lines("a = 1 &\n  + 2 ! c\n100 continue\nb = 'x!y'")
    assert [(lab, s) for _, lab, s in ll] == [(None, "a = 1 + 2"), ("100", "continue"), (None, "b = 'x!y'")]


SRC2 = """
module props
   implicit none
   real, parameter :: w0 = .75
   real :: scale
   real, allocatable :: field(:,:,:)
   character(len=16) :: mode
   procedure(pa), pointer :: pick => null()
   private
   public :: scale, field
contains
   elemental real function clip(x, hi)
      real, intent(IN) :: x, hi
      clip = max(min(w0 - w0 * (x / scale), hi), 0.0)
   end function clip

   real function pa() result (val)
      val = 2.d0 * scale
   end function pa

   real function pb() result (val)
      if (scale > 1.) then
         val = -1.
         return
      end if
      val = scale
   end function pb

   subroutine shifted(t, np)
      integer, intent(IN)    :: np
      real,    intent(INOUT) :: t(0:np+1, 0:np+1)
      t(0, np+1) = 7.
      t(np+1, 0) = t(np+1, 0) + 1.
   end subroutine shifted
end module props
"""


def test_sections_array_arithmetic_elemental_result_clause_strings_and_dummy_bounds():
    x = Interpreter()
    x.load_text(SRC2)
    assert x.procs["clip"].elemental and x.procs["pa"].result_name == "val" and x.var("props", "mode").t == "c"
    g = x.allocate("props", "field", (4, 4, 3), (0, 0, 1))
    f = {"field": g, "scale": x.var("props", "scale"), "mode": x.var("props", "mode"), "n": Cell("i", 2), "q": FArray("r", (2, 2)),
         "w": FArray("r", (2, 2)), "flag": Cell("l")}
    src = "\n".join([
        "scale = 4.",
        "field = 1.5",                         # whole array
        "field(n+1,:,:) = 2.",                 # a plane
        "field(1:n,1:n,2) = 9.",               # a box of one plane
        "field(:,:,3) = field(:,:,1) * (scale/8.) + 1.",   # section on both sides, array arithmetic
        "q = field(1:n,1:n,2)",
        "w = clip(q, .5)",                     # elemental over an array
        "mode = 'gaussian'",
        "flag = trim(mode) == \"gaussian\"",
    ])
    import tempfile, os
    with tempfile.NamedTemporaryFile("w", suffix=".f90", delete=False) as t:
        t.write(src)
    try:
        x.run_block(t.name, 1, 9, f)
    finally:
        os.unlink(t.name)
    a = g.a                                  # stored 0-based: a[i, j, k-1]
    assert a[3, 0, 0] == 2.0 and a[0, 0, 0] == 1.5 and a[1, 1, 1] == 9.0 and a[0, 1, 1] == 1.5 and a[3, 2, 1] == 2.0
    assert a[3, 1, 2] == 2.0 * 0.5 + 1 and a[0, 0, 2] == 1.5 * 0.5 + 1
    assert (f["q"].a == 9.0).all()
    assert (f["w"].a == max(min(0.75 - 0.75 * (9.0 / 4.0), 0.5), 0.0)).all() and f["flag"].v is True
    # result clause, RETURN with the result set, a procedure pointer bound by the harness
    assert x.call("pa", [], want_result=True) == 8.0 and x.call("pb", [], want_result=True) == -1.0
    x.alias["pick"] = "pb"
    x.var("props", "scale").set(0.25)
    assert x.call("pick", [], want_result=True) == 0.25
    # an explicit-shape dummy with its own lower bounds over the caller's storage
    t2 = FArray("r", (4, 4))                 # bounds 1:4 in the caller
    x.call("shifted", [t2, Cell("i", 2)])
    assert t2.a[0, 3] == 7.0 and t2.a[3, 0] == 1.0
    with pytest.raises(FortranError, match="shape"):
        f["q"].fill(g)
