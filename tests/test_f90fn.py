"""Self-test of the mechanical Fortran -> Python translator (tools/f90fn.py) on a small Fortran text written for this test:
the language rules the reference goldens depend on -- integer division and real->integer truncation, implicit typing,
1-based and custom-lower-bound arrays, array sections passed by reference, sequence association, write-back of modified
scalar dummies (also through a function reference), COMMON storage association across different partitions, DATA with
implied DO and repeat counts, OPTIONAL / keyword arguments, forward GOTO out of and inside loops, labelled DO with a shared
terminator, SELECT CASE, derived types with value semantics."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import f90fn  # noqa: E402
import f90rt as rt  # noqa: E402

SRC = """
module consts
  implicit none
  real :: scale_ref = 2.5
  integer, parameter :: nn = 4, mm = 1.9*nn      ! real expression truncated into an integer parameter
  type pair_t
     real :: a, b
  end type pair_t
contains
  subroutine getc(scale, twice)
    real, intent(out), optional :: scale, twice
    if (present(scale)) scale = scale_ref
    if (present(twice)) twice = 2*scale_ref
  end subroutine getc
end module consts

      subroutine intrules(i1, i2, r, iq, ir, itrunc, x)
      ! implicit typing: i*, j* ... integer; everything else real
      iq = i1/i2                 ! integer division truncates toward zero
      ir = mod(i1,i2)
      itrunc = r                 ! real -> integer assignment truncates toward zero
      x = i1/i2 + r/2            ! mixed: the integer quotient is formed first
      end

      function bump(a, k)
      ! a function that also modifies its argument (like HALFWHM_C does with AS)
      if (a .eq. 0.) a = 5*k
      bump = a + 1
      end function bump

      subroutine usebump(v, res)
      dimension v(3)
      res = bump(v(2), 3) + bump(v(1), 2)
      end

      subroutine sections(a, n, tot)
      real a(0:n, 3)
      call fill(a(0:n, 2), n+1, 7.)
      call seq(a(1, 3), 2)
      tot = sum(a(:, 2)) + a(1, 3) + a(2, 3)
      end
      subroutine fill(v, n, val)
      dimension v(n)
      do 10 i = 1, n
   10 v(i) = val + i
      end
      subroutine seq(v, n)
      dimension v(*)
      do i = 1, n
         v(i) = 100.*i
      end do
      end

      block data bd
      common /blk/ v1, npt, s0(2), s1(3)
      data v1, npt /-20.0, 5/
      data s0 /1., 2./
      data (s1(i), i=1,3) /3., 2*4.5/
      end
      function sumblk(k)
      common /blk/ v1, npt, s(5)
      sumblk = v1*0
      do i = 1, npt
         if (i .gt. k) goto 20
         sumblk = sumblk + s(i)
      enddo
   20 continue
      end

      subroutine gotos(n, cnt, last)
      cnt = 0
      do 30 i = 1, n
      do 30 j = 1, 3
         if (j .eq. 2) goto 30
         cnt = cnt + 1
         if (i .eq. 3 .and. j .eq. 3) goto 40
   30 continue
   40 last = 10*i + j
      end

      subroutine sel(k, r)
      use consts
      type(pair_t) :: p, q
      p = pair_t(1., 2.)
      q = p
      q%a = 9.
      select case (k)
      case (1)
         r = p%a
      case (2:4)
         r = q%a
      case default
         r = -1.
      end select
      call getc(twice=t2)
      r = r + t2 + mm
      end
"""


def _load():
    with tempfile.NamedTemporaryFile("w", suffix=".f90", delete=False) as f:
        f.write(SRC)
        path = f.name
    try:
        return f90fn.load([path])
    finally:
        os.unlink(path)


def test_translator_language_rules():
    ns = _load()
    r = ns["intrules"](-7, 2, -2.7, 0, 0, 0, 0.0)
    assert r[1:] == (-3, -1, -2, -3 + (-2.7) / 2)
    v = rt.FArr(np.array([0.0, 0.0, 1.0]))
    r = ns["usebump"](v, 0.0)
    assert r[1] == (15.0 + 1) + (10.0 + 1) and list(v.a) == [10.0, 15.0, 1.0]      # written back through the function reference
    a = rt.FArr.zeros((4, 3), (0, 1))
    r = ns["sections"](a, 3, 0.0)
    assert list(a.a[:, 1]) == [8.0, 9.0, 10.0, 11.0] and a.a[1, 2] == 100.0 and a.a[2, 2] == 200.0
    assert r[-1] == 38.0 + 300.0
    assert ns["sumblk"](5)[0] == 1 + 2 + 3 + 4.5 + 4.5 and ns["sumblk"](2)[0] == 3.0     # storage association across partitions
    r = ns["gotos"](5, 0.0, 0)
    assert r[1] == 2 * 3 and r[2] == 33       # cycle through the shared terminator, then jump out of both loops
    assert ns["sel"](1, 0.0)[1] == 1.0 + 5.0 + 7 and ns["sel"](3, 0.0)[1] == 9.0 + 5.0 + 7 and ns["sel"](9, 0.0)[1] == -1.0 + 5.0 + 7


def test_translator_refuses_what_it_does_not_know():
    import pytest
    with tempfile.NamedTemporaryFile("w", suffix=".f90", delete=False) as f:
        f.write("      subroutine bad(i)\n      go to (10, 20), i\n   10 continue\n   20 continue\n      end\n")
        path = f.name
    try:
        with pytest.raises(f90fn.TranslateError):
            f90fn.load([path])
    finally:
        os.unlink(path)
