"""The statement-by-statement Fortran translator of oracle/fortran_exec.py (test infrastructure) on snippets with
known answers: 1-based inclusive sections, first-index-fastest storage, loops, one-line if, else if, cycle/exit,
continuations, comments, parameters, complex arithmetic, locals vs module variables, mpi_allreduce on one rank."""
import numpy as np

from oracle import fortran_exec as fx

SRC = """
module m
    contains
        subroutine fill(n)
            implicit none
            integer, intent(in) :: n
            integer :: i, j   ! locals
            real, parameter :: half = 1.0d0/2.0
            real :: radius    ! shadows the module variable of the same name
            radius = -1.0
            total = 0
            do j = 1, 3
                do i = 1, n
                    if (i == 2 .and. j == 2) cycle
                    a(i,j) = 10*i + j &
                        + half      ! continuation and trailing comment
                    if (a(i,j) > 40.) then
                        total = total + 1
                    else if (a(i,j) > 30.) then
                        total = total + 100
                    else
                        total = total + 10000
                    endif
                enddo
            enddo
            b(2:3) = a(1,2:3) ; b(1) = real(size(a))
            z(:) = cmplx(0, b(:)) * conjg(cmplx(1., 2.))
            do i = 5, 1, -2
                last = i
                if (i < 3) exit
            end do
            call mpi_allreduce(total, total_all, 1, mpi_realtype, mpi_sum, mpi_comm_world, ierr)
        end subroutine fill
end module m
"""


def test_translator_semantics(tmp_path):
    f = tmp_path / "m.f90"
    f.write_text(SRC)
    ns = fx.base_namespace()
    a = np.zeros((3, 4))                     # C storage [j, i] = Fortran a(i, j), i = 1..4 fastest
    ns.update(a=fx.FArray(a.T), b=fx.FArray(np.zeros(3)), z=fx.FArray(np.zeros(3, dtype=complex)), radius=30.0, total=-5,
              total_all=None, last=None, mpi_realtype=None, mpi_sum=None, mpi_comm_world=None, ierr=0)
    src = fx.load(ns, str(f), ["fill"])
    assert "global last, total, total_all" in src["fill"]          # radius, i, j stay local
    ns["fill"](4)
    want = np.array([[10 * i + j + 0.5 for i in range(1, 5)] for j in range(1, 4)])
    want[1, 1] = 0.0                                                 # skipped by `cycle`
    assert np.array_equal(a, want)
    assert ns["radius"] == 30.0                                      # the local did not leak
    # a > 40: i = 4 (3 values); 30 < a <= 40: i = 3 (3 values); else 5 values (one skipped)
    assert ns["total"] == 3 + 300 + 50000 and ns["total_all"] == ns["total"]
    assert np.array_equal(ns["b"].a, [12.0, 12.5, 13.5])
    assert np.array_equal(ns["z"].a, 1j * ns["b"].a * (1 - 2j))
    assert ns["last"] == 1                                           # 5, 3, 1 with exit at 1


def test_statement_splitting():
    st = fx.statements("  x = 1 ! c\n  y = 'a!b' ; z = 2 &\n    & + 3\n")
    assert [' '.join(s.split()) for s in st] == ["x = 1", "y = 'a!b'", "z = 2 + 3"]


SRC2 = """
subroutine dec(ngrid, nproc, i_offset, i_size)
    integer,intent(in) :: ngrid, nproc
    integer,intent(out),allocatable,dimension(:) :: i_offset, i_size
    integer :: normal_size, p
    allocate(i_offset(nproc), i_size(nproc))
    normal_size = ngrid / nproc          ! INTEGER division
    half = 7/2 + (-7)/2 + nx/2+1          ! literals and a module integer
    i_offset(1) = 0
    do p = 1, nproc
        i_size(p) = normal_size
        if (p > 1) i_offset(p) = i_offset(p-1) + i_size(p-1)
    enddo
    call mpi_comm_rank(mpi_comm_world, me, ierr)
    call mpi_type_create_subarray(3, [nx/2+1, 2, 3], [1, 1, 1], [0, 0, 0], 0, 0, types(2), ierr)
end subroutine
"""


def test_integer_semantics_and_output_arguments(tmp_path):
    f = tmp_path / "d.f90"
    f.write_text(SRC2)
    ns = fx.base_namespace()
    off, siz = fx.FArray(np.zeros(3, dtype=np.int64)), fx.FArray(np.zeros(3, dtype=np.int64))
    types = fx.FArray(np.empty(2, dtype=object))
    ns.update(nx=fx.FInt(16), half=None, me=None, ierr=0, mpi_comm_world="w", types=types,
              mpi_comm_rank=lambda comm: 5, mpi_type_create_subarray=lambda nd, full, sub, st, o, t: (list(full), list(sub), list(st)))
    fx.load(ns, str(f), ["dec"])
    ns["dec"](fx.FInt(16), fx.FInt(3), off, siz)
    assert list(siz.a) == [5, 5, 5] and list(off.a) == [0, 5, 10]          # 16 / 3 = 5, not 5.33
    assert ns["half"] == 3 + (-3) + 9 and isinstance(ns["half"], int)       # truncation toward zero; 16/2+1 = 9
    assert ns["me"] == 5
    assert types.a[1] == ([9, 2, 3], [1, 1, 1], [0, 0, 0]) and types.a[0] is None
