"""Small dense-matrix helpers with the reference's names (matrix.py:7-51).  Setup-time
only; the ensemble contractions run in libpymd_b200."""
import numpy as N


def chkShape(a):
    """n for an n x n matrix (matrix.py:7-17); raises instead of sys.exit."""
    sh = N.shape(N.asarray(a))
    if len(sh) != 2 or sh[0] != sh[1]:
        raise ValueError("The matrix should be a n by n matrix")
    return sh[0]


def symmetrize(a):
    a = N.asarray(a)
    return 0.5 * (a + a.T)


def antisymmetrize(a):
    a = N.asarray(a)
    return 0.5 * (a - a.T)


def dagger(a):
    a = N.asarray(a)
    chkShape(a)
    return a.conj().T


def hermitianize(a):
    a = N.asarray(a)
    return 0.5 * (a + a.conj().T)


def mdot(*args):
    """Chained matrix product (matrix.py:39-44)."""
    out = N.asarray(args[0])
    for a in args[1:]:
        out = N.dot(out, a)
    return out


mm = mdot
