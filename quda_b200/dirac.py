"""Thin Python handles on the C++ operator / solver layer (quda_b200/csrc/host/dirac.h): DiracWilson[PC],
DiracClover[PC] (reference: lib/dirac_wilson.cpp, lib/dirac_clover.cpp) and CG with reliable updates
(lib/inv_cg_quda.cpp).  All arithmetic happens in libquda_b200.so; this module only marshals descriptors."""
import ctypes as C

from . import lib as L

MATPC_EVEN_EVEN, MATPC_ODD_ODD, MATPC_EVEN_EVEN_ASYMMETRIC, MATPC_ODD_ODD_ASYMMETRIC = 0, 1, 2, 3
_TYPES = {"wilson": L.DIRAC_WILSON, "wilsonpc": L.DIRAC_WILSONPC, "clover": L.DIRAC_CLOVER, "cloverpc": L.DIRAC_CLOVERPC,
          "twistedmass": L.DIRAC_TWISTED_MASS, "twistedmasspc": L.DIRAC_TWISTED_MASSPC}


class Dirac:
    def __init__(self, kind, U, kappa, clover=None, clover_inv=None, matpc_type=MATPC_EVEN_EVEN, comm=None, stream=None,
                 mu=0.0):
        self.lib = L.load()
        self.kind, self.U, self.clover, self.clover_inv, self.comm = kind, U, clover, clover_inv, comm  # keep fields alive
        self.prec = U.prec
        h = C.c_void_p()
        X = (C.c_int * 4)(*U.X)
        g = U.desc()
        a = clover.desc() if clover is not None else None
        ai = clover_inv.desc() if clover_inv is not None else None
        L.check(self.lib.b200_dirac_create(C.byref(h), _TYPES[kind], U.prec, X, C.byref(g),
                                           C.byref(a) if a is not None else None,
                                           C.byref(ai) if ai is not None else None, float(kappa), int(matpc_type),
                                           C.byref(comm) if comm is not None else None, stream))
        self.h = h
        if kind.startswith("twistedmass"):
            L.check(self.lib.b200_dirac_set_twist(self.h, float(mu)))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.b200_dirac_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _apply(self, what, out, in_, parity=0, x=None, k=0.0, dagger=False):
        o, i = out.desc(), in_.desc()
        xd = x.desc() if x is not None else None
        L.check(self.lib.b200_dirac_apply(self.h, what, C.byref(o), C.byref(i), parity,
                                          C.byref(xd) if xd is not None else None, float(k), int(bool(dagger))))

    def M(self, out, in_, dagger=False):
        self._apply(L.APPLY_M, out, in_, dagger=dagger)

    def Mdag(self, out, in_):
        self._apply(L.APPLY_MDAG, out, in_)

    def MdagM(self, out, in_):
        self._apply(L.APPLY_MDAGM, out, in_)

    def Dslash(self, out, in_, parity, dagger=False):
        self._apply(L.APPLY_DSLASH, out, in_, parity, dagger=dagger)

    def DslashXpay(self, out, in_, parity, x, k, dagger=False):
        self._apply(L.APPLY_DSLASH_XPAY, out, in_, parity, x, k, dagger=dagger)

    def prepare(self, x, b):
        sp, so = C.c_int(-1), C.c_int(-1)
        xd, bd = x.desc(), b.desc()
        L.check(self.lib.b200_dirac_prepare(self.h, C.byref(xd), C.byref(bd), C.byref(sp), C.byref(so)))
        return sp.value, so.value

    def reconstruct(self, x, b):
        xd, bd = x.desc(), b.desc()
        L.check(self.lib.b200_dirac_reconstruct(self.h, C.byref(xd), C.byref(bd)))


def invert_cg(precise, sloppy, x, b, tol=1e-10, maxiter=10000, delta=0.1):
    """CG on MdagM x = b; returns the filled SolverParam (iter, true_res, secs, gflops, reliable_updates)."""
    p = L.SolverParam()
    p.tol, p.maxiter, p.delta = tol, maxiter, delta
    xd, bd = x.desc(), b.desc()
    L.check(precise.lib.b200_invert_cg(precise.h, sloppy.h if sloppy is not None else None, C.byref(xd), C.byref(bd), C.byref(p)))
    return p
