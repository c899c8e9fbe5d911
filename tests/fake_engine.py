"""TEST DOUBLE for wallgo_b200.engine.BoltzmannEngine backed by the CPU oracle.

Only for CPU tests of the host-side plumbing (drop-in subclass, result wrapping, truncation logic)
in a container without a GPU.  It implements the engine's method surface with NumPy; the product
never imports this file."""
import numpy as np

from oracle import boltzmann_oracle as bo


class _NP:
    """numpy array with the two tensor methods the host code uses."""

    def __init__(self, arr):
        self.arr = arr

    def cpu(self):
        return self

    def numpy(self):
        return self.arr

    def clone(self):
        return _NP(self.arr.copy())

    def copy_(self, other, non_blocking=False):
        self.arr[...] = other.arr if isinstance(other, _NP) else other
        return self

    def __mul__(self, other):
        return _NP(self.arr * (other.arr if isinstance(other, _NP) else other))

    def __getitem__(self, idx):
        return _NP(self.arr[idx])


class _FakeStream:
    cuda_stream = 0

    def __init__(self, device=None):
        pass

    def synchronize(self):
        pass


class _FakeEvent:
    def record(self, stream=None):
        pass

    def query(self):
        return True

    def synchronize(self):
        pass


class _FakeStreamCtx:
    def __init__(self, stream):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class _FakeCuda:
    Stream = _FakeStream
    Event = _FakeEvent
    stream = _FakeStreamCtx

    @staticmethod
    def current_stream(dev=None):
        return _FakeStream()

    @staticmethod
    def mem_get_info(dev=None):
        return (64 << 30, 128 << 30)


class _FakeTorch:
    cuda = _FakeCuda()

    @staticmethod
    def from_numpy(a):
        return a

    @staticmethod
    def empty_like(x):
        return _NP(np.empty_like(x.arr))


class FakeEngine:
    def __init__(self, P, M, N, device=None):
        self.P, self.M, self.N = P, M, N
        self.M1, self.N1 = M - 1, N - 1
        self.q = self.N1 ** 2
        self.n = P * self.M1 * self.q
        self.torch = _FakeTorch()
        self.S = _NP(np.zeros(self.n))
        self.dF_in = _NP(np.zeros(self.n))
        self.A = None
        self.tables = None
        self.launches = 0
        self.dev = None
        self.mats = [None] * 9
        self.stream_ptr = None

    # the slot pool (BoltzmannSolver.slotPool) reads results through these pinned-buffer stand-ins
    @property
    def spectra(self):
        return _NP(self._spectra)

    @property
    def spectra_host(self):
        return _NP(self._spectra.copy())

    @property
    def out(self):
        return self.out_host

    @property
    def out_host(self):
        dF, deltas, err = self._out
        return _NP(np.concatenate([np.ravel(dF), np.ravel(deltas), np.ravel(err), np.zeros(2)]))

    def tmunu_out(self, deltas, msqInterior, velocityMid):
        d = np.asarray(deltas, dtype=float).reshape(4, self.P, self.M1)
        msq = np.asarray(msqInterior, dtype=float).reshape(self.P, self.M1)
        u0 = np.sqrt(1.0 / (1.0 - velocityMid**2))
        u3 = u0 * velocityMid
        ub0, ub3 = u3, u0
        a1 = 3 * d[2] - d[1] - msq * d[0]
        a2 = 3 * d[1] - d[2] + msq * d[0]
        dof = self.dofs[:, None]
        t30 = np.sum(dof * (a1 * u3 * u0 + a2 * ub3 * ub0 + 2 * d[3] * (u3 * ub0 + ub3 * u0)) / 2.0, axis=0)
        t33 = np.sum(dof * ((a1 * u3 * u3 + a2 * ub3 * ub3 + 4 * d[3] * u3 * ub3) / 2.0
                            - (msq * d[0] + d[1] - d[2]) / 2.0), axis=0)
        return t30, t33

    def close(self):
        pass

    # static
    def set_grid(self, chi, rz, rp, pz, pp, dpzdrz, dppdrp):
        self.grid = [np.array(x, dtype=float) for x in (chi, rz, rp, pz, pp, dpzdrz, dppdrp)]

    def set_statistics(self, s):
        self.stat = np.array(s, dtype=float)

    def set_dofs(self, d):
        self.dofs = np.array(d, dtype=float)

    def build_basis(self):
        self.tab = bo.build_tables(self.M, self.N, "Cardinal", "Chebyshev", "Spectral")

    def set_tables(self, Tchi, Dchi, DchiFull, Trz, Drz, Trp):
        ident = lambda m: np.allclose(m, np.identity(m.shape[0]))  # noqa: E731
        basisM = "Cardinal" if ident(Tchi) else "Chebyshev"
        basisN = "Cardinal" if ident(Trz) else "Chebyshev"
        fd = basisM == "Cardinal" and np.count_nonzero(Dchi) <= 3 * Dchi.shape[0]
        self.tab = bo.build_tables(self.M, self.N, basisM, basisN, "Finite Difference" if fd else "Spectral")

    def set_transforms(self, mats):
        self.mats = [None if m is None else np.array(m, dtype=float) for m in mats]

    def get_tables(self):
        t = self.tab
        return dict(Tchi=t.Tchi, Dchi=t.Dchi, DchiFull=t.DchiFull, Trz=t.Trz, Drz=t.Drz, Trp=t.Trp)

    def change_basis(self, stage, src, dst):
        x = src.arr.reshape(self.P, self.M1, self.N1, self.N1)
        for ax in range(3):
            m = self.mats[3 * stage + ax]
            if m is not None:
                x = np.moveaxis(np.tensordot(m, x, axes=([1], [ax + 1])), 0, ax + 1)
        dst.arr[...] = x.reshape(-1)

    def collision_apply(self, x, y):
        _, _, coll = bo.assemble(self.pb, want_collision=True)
        y.arr[...] = coll @ x.arr

    def set_collision(self, c):
        self.coll = np.array(c, dtype=float).reshape(self.P, self.N1, self.N1, self.P, self.N1, self.N1)

    # per solve
    def upload_profiles(self, T, v, msq, dxidchi):
        self.prof = (np.array(T, float), np.array(v, float), np.array(msq, float).reshape(self.P, self.M + 1),
                     np.array(dxidchi, float))

    def _problem(self, vw, mult):
        chi, rz, rp, pz, pp, dpzdrz, dppdrp = self.grid
        T, v, msq, dxidchi = self.prof
        return bo.Problem(tab=self.tab, xi_dxidchi=dxidchi, pz=pz, pp=pp, dpzdrz=dpzdrz, dppdrp=dppdrp, T=T, v=v,
                          msq=msq, vw=vw, statistics=self.stat, collision=self.coll, collisionMultiplier=mult)

    factorTag = None

    def assemble(self, vw, collisionMultiplier=1.0, parts=3, A=None):
        self.factorTag = None
        self.pb = self._problem(vw, collisionMultiplier)
        op, src = bo.assemble(self.pb)
        self.A, self.S = _NP(op), _NP(src.copy())
        return self.A

    def factor(self):
        import scipy.linalg as sla
        self._lu = sla.lu_factor(self.A.arr)     # kept until the next factor(), like the device factors

    def solve(self, b=None):
        import scipy.linalg as sla
        b = self.S if b is None else b
        b.arr[...] = sla.lu_solve(self._lu, b.arr)
        return b

    MAX_REUSE_STEPS = 64

    def refine_steps(self, vw, collisionMultiplier, first, count):
        """Richardson steps of self.S towards the solution of the CURRENT profiles' system with the factors of
        the last factor() (BoltzmannEngine.refine_steps)."""
        import scipy.linalg as sla
        if getattr(self, "rstats", None) is None:
            self.rstats = _NP(np.zeros((self.MAX_REUSE_STEPS, 2)))
            self.rstats_host = self.rstats
        op, src = bo.assemble(self._problem(vw, collisionMultiplier))
        for i in range(first, first + count):
            c = sla.lu_solve(self._lu, src - op @ self.S.arr)
            self.S.arr += c
            self.rstats.arr[i] = (np.max(np.abs(c)), np.max(np.abs(self.S.arr)))
        self.refine_calls = getattr(self, "refine_calls", 0) + count

    def refine_history(self, steps):
        return self.rstats.arr[:steps].copy()

    def _onEngineStream(self):
        import contextlib
        return contextlib.nullcontext()

    @property
    def nout(self):
        return self.n + 4 * self.P * self.M1 + 4 * self.P + 2

    def factor_solve(self, b=None):
        self.factor()
        return self.solve(b)

    last_refinement = (0.0, 1.0)

    def refine(self, vw, collisionMultiplier=1.0, x=None):
        return self.S if x is None else x      # the LAPACK solution stands in for the refined one

    def info(self):
        return 0

    def spectra_async(self, dF):
        self._dF = np.array(dF.arr if isinstance(dF, _NP) else dF).reshape(self.P, self.M1, self.N1, self.N1)
        c = np.abs(bo.to_chebyshev(self._dF, self.tab))
        self._spectra = np.concatenate([c.sum(axis=(0, 2, 3)), c.sum(axis=(0, 1, 3)), c.sum(axis=(0, 1, 2))])

    def spectra_to_host(self):
        return self._spectra.copy()

    def moments_async(self, dF, keep, roundtrip):
        c = bo.to_chebyshev(self._dF, self.tab).copy()
        c[:, keep[0]:] = 0
        c[:, :, keep[1]:] = 0
        c[:, :, :, keep[2]:] = 0
        a = np.abs(c)
        err = np.stack([a.sum(axis=(1, 2, 3)), a[:, keep[0] - 1].sum(axis=(1, 2)),
                        a[:, :, keep[1] - 1].sum(axis=(1, 2)), a[:, :, :, keep[2] - 1].sum(axis=(1, 2))], axis=1)
        dFo = bo.from_chebyshev(c, self.tab) if roundtrip else self._dF
        pb = self._problem(0.0, 1.0)
        self._out = (np.array(dFo), bo.moments(dFo, pb), err)

    def results_to_host(self):
        return self._out
