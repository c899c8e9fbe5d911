"""B200-native `BoltzmannSolver`: same public interface as the reference class
(`/root/reference/src/WallGo/boltzmann.py:37-886`), every numerical step on the GPU.

Host Python keeps only what the reference keeps outside NumPy kernels: evaluating the user's
`msqVacuum` callbacks on the M+1 profile points, the branchy truncation decision on the
(M-1)+2(N-1) device-reduced spectra (`np.polyfit`, polynomial.py:994-1033) and wrapping the
results.  Per solve the boundary traffic is one upload of O(P*M) doubles and one download of
O(n) doubles.

There is no CPU path: constructing the engine without an sm_100 device raises.
"""
from __future__ import annotations

import sys
import typing
from copy import deepcopy

import numpy as np

from . import containers as _own
from .containers import ETruncationOption
from .spectral import basisChangeMatrices, buildTables, spectralPeak, tailConverging


class _OwnTypes:
    Polynomial = _own.Polynomial
    BoltzmannDeltas = _own.BoltzmannDeltas
    BoltzmannResults = _own.BoltzmannResults


class _Keep(list):
    """Truncated shape per axis as the reference reports it, plus `.dev`: what the device keeps."""

    def __init__(self, sizes):
        super().__init__(sizes)
        self.dev = list(sizes)


REUSE_FROM = 2          # first position in a sequence that tries the stale factors (0 and 1 always factor)
REUSE_CHUNK = 4         # refinement steps enqueued between two looks at the history
REUSE_TOL = 1e-13       # max|correction| / max|x| at which the iteration has converged
REUSE_MAX_STEPS = 32
REUSE_MIN_GAIN = 0.2    # a chunk must shrink the correction at least to this fraction (0.67 per step)


def reuseVerdict(history, maxSteps=REUSE_MAX_STEPS):
    """What to do behind a chunk of stale-factor refinement steps.  history[k] = (max|correction|, max|x|) of step
    k.  "done": converged (last ratio <= REUSE_TOL); "more": contracting, enqueue the next chunk; "fail": not
    contracting fast enough (after the second chunk: less than 5x per chunk), not finite, or out of steps --
    factor afresh."""
    h = np.asarray(history, dtype=np.float64).reshape(-1, 2)
    steps = h.shape[0]
    with np.errstate(all="ignore"):
        ratio = h[:, 0] / h[:, 1]
    last = ratio[-1]
    if not np.isfinite(last):
        return "fail"
    if last <= REUSE_TOL:
        return "done"
    contracting = steps < 2 * REUSE_CHUNK or last <= REUSE_MIN_GAIN * ratio[-1 - REUSE_CHUNK]
    return "more" if contracting and steps + REUSE_CHUNK <= maxSteps else "fail"


class _Slot:
    """What one busy slot of a _SlotPool is doing."""

    __slots__ = ("token", "phase", "event", "keep", "peaks", "raw", "steps", "fresh")

    def __init__(self, token, phase, event, raw=None, steps=0, fresh=True):
        self.token, self.phase, self.event = token, phase, event
        self.keep = self.peaks = None
        self.raw, self.steps, self.fresh = raw, steps, fresh


class _SlotPool:
    """Slots of independent systems in flight on one GPU (see BoltzmannSolverB200Mixin.slotPool).

    reuseFactors=True adds FACTOR RE-USE for sequences of nearby systems (the successive Boltzmann solves inside
    one `EOM.wallPressure`, equationOfMotion.py:1133-1449: same wall velocity and grid mapping, slowly changing
    profiles).  A caller that submits such a sequence to a FIXED slot with its position `seq` in the sequence
    gets, from REUSE_FROM on, the solution of the new system by Richardson iteration preconditioned with the LU
    factors the slot still holds from the sequence's last factorisation, started from the previous solution:
    x <- x + (LU_old)^-1 (S_new - A_new x), the residual formed in double-double arithmetic from the operator
    regenerated on the fly (wg_refine; A_new is never stored).  Steps are enqueued REUSE_CHUNK at a time; the
    host looks at the history of max|correction| / max|x| behind every chunk and stops at REUSE_TOL (the answer
    then agrees with a fresh LU + refinement step to ~1e-14), or falls back to the fresh factorisation when the
    iteration does not contract (profiles changed too much) or REUSE_MAX_STEPS are used up.  A step costs ~1 ms
    of mostly idle GPU against ~40-100 ms of full-GPU tensor work for a factorisation.  Everything is
    deterministic, so a sequence gives the same bits wherever and beside whatever it runs."""

    REUSE_FROM, REUSE_CHUNK, REUSE_TOL, REUSE_MAX_STEPS = REUSE_FROM, REUSE_CHUNK, REUSE_TOL, REUSE_MAX_STEPS

    def __init__(self, solver, count, reuseFactors=False):
        import ctypes

        self.sibs = solver._siblings(count)
        self.torch = self.sibs[0]._getEngine().torch
        for s_ in self.sibs:
            eng = s_._getEngine()
            if getattr(s_, "_stream", None) is None:
                s_._stream = self.torch.cuda.Stream(device=eng.dev)
            eng.stream_ptr = ctypes.c_void_p(s_._stream.cuda_stream)   # pinned while the pool is open
        self.state = [None] * count      # per slot: None (free) or a _Slot
        self.reuse = bool(reuseFactors)
        self.maxFresh = None             # at most this many factorisations in flight (None: every slot may factor)
        self.reuseStats = dict(fresh=0, stale=0, fallbacks=0, steps=0)

    def close(self):
        for s_ in self.sibs:
            if s_._engine is not None:
                s_._engine.stream_ptr = None

    def freeSlots(self) -> int:
        return sum(1 for st in self.state if st is None)

    def busy(self) -> bool:
        return any(st is not None for st in self.state)

    def _factoring(self) -> int:
        return sum(1 for st in self.state if st is not None and st.fresh and st.phase == 0)

    @staticmethod
    def _factorTag(raw):
        return (float(raw[4]), np.asarray(raw[3], dtype=np.float64).tobytes())

    def _finishSolve(self, s_, eng, token, raw, fresh):
        """spectra of the solution in eng.S + their download; the slot then waits in phase 0"""
        torch = self.torch
        eng.spectra_async(eng.S)
        eng.spectra_host.copy_(eng.spectra, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(s_._stream)
        return _Slot(token, 0, ev, raw=raw, fresh=fresh)

    def submit(self, token, background=None, raw=None, slot=None, seq=None) -> bool:
        """Enqueue one system (a BoltzmannBackground, or raw = (T, v, msq, dxidchi, vw)) on a free slot (or on
        slot `slot`).  seq: position of the system in a sequence of nearby systems submitted to this slot (see the
        class docstring; needs reuseFactors, raw and a fixed slot).  Returns False when the slot (every slot) is
        busy or maxFresh factorisations are already in flight."""
        if slot is None:
            for slot, st in enumerate(self.state):
                if st is None:
                    break
            else:
                return False
        elif self.state[slot] is not None:
            return False
        s_ = self.sibs[slot]
        torch = self.torch
        stale = False
        if self.reuse and raw is not None and seq is not None and int(seq) >= self.REUSE_FROM:
            eng = s_._getEngine()      # (a change of the static data re-uploads it and drops the factor tag)
            stale = getattr(eng, "factorTag", None) == self._factorTag(raw)
        if not stale and self.maxFresh is not None and self._factoring() >= self.maxFresh:
            return False
        with torch.cuda.stream(s_._stream):
            if stale:
                eng = s_._staleStartOnDevice(raw, 0, self.REUSE_CHUNK)
                ev = torch.cuda.Event()
                ev.record(s_._stream)
                self.state[slot] = _Slot(token, "R", ev, raw=raw, steps=self.REUSE_CHUNK, fresh=False)
                return True
            if raw is None:
                s_.setBackground(background)
                eng = s_._solveOnDevice()
                eng.factorTag = None
            else:
                eng = s_._solveRawOnDevice(*raw)
                eng.factorTag = self._factorTag(raw) if self.reuse else None
            self.reuseStats["fresh"] += 1
            self.state[slot] = self._finishSolve(s_, eng, token, raw, True)
        return True

    def _advanceReuse(self, slot):
        """Behind a chunk of stale-factor refinement steps: converged -> spectra; contracting -> next chunk;
        otherwise -> fresh factorisation of the same system."""
        st = self.state[slot]
        s_ = self.sibs[slot]
        eng = s_._engine
        torch = self.torch
        verdict = reuseVerdict(eng.rstats_host.numpy()[:st.steps], min(REUSE_MAX_STEPS, eng.MAX_REUSE_STEPS))
        C = REUSE_CHUNK
        with torch.cuda.stream(s_._stream):
            if verdict == "done":
                eng.out[eng.nout - 2:].copy_(eng.rstats[st.steps - 1])     # lastRefinement of this solve
                self.reuseStats["stale"] += 1
                self.reuseStats["steps"] += st.steps
                self.state[slot] = self._finishSolve(s_, eng, st.token, st.raw, False)
                return
            if verdict == "more":
                s_._staleStartOnDevice(st.raw, st.steps, C)
                ev = torch.cuda.Event()
                ev.record(s_._stream)
                st.event, st.steps = ev, st.steps + C
                return
            self.reuseStats["fallbacks"] += 1
            self.reuseStats["steps"] += st.steps
            self.reuseStats["fresh"] += 1
            eng = s_._solveRawOnDevice(*st.raw)
            eng.factorTag = self._factorTag(st.raw)
            self.state[slot] = self._finishSolve(s_, eng, st.token, st.raw, True)

    def _advance(self, slot, finished):
        st = self.state[slot]
        s_ = self.sibs[slot]
        eng = s_._engine
        torch = self.torch
        if st.phase == "R":
            self._advanceReuse(slot)
            return
        if st.phase == 0:
            keep, peaks = s_._decideTruncation(eng.spectra_host.numpy().copy())
            with torch.cuda.stream(s_._stream):
                eng.moments_async(eng.S, keep.dev, s_.truncationOption.name != "NONE")
                eng.out_host.copy_(eng.out, non_blocking=True)
                ev2 = torch.cuda.Event()
                ev2.record(s_._stream)
            st.phase, st.event, st.keep, st.peaks = 1, ev2, keep, peaks
            return
        keep, peaks = st.keep, st.peaks
        if eng.info() != 0:
            eng.factorTag = None
            raise np.linalg.LinAlgError("Singular matrix")
        h = eng.out_host.numpy()
        pm = 4 * eng.P * eng.M1
        dF = h[:eng.n].reshape(eng.P, eng.M1, eng.N1, eng.N1).copy()
        deltas = h[eng.n:eng.n + pm].reshape(4, eng.P, eng.M1).copy()
        err = h[eng.n + pm:eng.n + pm + 4 * eng.P].reshape(eng.P, 4).copy()
        s_.lastRefinement = (float(h[-2]), float(h[-1])) if getattr(s_, "refinementSteps", 1) else None
        self.state[slot] = None
        finished.append((st.token, (dF, deltas, err, keep, peaks) if self.rawResults else
                         s_._wrapResults(dF, deltas, err, keep, peaks)))

    rawResults = False   # True: poll() returns the (dF, deltas, err, keep, peaks) tuples instead of BoltzmannResults

    def poll(self):
        finished = []
        for slot, st in enumerate(self.state):
            if st is not None and st.event.query():
                self._advance(slot, finished)
        return finished

    def waitOldest(self):
        live = [(st.token if isinstance(st.token, int) else k, k) for k, st in enumerate(self.state) if st is not None]
        if live:
            self.state[min(live)[1]].event.synchronize()


class BoltzmannSolverB200Mixin:
    """Implementation shared by the standalone class below and the WallGo drop-in subclass."""

    MAX_EXPONENT: typing.Final[float] = sys.float_info.max_exp * np.log(2)
    _types = _OwnTypes

    # ------------------------------------------------------------------ construction / setters
    def _initB200(self, grid, basisM, basisN, derivatives, collisionMultiplier, truncationOption, device):
        self.grid = grid
        self._checkDerivatives(derivatives)
        self.derivatives = derivatives
        self._checkBasis(basisM)
        self._checkBasis(basisN)
        if derivatives == "Finite Difference":
            assert basisM == "Cardinal" and basisN == "Cardinal", \
                "Must use Cardinal basis for Finite Difference method"
        self.basisM, self.basisN = basisM, basisN
        self.collisionMultiplier = collisionMultiplier
        self.truncationOption = truncationOption
        self.background = None
        self.collisionArray = None
        self.offEqParticles = []
        self.device = device
        self._engine = None
        self._staticKey = None
        # Steps of mixed-precision iterative refinement after the LU solve (0 or 1; wg_refine): removes the
        # cond(A)*eps forward error of the FP64 solve for ~1 % of its cost.  lastRefinement = (max|correction|,
        # max|deltaF|) of the most recent getDeltas(); the ratio / 2^-53 estimates the condition number.
        self.refinementSteps = 1
        self.lastRefinement = None
        # Opt-in: consecutive getDeltas() calls with unchanged wall velocity and grid mapping re-use the LU factors
        # of their last factorisation from the third call on (_solveMaybeStale); results agree with the plain path
        # to ~1e-13 instead of bit for bit.  The wall-velocity scan server has its own switch (vwscan).
        self.reuseFactors = bool(getattr(type(self), "reuseFactorsDefault", False))

    def setBackground(self, background) -> None:
        """Deep-copies and boosts to the plasma frame (boltzmann.py:116-123)."""
        self.background = deepcopy(background)
        self.background.boostToPlasmaFrame()

    def setCollisionArray(self, collisionArray) -> None:
        self.collisionArray = collisionArray
        self._staticKey = None           # re-upload the tensor at the next solve

    def updateParticleList(self, offEqParticles) -> None:
        for p in offEqParticles:
            assert hasattr(p, "msqVacuum") and hasattr(p, "statistics"), "not a Particle"
        self.offEqParticles = offEqParticles
        self._staticKey = None

    def loadCollisions(self, directoryPath) -> None:
        """Loads one collision file per ordered pair of out-of-equilibrium species from `directoryPath`
        (boltzmann.py:822-844 -> CollisionArray.newFromDirectory); raises CollisionLoadError when a file is
        missing, larger grids are requested than the files hold, or the format cannot be read here."""
        self.collisionArray = _own.CollisionArray.newFromDirectory(
            directoryPath, self.grid, self.basisN, self.offEqParticles, device=self.device)
        self._staticKey = None

    # deepcopy (used by EOM.getBoltzmannFiniteDifference, equationOfMotion.py:2135) must not clone
    # the device handle; the copy builds its own engine lazily.
    def __getstate__(self):
        st = dict(self.__dict__)
        st["_engine"] = None
        st["_staticKey"] = None
        st["_sibs"] = None
        st["_stream"] = None
        st["_batchEng"] = None
        return st

    def __deepcopy__(self, memo):
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_engine", "_staticKey", "_sibs", "_stream", "_batchEng"):
                setattr(new, k, None)
            elif k == "grid":
                setattr(new, k, deepcopy(v, memo))
            else:
                setattr(new, k, deepcopy(v, memo))
        return new

    # ------------------------------------------------------------------ engine plumbing
    def _collisionData(self):
        ca = self.collisionArray
        assert ca is not None, "BoltzmannSolver error: collision array not set"
        return np.asarray(ca[...])

    def _getEngine(self):
        from .engine import BoltzmannEngine

        P = len(self.offEqParticles)
        assert P >= 1, "BoltzmannSolver error: no out-of-equilibrium particles"
        g = self.grid
        if self._engine is None or (self._engine.P, self._engine.M, self._engine.N) != (P, g.M, g.N):
            if self._engine is not None:
                cur = getattr(self, "_batchEng", None)
                if cur is not None:                 # batch contexts borrow the engine's handle: they go first
                    for b in cur[2]:
                        b.close()
                    self._batchEng = None
                self._engine.close()
            self._engine = BoltzmannEngine(P, g.M, g.N, self.device)
            self._staticKey = None
        eng = self._engine
        coll = self._collisionData()
        stats = tuple(-1.0 if p.statistics == "Fermion" else 1.0 for p in self.offEqParticles)
        _, dpzdrz, dppdrp = g.getCompactificationDerivatives()
        # content fingerprint of the tensor (an in-place edit by the caller must not leave a stale copy on the
        # GPU): the full sum up to 2^21 entries, a strided sample of that size beyond
        flat = coll.reshape(-1)
        step = max(1, flat.size >> 21)
        dofs = tuple(float(p.totalDOFs) for p in self.offEqParticles)
        key = (self.basisM, self.basisN, self.derivatives, stats, dofs, id(g), coll.__array_interface__["data"][0],
               coll.shape, float(g.momentumFalloffT), float(np.sum(flat[::step])), float(np.sum(np.abs(flat[::step]))),
               float(flat[0]), float(flat[-1]))
        if key != self._staticKey:
            chi, rz, rp = g.getCompactCoordinates()
            _, pz, pp = g.getCoordinates()
            eng.set_grid(chi, rz, rp, pz, pp, dpzdrz, dppdrp)
            eng.set_statistics(np.array(stats))
            eng.set_dofs(np.array([float(p.totalDOFs) for p in self.offEqParticles]))
            if (self.basisM, self.basisN, self.derivatives) == ("Cardinal", "Chebyshev", "Spectral"):
                eng.build_basis()
            else:
                eng.set_tables(*buildTables(g.M, g.N, self.basisM, self.basisN, self.derivatives))
            eng.set_transforms(basisChangeMatrices(g.M, g.N, self.basisM, self.basisN))
            assert coll.shape == (P, g.N - 1, g.N - 1, P, g.N - 1, g.N - 1), "collision tensor has the wrong shape"
            eng.set_collision(coll)
            self._staticKey = key
            eng.factorTag = None       # factors of an operator built from other static data are not re-used
        return eng

    def _msq(self, fields):
        return np.array([np.asarray(p.msqVacuum(fields), dtype=np.float64) for p in self.offEqParticles])

    def _uploadBackground(self, eng):
        bg = self.background
        assert bg is not None, "BoltzmannSolver error: background not set"
        dxidchi, _, _ = self.grid.getCompactificationDerivatives()
        msq = self._msq(bg.fieldProfiles)
        eng.upload_profiles(np.asarray(bg.temperatureProfile, dtype=np.float64),
                            np.asarray(bg.velocityProfile, dtype=np.float64), msq, dxidchi)

    def _solveRawOnDevice(self, T, v, msq, dxidchi, vw):
        """assemble + LU + solves from already evaluated profiles (boosted T, v on the M+1 nodes, m^2 of every
        species there, dxi/dchi of the caller's own grid mapping): the form the scan server receives from its
        host workers, whose grids differ per wall velocity (equationOfMotion.py:1457-1503)."""
        eng = self._getEngine()
        eng.upload_profiles(np.asarray(T, dtype=np.float64), np.asarray(v, dtype=np.float64),
                            np.asarray(msq, dtype=np.float64), np.asarray(dxidchi, dtype=np.float64))
        eng.assemble(float(vw), float(self.collisionMultiplier))
        eng.factor()
        eng.solve()
        if getattr(self, "refinementSteps", 1):
            eng.refine(float(vw), float(self.collisionMultiplier))
        return eng

    def _staleStartOnDevice(self, raw, first, count):
        """`count` Richardson steps towards the solution of the system `raw` = (T, v, msq, dxidchi, vw), starting
        from whatever eng.S holds (the previous solution of the sequence) with the factors eng.A holds
        (_SlotPool, reuseFactors).  first = 0 uploads the new profiles."""
        eng = self._engine
        T, v, msq, dxidchi, vw = raw
        if first == 0:
            eng.upload_profiles(np.asarray(T, dtype=np.float64), np.asarray(v, dtype=np.float64),
                                np.asarray(msq, dtype=np.float64), np.asarray(dxidchi, dtype=np.float64))
        eng.refine_steps(float(vw), float(self.collisionMultiplier), first, count)
        return eng

    def _solveMaybeStale(self):
        """_solveOnDevice() for getDeltas(), with FACTOR RE-USE when self.reuseFactors is set: consecutive
        getDeltas() calls with the same wall velocity and grid mapping (the pressure iteration of one
        EOM.wallPressure, equationOfMotion.py:1133-1449) form a sequence, and from its third member on the
        solution is first sought by refinement steps on the factors of the sequence's last factorisation (see
        _SlotPool; same constants, same verdicts, synchronous here).  self.reuseStats counts what happened."""
        if not getattr(self, "reuseFactors", False):
            return self._solveOnDevice()
        eng = self._getEngine()
        bg = self.background
        assert bg is not None, "BoltzmannSolver error: background not set"
        dxidchi, _, _ = self.grid.getCompactificationDerivatives()
        tag = (float(bg.velocityWall), np.asarray(dxidchi, dtype=np.float64).tobytes())
        seq = getattr(self, "_seq", -1) + 1 if tag == getattr(self, "_seqTag", None) else 0
        self._seqTag, self._seq = tag, seq
        stats = self.__dict__.setdefault("reuseStats", dict(fresh=0, stale=0, fallbacks=0, steps=0))
        if seq >= REUSE_FROM and getattr(eng, "factorTag", None) == tag:
            self._uploadBackground(eng)
            vw, mult = float(bg.velocityWall), float(self.collisionMultiplier)
            steps, verdict = 0, "more"
            while verdict == "more":
                eng.refine_steps(vw, mult, steps, REUSE_CHUNK)
                steps += REUSE_CHUNK
                verdict = reuseVerdict(eng.refine_history(steps), min(REUSE_MAX_STEPS, eng.MAX_REUSE_STEPS))
            stats["steps"] += steps
            if verdict == "done":
                stats["stale"] += 1
                with eng._onEngineStream():
                    eng.out[eng.nout - 2:].copy_(eng.rstats[steps - 1])
                return eng
            stats["fallbacks"] += 1
        stats["fresh"] += 1
        eng = self._solveOnDevice()
        eng.factorTag = tag
        return eng

    def _solveOnDevice(self, fused: bool = False):
        """assemble + LU + triangular solves; leaves deltaF in eng.S (device).  fused=True uses wg_factor_solve
        (the forward solve rides along inside the factorisation): 0.6 ms lower latency for one n=15600 system
        alone, but 1.7 % less throughput when several systems share the GPU, so the default is the two
        separate calls; one path for single solves and batches keeps their results bit-identical."""
        eng = self._getEngine()
        self._uploadBackground(eng)
        eng.assemble(float(self.background.velocityWall), float(self.collisionMultiplier))
        if fused:
            eng.factor_solve()
        else:
            eng.factor()
            eng.solve()
        if getattr(self, "refinementSteps", 1):
            eng.refine(float(self.background.velocityWall), float(self.collisionMultiplier))
        return eng

    def _shape(self):
        g = self.grid
        return (len(self.offEqParticles), g.M - 1, g.N - 1, g.N - 1)

    # ------------------------------------------------------------------ queries
    def solveBoltzmannEquations(self) -> np.ndarray:
        """delta f = operator^-1 source (boltzmann.py:240-294)."""
        eng = self._solveOnDevice()
        info = eng.info()
        if info != 0:
            raise np.linalg.LinAlgError("Singular matrix")
        return eng.S.cpu().numpy().reshape(self._shape())

    def buildLinearEquations(self):
        """(operator (n,n), source (n,), liouville, collision) on the host (boltzmann.py:628-820).
        Debug/test entry point: the production path never moves the operator off the device."""
        eng = self._getEngine()
        self._uploadBackground(eng)
        vw, mult = float(self.background.velocityWall), float(self.collisionMultiplier)
        shape8 = self._shape() * 2
        eng.assemble(vw, mult, 1)
        liouville = eng.A.cpu().numpy().reshape(shape8)
        eng.assemble(vw, mult, 2)
        collision = eng.A.cpu().numpy().reshape(shape8)
        eng.assemble(vw, mult, 3)
        operator = eng.A.cpu().numpy()
        source = eng.S.cpu().numpy()
        return operator, source, liouville, collision

    def _decideTruncation(self, spectra):
        """Host half of checkSpectralConvergence (boltzmann.py:399-483) on device-reduced spectra.

        Returns (keep, peaks): keep[axis] is the truncated shape the reference reports; keep.dev[axis] is the
        number of coefficients the device actually retains.  They differ only on grids with fewer than three
        coefficients along an axis: there the reference's cut is `-(size // 3) == 0`, its slice `[-0:]` is the
        WHOLE axis, so a truncation zeroes everything while the reported shape stays unchanged
        (boltzmann.py:400-446) -- reproduced here for drop-in fidelity."""
        g = self.grid
        M1, N1 = g.M - 1, g.N - 1
        sChi, sPz, sPp = spectra[:M1], spectra[M1:M1 + N1], spectra[M1 + N1:]
        cutM, cutN = M1 // 3, N1 // 3
        conv = (tailConverging(sChi[M1 - cutM:] if cutM else sChi, M1 - cutM),
                tailConverging(sPz[N1 - cutN:] if cutN else sPz, N1 - cutN),
                tailConverging(sPp[N1 - cutN:] if cutN else sPp, N1 - cutN))
        keep = _Keep([M1, N1, N1])
        if self.truncationOption == ETruncationOption.AUTO or self.truncationOption.name == "AUTO":
            cut = [not c for c in conv]
        elif self.truncationOption.name == "THIRD":
            cut = [True, True, True]
        else:
            cut = [False, False, False]
        for axis, (size, c) in enumerate(((M1, cutM), (N1, cutN), (N1, cutN))):
            if cut[axis]:
                keep[axis] = size - c
                keep.dev[axis] = size - c if c > 0 else 0
        peaks = (spectralPeak(sChi[:keep[0]], 0), spectralPeak(sPz[:keep[1]], 1),
                 spectralPeak(sPp[:keep[2]], 2))
        return keep, peaks

    def _toDevice(self, eng, deltaF):
        arr = np.ascontiguousarray(deltaF, dtype=np.float64).reshape(-1)
        assert arr.size == eng.n, "deltaF has the wrong shape"
        eng.dF_in.copy_(eng.torch.from_numpy(arr), non_blocking=False)
        return eng.dF_in

    def _spectralPipeline(self, eng, dFdev):
        eng.spectra_async(dFdev)
        spectra = eng.spectra_to_host()
        keep, peaks = self._decideTruncation(spectra)
        roundtrip = self.truncationOption.name != "NONE"
        eng.moments_async(dFdev, keep.dev, roundtrip)
        dF, deltas, err = eng.results_to_host()
        return dF, deltas, err, keep, peaks

    def checkSpectralConvergence(self, deltaF):
        """(deltaFTruncated, shapeTruncated, spectralPeaks) (boltzmann.py:353-491)."""
        eng = self._getEngine()
        if self.background is not None:
            self._uploadBackground(eng)
        dF, _, _, keep, peaks = self._spectralPipeline(eng, self._toDevice(eng, deltaF))
        shape = (deltaF.shape[0], keep[0], keep[1], keep[2])
        if self.truncationOption.name == "NONE":
            return deltaF, shape, peaks
        return dF.reshape(deltaF.shape), shape, peaks

    def estimateTruncationError(self, deltaF, shapeTruncated) -> float:
        """max over axes/species of |last kept coefficient| / sum |coefficients| (boltzmann.py:296-351)."""
        eng = self._getEngine()
        if self.background is not None:
            self._uploadBackground(eng)
        dev = self._toDevice(eng, deltaF)
        eng.spectra_async(dev)
        eng.moments_async(dev, shapeTruncated[1:], False)
        _, _, err = eng.results_to_host()
        return float(np.max(err[:, 1:] / err[:, :1]))

    def getDeltas(self, deltaF=None):
        """Solve (unless deltaF is given), truncate, integrate the four moments (boltzmann.py:141-238)."""
        if deltaF is None:
            eng = self._solveMaybeStale()
            dFdev = eng.S
        else:
            eng = self._getEngine()
            self._uploadBackground(eng)
            dFdev = self._toDevice(eng, deltaF)
        dF, deltas, err, keep, peaks = self._spectralPipeline(eng, dFdev)
        if deltaF is None:
            info = eng.info()
            if info != 0:
                eng.factorTag = None
                raise np.linalg.LinAlgError("Singular matrix")
            self.lastRefinement = eng.last_refinement if getattr(self, "refinementSteps", 1) else None
        return self._wrapResults(dF, deltas, err, keep, peaks)

    def conditionEstimate(self):
        """Condition number the last solve experienced, from the refinement step: (max|correction| / max|deltaF|)
        / 2^-53, i.e. the forward error of the plain LU solve in units of the rounding unit.  A by-product (no extra
        device work); None when refinement is off or nothing has been solved yet.  SURVEY.md section 7 asks for a
        cheap estimate to decide on refinement; here the step always runs (about 1 % of a solve) and reports it."""
        lr = getattr(self, "lastRefinement", None)
        if not lr or not lr[1] > 0.0:
            return None
        return float(lr[0] / lr[1] / 2.0 ** -53)

    def getFeedback(self):
        """Out-of-equilibrium source terms of the wall EOM computed on the device from the moments of the
        LAST getDeltas() call: dict with T30, T33 (EOM.deltaToTmunu, equationOfMotion.py:2021-2127, all
        interior z at once), potentialOut (:1545-1555) and dVout (:1418-1429)."""
        eng = self._getEngine()
        bg = self.background
        fields = bg.fieldProfiles
        dmsq = None
        if all(getattr(p, "msqDerivative", None) is not None for p in self.offEqParticles):
            vals = [np.asarray(p.msqDerivative(fields), dtype=np.float64) for p in self.offEqParticles]
            vals = [v.reshape(v.shape[0], -1) for v in vals]
            dmsq = np.stack([v[1:-1] for v in vals])          # [P, M-1, F]
        t30, t33, pot, dv = eng.feedback(dmsq, float(bg.velocityMid))
        return dict(T30=t30, T33=t33, potentialOut=pot, dVout=dv)

    def tmunuOut(self, offEquilDeltas, fields, velocityMid: float):
        """(T30_out, T33_out) on all interior z nodes for an arbitrary BoltzmannDeltas (EOM.deltaToTmunu,
        equationOfMotion.py:2021-2127, which the reference evaluates point by point): one wg_feedback launch.
        `fields` are the M-1 interior field points the EOM hands to findPlasmaProfile."""
        eng = self._getEngine()
        d = offEquilDeltas
        table = np.stack([np.asarray(d.Delta00.coefficients), np.asarray(d.Delta02.coefficients),
                          np.asarray(d.Delta20.coefficients), np.asarray(d.Delta11.coefficients)])
        return eng.tmunu_out(table, self._msq(fields), float(velocityMid))

    # ------------------------------------------------------------------ batches of independent systems
    def _siblings(self, count):
        """This solver plus (count-1) shallow clones, each with its own engine, operator buffer and CUDA
        stream; grid, particles and collision tensor are shared (read-only while a batch runs)."""
        import copy

        sibs = getattr(self, "_sibs", None)
        if sibs is None or len(sibs) < count:
            sibs = [self] if sibs is None else sibs
            while len(sibs) < count:
                c = copy.copy(self)
                c._engine, c._staticKey, c._sibs, c._stream, c._batchEng = None, None, None, None, None
                sibs.append(c)
            self._sibs = sibs
        for s_ in sibs[:count]:
            s_.basisM, s_.basisN, s_.derivatives = self.basisM, self.basisN, self.derivatives
            s_.collisionMultiplier, s_.truncationOption = self.collisionMultiplier, self.truncationOption
            s_.refinementSteps = getattr(self, "refinementSteps", 1)
            s_.collisionArray, s_.offEqParticles, s_.grid = self.collisionArray, self.offEqParticles, self.grid
        return sibs[:count]

    @staticmethod
    def autoConcurrency(n: int, freeBytes: int | None = None) -> int:
        """How many independent systems of size n to keep in flight on one B200.  A single LU of this size
        cannot fill the GPU during its latency-bound panel chain (16 SMs), so a few large systems or many
        small ones share it.  Measured with tools/knob_sweep.py (round 2, solves/s through getDeltasBatch):
        n=15600: 2/3/4 slots 10.48/10.55/10.60; n=11600: 3/4/6/8 slots 23.7/24.0/24.2/24.3; n=8700: 4/6/8/12
        slots 50.6/51.9/51.9/52.3 -- flat beyond the first few, so the table only has to stay clear of too few."""
        c = 4 if n >= 12000 else 8 if n >= 2500 else 32
        if freeBytes is not None:
            perSlot = 8 * n * n + 2 * 8 * 256 * n + 64 * n + (1 << 20)
            c = max(1, min(c, int(0.6 * freeBytes) // perSlot))
        return c

    def _autoSlotCount(self) -> int:
        eng0 = self._getEngine()
        if getattr(self, "_autoSlots", None) is None or self._autoSlots[0] != eng0.n:
            free = eng0.torch.cuda.mem_get_info(eng0.dev)[0]      # ~30 ms: asked once per system size
            have = sum(8 * eng0.n * eng0.n for s_ in (getattr(self, "_sibs", None) or [self])
                       if s_._engine is not None and s_._engine.A is not None)
            self._autoSlots = (eng0.n, self.autoConcurrency(eng0.n, free + have))
        return self._autoSlots[1]

    def slotPool(self, concurrency: int | None = None, reuseFactors: bool = False):
        """Asynchronous form of getDeltas() for many independent systems: `concurrency` slots (None:
        autoConcurrency), each with its own engine, operator buffer, CUDA stream and captured graphs.
        submit() enqueues one system on a free slot without waiting; poll() advances the slots whose device
        work has finished (host half of the truncation, moments, result download) and returns the
        (token, BoltzmannResults) pairs that completed.  getDeltasBatch() and the wall-velocity scan server
        (wallgo_b200.vwscan) are built on it."""
        if concurrency is None or int(concurrency) <= 0:
            concurrency = self._autoSlotCount()
        return _SlotPool(self, max(1, int(concurrency)), reuseFactors)

    def getDeltasBatch(self, backgrounds, concurrency: int | None = None):
        """getDeltas() for a list of independent backgrounds (wall-velocity scan points, parameter points),
        `concurrency` systems in flight on this GPU (None: autoConcurrency).  Each system is solved exactly
        as getDeltas() would; every slot replays its factorisation as one CUDA graph on its own stream, and
        the slots are staggered so that the latency-bound end of one LU overlaps the GEMM-rich start of the
        next.  Returns the BoltzmannResults in input order."""
        import time

        nItems = len(backgrounds)
        if nItems == 0:
            return []
        if concurrency is None and nItems >= 2 and self._getEngine().n <= self.STRIDED_MAX_N:
            return self.getDeltasStrided(backgrounds)
        if concurrency is None or int(concurrency) <= 0:
            concurrency = self._autoSlotCount()
        C = max(1, min(int(concurrency), nItems))
        pool = self.slotPool(C)
        results = [None] * nItems
        n = self._getEngine().n
        stagger = (2.0 / 3.0) * float(n) ** 3 / 20e12 / C   # rough LU time / C
        nextItem = 0
        try:
            while nextItem < min(C, nItems):
                if nextItem > 0 and stagger > 1e-3:
                    time.sleep(stagger)
                pool.submit(nextItem, background=backgrounds[nextItem])
                nextItem += 1
            done = 0
            while done < nItems:
                finished = pool.poll()
                if not finished:
                    pool.waitOldest()
                    continue
                for token, res in finished:
                    results[token] = res
                    done += 1
                    if nextItem < nItems:
                        pool.submit(nextItem, background=backgrounds[nextItem])
                        nextItem += 1
        finally:
            pool.close()
        return results

    # ------------------------------------------------------------------ strided batches (many small systems)
    STRIDED_MAX_N = 4096     # systems up to this order go through the lock-step strided-batch kernels
    STRIDED_BATCH = 64       # systems per launch sequence

    def _batchEngines(self, count):
        from .engine import BatchEngine

        eng = self._getEngine()
        cur = getattr(self, "_batchEng", None)
        if cur is None or cur[0] is not eng or cur[1] != count:
            if cur is not None:
                for b in cur[2]:
                    b.close()
            cur = (eng, count, [BatchEngine(eng, count), BatchEngine(eng, count)])
            self._batchEng = cur
        return cur[2]

    def getDeltasStrided(self, backgrounds, batch: int | None = None):
        """getDeltas() for many independent backgrounds of a SMALL system (north star (2): batched LU): the
        systems are processed `batch` at a time by launch sequences whose kernels carry the system index as a
        grid dimension (wg_batch_*), two batches in flight so the host half of the truncation decision and the
        result wrapping of one overlap the device work of the other.  Each result is bit-identical to
        getDeltas() on the same background.  A short last batch is padded with copies of its last system."""
        nItems = len(backgrounds)
        if nItems == 0:
            return []
        eng = self._getEngine()
        free = eng.torch.cuda.mem_get_info(eng.dev)[0]
        cap = max(1, int(0.4 * free) // (2 * (8 * eng.n * eng.n + 2 * 8 * 256 * eng.n + 200 * eng.n)))
        B = max(1, min(int(batch) if batch else self.STRIDED_BATCH, nItems, cap))
        bes = self._batchEngines(B)
        dxidchi, _, _ = self.grid.getCompactificationDerivatives()
        roundtrip = self.truncationOption.name != "NONE"
        mult = float(self.collisionMultiplier)
        chunks = [list(range(i, min(i + B, nItems))) for i in range(0, nItems, B)]
        results = [None] * nItems
        state = {}       # chunk -> (event, keeps)

        def start(c):
            be = bes[c % 2]
            for slot in range(B):
                item = chunks[c][min(slot, len(chunks[c]) - 1)]
                if slot >= len(chunks[c]) and slot > 0:
                    be.inp_host[slot].copy_(be.inp_host[slot - 1])
                    continue
                bg = backgrounds[item]
                if type(bg).__name__ == "BoltzmannBackground" and type(bg).boostToPlasmaFrame.__qualname__ == \
                        "BoltzmannBackground.boostToPlasmaFrame":
                    # setBackground without the deep copy: the boost of containers.py:152-157 (same expression,
                    # same bits) applied to the two quantities it changes
                    vm = bg.velocityMid
                    vprof = np.asarray(bg.velocityProfile, dtype=np.float64)
                    v = (vprof - vm) / (1.0 - vprof * vm)
                    vw = (bg.velocityWall - vm) / (1.0 - bg.velocityWall * vm)
                    T, fields = bg.temperatureProfile, bg.fieldProfiles
                else:
                    self.setBackground(bg)
                    b2 = self.background
                    T, v, fields, vw = b2.temperatureProfile, b2.velocityProfile, b2.fieldProfiles, b2.velocityWall
                be.set_system(slot, np.asarray(T, dtype=np.float64), np.asarray(v, dtype=np.float64), self._msq(fields),
                              dxidchi, float(vw))
            state[c] = (be.solve_async(mult, bool(getattr(self, "refinementSteps", 1))), None)

        def decide(c):
            be = bes[c % 2]
            state[c][0].synchronize()
            sp = be.spectra_host.numpy()
            keeps = [self._decideTruncation(sp[slot].copy()) for slot in range(B)]
            state[c] = (be.moments_async([k.dev for k, _ in keeps], roundtrip), keeps)

        def finish(c):
            be = bes[c % 2]
            ev, keeps = state.pop(c)
            ev.synchronize()
            if any(be.infos()[:len(chunks[c])]):
                raise np.linalg.LinAlgError("Singular matrix")
            P, M1, N1 = eng.P, eng.M1, eng.N1
            dFh, dh, eh = be.dF_host.numpy(), be.deltas_host.numpy(), be.err_host.numpy()
            rs = be.refstats_host.numpy()
            self.lastRefinement = (float(rs[len(chunks[c]) - 1, 0]), float(rs[len(chunks[c]) - 1, 1])) \
                if getattr(self, "refinementSteps", 1) else None
            for slot, item in enumerate(chunks[c]):
                keep, peaks = keeps[slot]
                results[item] = self._wrapResults(dFh[slot].reshape(P, M1, N1, N1).copy(),
                                                  dh[slot].reshape(4, P, M1).copy(), eh[slot].reshape(P, 4).copy(),
                                                  keep, peaks)

        nC = len(chunks)
        for c in range(nC + 2):
            if c >= 2:
                finish(c - 2)
            if c < nC:
                start(c)
            if 1 <= c <= nC:
                decide(c - 1)
        self.setBackground(backgrounds[-1])     # what a sequential loop would leave behind
        return results

    def _wrapResults(self, dF, deltas, err, keep, peaks):
        g = self.grid
        M1, N1 = g.M - 1, g.N - 1
        truncationError = float(np.max(err[:, 1:] / err[:, :1]))
        truncatedTail = (keep[0] != M1, keep[1] != N1, keep[2] != N1)
        T = self._types
        polys = [T.Polynomial(deltas[i], g, ("Array", "Cardinal"), ("Array", "z"), (False, False))
                 for i in range(4)]
        Deltas = T.BoltzmannDeltas(Delta00=polys[0], Delta02=polys[1], Delta20=polys[2], Delta11=polys[3])
        return T.BoltzmannResults(deltaF=dF, Deltas=Deltas, truncationError=truncationError,
                                  truncatedTail=truncatedTail, spectralPeaks=peaks)

    def checkLinearization(self, deltaF=None):
        """The two linearisation criteria (boltzmann.py:500-626).  The second solve re-uses the LU
        factors of the first instead of re-assembling and re-factoring."""
        eng = self._solveOnDevice()          # operator factors + deltaF in eng.S
        torch = eng.torch
        shape = self._shape()
        if deltaF is not None:
            dFdev = self._toDevice(eng, deltaF).clone()
        else:
            dFdev = eng.S.clone()
        card = torch.empty_like(dFdev)
        eng.change_basis(2, dFdev, card)                     # solver basis -> Cardinal
        sq = card * card
        sqB = torch.empty_like(sq)
        # Cardinal -> solver basis = inverse of stage 2; identical to stage 0 on the axes that differ
        self._cardinalToSolver(eng, sq, sqB)
        src2 = torch.empty_like(sq)
        eng.collision_apply(sqB, src2)
        eng.solve(src2)                                      # deltaNonlin in src2
        nlCard = torch.empty_like(src2)
        eng.change_basis(2, src2, nlCard)
        # pressure integrals: small host-side quadrature on three n-vectors
        g = self.grid
        bg = self.background
        msqFull = self._msq(bg.fieldProfiles)
        tabs = eng.get_tables()
        dmsq = (msqFull @ tabs["DchiFull"].T)[:, 1:-1, None, None]
        chi, rz, rp = g.getCompactCoordinates()
        _, pz, pp = g.getCoordinates()
        _, dpzdrz, dppdrp = g.getCompactificationDerivatives()
        pz4, pp4 = pz[None, None, :, None], pp[None, None, None, :]
        energy = np.sqrt(msqFull[:, 1:-1, None, None] + pz4**2 + pp4**2)
        temp = np.asarray(bg.temperatureProfile)[None, 1:-1, None, None]
        stat = np.array([-1 if p.statistics == "Fermion" else 1 for p in self.offEqParticles])[:, None, None, None]
        x = energy / temp
        with np.errstate(over="ignore"):
            fEq = np.where(x > self.MAX_EXPONENT, 0.0, 1 / (np.exp(x) - stat))
        dofs = np.array([p.totalDOFs for p in self.offEqParticles])[:, None, None, None]
        integrand = dofs * dmsq * dpzdrz[None, None, :, None] * dppdrp[None, None, None, :] * pp4 / (
            4 * np.pi**2 * energy)
        wchi = np.sqrt(1 - chi**2) * (np.pi / g.M)
        wz = np.sqrt(1 - rz**2) * (np.pi / g.N)
        wp = np.full(g.N - 1, np.pi / (g.N - 1))
        wp[0] /= 2
        wp = np.sqrt(1 - rp**2) * wp
        W = wchi[None, :, None, None] * wz[None, None, :, None] * wp[None, None, None, :]
        pEq = np.sum(fEq * integrand * W)
        pDf = np.sum(card.cpu().numpy().reshape(shape) * integrand * W)
        pNl = np.sum(nlCard.cpu().numpy().reshape(shape) * integrand * W)
        return abs(pDf / pEq), abs(pNl / (pEq + pDf))

    def _cardinalToSolver(self, eng, src, dst):
        """Cardinal values -> solver basis.  On axes where the solver basis is Chebyshev this is the
        inverse Chebyshev matrix; such matrices are stage 0 only for Cardinal solver axes, so the
        (rare) Chebyshev case is done through a temporary transform set."""
        g = self.grid
        from .spectral import chebyshevMatrix, lobattoNodes
        chi, rz, rp = lobattoNodes(g.M, g.N)
        mats = basisChangeMatrices(g.M, g.N, self.basisM, self.basisN)
        inv = [None] * 9
        if self.basisM == "Chebyshev":
            inv[0] = np.linalg.inv(chebyshevMatrix(chi, g.M - 1, "full"))
        if self.basisN == "Chebyshev":
            inv[1] = np.linalg.inv(chebyshevMatrix(rz, g.N - 1, "full"))
            inv[2] = np.linalg.inv(chebyshevMatrix(rp, g.N - 1, "partial"))
        if all(m is None for m in inv):
            dst.copy_(src)
            return
        eng.set_transforms(inv[:3] + mats[3:])
        eng.change_basis(0, src, dst)
        eng.torch.cuda.current_stream(eng.dev).synchronize()
        eng.set_transforms(mats)

    # ------------------------------------------------------------------ validation helpers
    @staticmethod
    def _checkBasis(basis: str) -> None:
        assert basis in ["Cardinal", "Chebyshev"], f"BoltzmannSolver error: unkown basis {basis}"

    @staticmethod
    def _checkDerivatives(derivatives: str) -> None:
        assert derivatives in ["Spectral", "Finite Difference"], \
            f"BoltzmannSolver error: unkown derivatives option {derivatives}"


class BoltzmannSolver(BoltzmannSolverB200Mixin):
    """Standalone class (no WallGo import needed).  Constructor signature follows
    boltzmann.py:52-60 plus an optional CUDA device index."""

    def __init__(self, grid, basisM: str = "Cardinal", basisN: str = "Chebyshev",
                 derivatives: str = "Spectral", collisionMultiplier: float = 1.0,
                 truncationOption=ETruncationOption.AUTO, device: int | None = None):
        self._initB200(grid, basisM, basisN, derivatives, collisionMultiplier, truncationOption, device)
