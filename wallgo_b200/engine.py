"""Device engine: owns the torch tensors and drives libwallgo_b200.so through ctypes.

PyTorch is used for device memory, pinned staging buffers and streams only; all
arithmetic happens in the hand-written CUDA kernels behind the C ABI.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import check, host_ptr


def default_device() -> int:
    """One process per GPU: LOCAL_RANK selects the device under torchrun."""
    return int(os.environ.get("LOCAL_RANK", "0"))


class BoltzmannEngine:
    """One Boltzmann system (P species, M x N grid) resident on one B200."""

    def __init__(self, P: int, M: int, N: int, device: int | None = None):
        import torch  # imported lazily so that CPU-only tooling can import the package

        if not torch.cuda.is_available():
            raise _lib.WallGoB200Error(
                "wallgo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.torch = torch
        self.lib = _lib.load()
        self.P, self.M, self.N = int(P), int(M), int(N)
        self.device = default_device() if device is None else int(device)
        self.dev = torch.device("cuda", self.device)
        h = C.c_void_p()
        check(self.lib.wg_create(self.P, self.M, self.N, self.device, C.byref(h)), "wg_create")
        self.h = h
        n, q = C.c_int(), C.c_int()
        check(self.lib.wg_dims(self.h, C.byref(n), C.byref(q)), "wg_dims")
        self.n, self.q = n.value, q.value
        self.M1, self.N1 = self.M - 1, self.N - 1
        self.nprof = (2 + self.P) * (self.M + 1) + self.M1
        f64 = torch.float64
        self.A = None  # n x n operator / LU factors, allocated on first use
        self.S = torch.empty(self.n, dtype=f64, device=self.dev)          # source -> deltaF (in place)
        self.prof = torch.empty(self.nprof, dtype=f64, device=self.dev)
        self.prof_host = torch.empty(self.nprof, dtype=f64).pin_memory()
        self.nspec = self.M1 + 2 * self.N1
        self.spectra = torch.empty(self.nspec, dtype=f64, device=self.dev)
        self.spectra_host = torch.empty(self.nspec, dtype=f64).pin_memory()
        # [ deltaF_out (n) | deltas (4 P M1) | err (4 P) | refinement stats (2) ] contiguous so one D2H copy
        # returns everything
        self.nout = self.n + 4 * self.P * self.M1 + 4 * self.P + 2
        self.out = torch.empty(self.nout, dtype=f64, device=self.dev)
        self.out_host = torch.empty(self.nout, dtype=f64).pin_memory()
        self.dF_in = torch.empty(self.n, dtype=f64, device=self.dev)
        self.launches = 0  # kernels launched through this engine (rough, for bench accounting)
        self.stream_ptr = None
        # Factor re-use (BoltzmannSolver.slotPool(reuseFactors=True)): what the LU factors in self.A belong to, as set
        # by the caller that enqueued the factorisation; every assemble() / factor() of this engine clears it.
        self.factorTag = None

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.wg_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        # batch drivers pin the slot's stream once (stream_ptr) instead of asking torch on every call
        if self.stream_ptr is not None:
            return self.stream_ptr
        return C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def _onEngineStream(self):
        """Context in which torch copies / synchronisation use the same stream as the kernels: the pinned
        slot stream while a batch runs (stream_ptr), torch's current stream otherwise."""
        import contextlib

        if self.stream_ptr is None:
            return contextlib.nullcontext()
        return self.torch.cuda.stream(self.torch.cuda.ExternalStream(int(self.stream_ptr.value), device=self.dev))

    def _ensure_A(self):
        if self.A is None:
            self.A = self.torch.empty((self.n, self.n), dtype=self.torch.float64, device=self.dev)
        return self.A

    # ------------------------------------------------------------------ static data
    def set_grid(self, chi, rz, rp, pz, pp, dpzdrz, dppdrp):
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (chi, rz, rp, pz, pp, dpzdrz, dppdrp)]
        assert arrs[0].size == self.M1 and all(a.size == self.N1 for a in arrs[1:])
        check(self.lib.wg_set_grid(self.h, *[host_ptr(a) for a in arrs]), "wg_set_grid")

    def set_statistics(self, stat):
        s = np.ascontiguousarray(stat, dtype=np.float64)
        assert s.size == self.P
        check(self.lib.wg_set_statistics(self.h, host_ptr(s)), "wg_set_statistics")

    def set_dofs(self, dofs):
        d = np.ascontiguousarray(dofs, dtype=np.float64)
        assert d.size == self.P
        check(self.lib.wg_set_dofs(self.h, host_ptr(d)), "wg_set_dofs")

    def feedback(self, dmsq, velocityMid: float):
        """T30_out, T33_out, potentialOut [M-1] and dVout [M-1, F] from the device-resident Deltas of the
        last moments call (one tiny upload of dm2/dphi, one tiny download)."""
        torch = self.torch
        n, pm = self.n, 4 * self.P * self.M1
        F = 0 if dmsq is None else int(np.asarray(dmsq).shape[-1])
        out = torch.empty((3 + F) * self.M1, dtype=torch.float64, device=self.dev)
        dptr = 0
        if F:
            d = torch.from_numpy(np.ascontiguousarray(dmsq, dtype=np.float64).reshape(self.P, self.M1, F)).to(self.dev)
            dptr = d.data_ptr()
        with self._onEngineStream():
            check(self.lib.wg_feedback(self.h, self.out[n:].data_ptr(), self.prof.data_ptr(), dptr, F,
                                       float(velocityMid), out.data_ptr(), self._stream()), "wg_feedback")
            h = out.cpu().numpy()
        M1 = self.M1
        return h[:M1].copy(), h[M1:2 * M1].copy(), h[2 * M1:3 * M1].copy(), h[3 * M1:].reshape(M1, F).copy()

    def tmunu_out(self, deltas, msqInterior, velocityMid: float):
        """T30_out, T33_out [M-1] (EOM.deltaToTmunu for all interior z at once) from HOST moment tables: the wall
        EOM mixes the moments of successive solves on the host (under-relaxation, equationOfMotion.py:1352-1355)
        before it asks for the plasma profile, so the table is uploaded (4 P (M-1) doubles), wg_feedback runs
        once and 2 (M-1) doubles come back."""
        torch = self.torch
        pm, L, M1 = 4 * self.P * self.M1, self.M + 1, self.M1
        if getattr(self, "_fb_host", None) is None:
            self._fb_host = torch.zeros(pm + self.nprof, dtype=torch.float64).pin_memory()
            self._fb_dev = torch.empty(pm + self.nprof, dtype=torch.float64, device=self.dev)
            self._fb_out = torch.empty(3 * M1, dtype=torch.float64, device=self.dev)
            self._fb_out_host = torch.empty(3 * M1, dtype=torch.float64).pin_memory()
        h = self._fb_host.numpy()
        h[:pm] = np.asarray(deltas, dtype=np.float64).reshape(-1)
        msq = np.asarray(msqInterior, dtype=np.float64).reshape(self.P, M1)
        for a in range(self.P):
            h[pm + (2 + a) * L + 1: pm + (2 + a) * L + 1 + M1] = msq[a]
        with self._onEngineStream():
            self._fb_dev.copy_(self._fb_host, non_blocking=True)
            check(self.lib.wg_feedback(self.h, self._fb_dev.data_ptr(), self._fb_dev[pm:].data_ptr(), 0, 0,
                                       float(velocityMid), self._fb_out.data_ptr(), self._stream()), "wg_feedback")
            self._fb_out_host.copy_(self._fb_out, non_blocking=True)
            torch.cuda.current_stream(self.dev).synchronize()
        o = self._fb_out_host.numpy()
        return o[:M1].copy(), o[M1:2 * M1].copy()

    def build_basis(self):
        check(self.lib.wg_build_basis(self.h), "wg_build_basis")

    def set_tables(self, Tchi, Dchi, DchiFull, Trz, Drz, Trp):
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (Tchi, Dchi, DchiFull, Trz, Drz, Trp)]
        check(self.lib.wg_set_tables(self.h, *[host_ptr(a) for a in arrs]), "wg_set_tables")

    def get_tables(self):
        M1, N1, L = self.M1, self.N1, self.M + 1
        out = [np.empty((M1, M1)), np.empty((M1, M1)), np.empty((L, L)), np.empty((N1, N1)),
               np.empty((N1, N1)), np.empty((N1, N1))]
        check(self.lib.wg_get_tables(self.h, *[host_ptr(a) for a in out]), "wg_get_tables")
        return dict(zip(("Tchi", "Dchi", "DchiFull", "Trz", "Drz", "Trp"), out))

    def set_transforms(self, mats):
        """mats: 9 entries (stage-major: toChebyshev, fromChebyshev, toCardinal; axis z, pz, pp) or None."""
        keep = [None if m is None else np.ascontiguousarray(m, dtype=np.float64) for m in mats]
        arr = (C.POINTER(C.c_double) * 9)()
        for i, m in enumerate(keep):
            if m is not None:
                size = self.M1 if i % 3 == 0 else self.N1
                assert m.shape == (size, size)
                arr[i] = host_ptr(m)
        check(self.lib.wg_set_transforms(self.h, arr), "wg_set_transforms")

    def set_collision(self, coll):
        c = np.ascontiguousarray(coll, dtype=np.float64)
        assert c.size == (self.P * self.q) ** 2, "collision tensor must be (P,N-1,N-1,P,N-1,N-1)"
        check(self.lib.wg_set_collision(self.h, C.c_void_p(c.ctypes.data), 0), "wg_set_collision")

    # ------------------------------------------------------------------ per solve
    def upload_profiles(self, T, v, msq, dxidchi):
        """Host -> device copy of the O(P*M) background numbers (one pinned buffer, one copy)."""
        L = self.M + 1
        ph = self.prof_host.numpy()
        ph[0:L] = T
        ph[L:2 * L] = v
        ph[2 * L:(2 + self.P) * L] = np.asarray(msq, dtype=np.float64).reshape(-1)
        ph[(2 + self.P) * L:] = dxidchi
        with self._onEngineStream():
            self.prof.copy_(self.prof_host, non_blocking=True)

    def assemble(self, vw: float, collisionMultiplier: float = 1.0, parts: int = 3, A=None):
        if A is None:
            self.factorTag = None      # the operator buffer no longer holds the factors of an earlier system
        A = self._ensure_A() if A is None else A
        check(self.lib.wg_assemble(self.h, self.prof.data_ptr(), float(vw), float(collisionMultiplier),
                                   int(parts), A.data_ptr(), A.stride(0), self.S.data_ptr(), self._stream()),
              "wg_assemble")
        return A

    def factor(self):
        A = self._ensure_A()
        self.factorTag = None
        check(self.lib.wg_factor(self.h, A.data_ptr(), A.stride(0), self._stream()), "wg_factor")

    def factor_solve(self, b=None):
        """np.linalg.solve in one call: LU with the forward solve of b riding along, then the backward solve."""
        b = self.S if b is None else b
        A = self._ensure_A()
        self.factorTag = None
        check(self.lib.wg_factor_solve(self.h, A.data_ptr(), A.stride(0), b.data_ptr(), self._stream()), "wg_factor_solve")
        return b

    def solve(self, b=None):
        """Triangular solves with the current factors; b (device, n) is overwritten by the solution."""
        b = self.S if b is None else b
        A = self._ensure_A()
        check(self.lib.wg_solve(self.h, A.data_ptr(), A.stride(0), b.data_ptr(), self._stream()), "wg_solve")
        return b

    def refine(self, vw: float, collisionMultiplier: float = 1.0, x=None):
        """One step of mixed-precision iterative refinement of x (default: the solution in self.S) with the
        current factors (wg_refine); max|correction| and max|x| land in the last two entries of self.out."""
        x = self.S if x is None else x
        A = self._ensure_A()
        check(self.lib.wg_refine(self.h, self.prof.data_ptr(), float(vw), float(collisionMultiplier), A.data_ptr(),
                                 A.stride(0), x.data_ptr(), self.out[self.nout - 2:].data_ptr(), self._stream()),
              "wg_refine")
        return x

    MAX_REUSE_STEPS = 64      # rows of the stale-factor refinement history (refine_steps)

    def refine_steps(self, vw: float, collisionMultiplier: float, first: int, count: int):
        """`count` further refinement steps of self.S against the system of the CURRENT profiles with whatever
        factors self.A holds (wg_refine: residual of the regenerated operator in double-double, triangular solves,
        update) -- the factors may belong to an earlier, nearby system (factor re-use across the pressure
        iteration of one scan point, BoltzmannSolver.slotPool(reuseFactors=True)): the iteration is then a
        Richardson iteration preconditioned by the stale factors and converges to the solution of the CURRENT
        system.  Step first+i writes (max|correction|, max|x|) to row first+i of self.rstats; the whole history
        is copied to pinned host memory (self.rstats_host) behind the last step."""
        torch = self.torch
        if getattr(self, "rstats", None) is None:
            self.rstats = torch.zeros((self.MAX_REUSE_STEPS, 2), dtype=torch.float64, device=self.dev)
            self.rstats_host = torch.zeros((self.MAX_REUSE_STEPS, 2), dtype=torch.float64).pin_memory()
        assert 0 <= first and first + count <= self.MAX_REUSE_STEPS
        A = self._ensure_A()
        for i in range(first, first + count):
            check(self.lib.wg_refine(self.h, self.prof.data_ptr(), float(vw), float(collisionMultiplier), A.data_ptr(),
                                     A.stride(0), self.S.data_ptr(), self.rstats[i].data_ptr(), self._stream()),
                  "wg_refine")
        with self._onEngineStream():
            self.rstats_host.copy_(self.rstats, non_blocking=True)

    def refine_history(self, steps: int):
        """(max|correction|, max|x|) of the first `steps` refinement steps of refine_steps(), after a synchronise."""
        with self._onEngineStream():
            self.torch.cuda.current_stream(self.dev).synchronize()
        return self.rstats_host.numpy()[:steps].copy()

    def info(self) -> int:
        v = C.c_int(0)
        check(self.lib.wg_info(self.h, self._stream(), C.byref(v)), "wg_info")
        return v.value

    def spectra_async(self, dF):
        check(self.lib.wg_spectra(self.h, dF.data_ptr(), self.spectra.data_ptr(), self._stream()), "wg_spectra")

    def spectra_to_host(self):
        with self._onEngineStream():
            self.spectra_host.copy_(self.spectra, non_blocking=True)
            self.torch.cuda.current_stream(self.dev).synchronize()
        return self.spectra_host.numpy().copy()

    def moments_async(self, dF, keep, roundtrip: bool):
        n, pm = self.n, 4 * self.P * self.M1
        o = self.out
        check(self.lib.wg_moments(self.h, dF.data_ptr(), self.prof.data_ptr(), int(keep[0]), int(keep[1]),
                                  int(keep[2]), 1 if roundtrip else 0, o.data_ptr(),
                                  o[n:].data_ptr(), o[n + pm:].data_ptr(), self._stream()), "wg_moments")

    def results_to_host(self):
        """One device -> host copy: (deltaF (P,M-1,N-1,N-1), Deltas (4,P,M-1), err (P,4))."""
        with self._onEngineStream():
            self.out_host.copy_(self.out, non_blocking=True)
            self.torch.cuda.current_stream(self.dev).synchronize()
        h = self.out_host.numpy()
        n, pm = self.n, 4 * self.P * self.M1
        dF = h[:n].reshape(self.P, self.M1, self.N1, self.N1).copy()
        deltas = h[n:n + pm].reshape(4, self.P, self.M1).copy()
        err = h[n + pm:n + pm + 4 * self.P].reshape(self.P, 4).copy()
        self.last_refinement = (float(h[-2]), float(h[-1]))   # (max|correction|, max|x|) of the last refine()
        return dF, deltas, err

    def collision_apply(self, x, y):
        check(self.lib.wg_collision_apply(self.h, x.data_ptr(), y.data_ptr(), self._stream()),
              "wg_collision_apply")

    def change_basis(self, stage: int, src, dst):
        check(self.lib.wg_change_basis(self.h, int(stage), src.data_ptr(), dst.data_ptr(), self._stream()),
              "wg_change_basis")

    # ------------------------------------------------------------------ batched form of the C ABI
    @staticmethod
    def solve_batched(engines, vws, collisionMultipliers, streams):
        """assemble + LU + solve of several engines' current backgrounds with one C call
        (wg_assemble_factor_solve_batched); engine b is enqueued on streams[b] (torch.cuda.Stream)."""
        B = len(engines)
        if B == 0:
            return
        lib = engines[0].lib
        vp = C.c_void_p * B
        hs = vp(*[e.h for e in engines])
        profs = vp(*[e.prof.data_ptr() for e in engines])
        As = vp(*[e._ensure_A().data_ptr() for e in engines])
        Ss = vp(*[e.S.data_ptr() for e in engines])
        sts = vp(*[st.cuda_stream for st in streams])
        ldas = (C.c_longlong * B)(*[e.A.stride(0) for e in engines])
        vw = (C.c_double * B)(*[float(v) for v in vws])
        cm = (C.c_double * B)(*[float(c) for c in collisionMultipliers])
        check(lib.wg_assemble_factor_solve_batched(B, hs, profs, vw, cm, As, ldas, Ss, sts),
              "wg_assemble_factor_solve_batched")

    # ------------------------------------------------------------------ fused device step (bench)
    def step_device(self, vw: float, collisionMultiplier: float, keep, roundtrip: bool = True, refine: bool = True):
        """assemble -> LU -> solve -> refinement step -> spectra -> truncate/moments, inputs already in HBM, no
        host sync."""
        self.assemble(vw, collisionMultiplier)
        self.factor()
        self.solve()
        if refine:
            self.refine(vw, collisionMultiplier)
        self.spectra_async(self.S)
        self.moments_async(self.S, keep, roundtrip)


class BatchEngine:
    """`count` systems of one engine's shape advanced in LOCK STEP by every kernel launch (wg_batch_*): the
    system index is a grid dimension, so one launch sequence / one CUDA graph serves the whole batch.  Shares
    the parent's static tables and collision tensor; owns operators, vectors and pinned staging for `count`
    systems and its own stream.  Per-system results are bit-identical to BoltzmannEngine's."""

    def __init__(self, parent: BoltzmannEngine, count: int):
        torch = parent.torch
        self.torch, self.lib, self.parent, self.count = torch, parent.lib, parent, int(count)
        self.dev = parent.dev
        h = C.c_void_p()
        check(self.lib.wg_batch_create(parent.h, self.count, C.byref(h)), "wg_batch_create")
        self.h = h
        B, n, f64 = self.count, parent.n, torch.float64
        self.n, self.nprof, self.nspec, self.nout = n, parent.nprof, parent.nspec, parent.nout
        self.strideA = n * n + (-(n * n)) % 4
        self.A = torch.empty(B * self.strideA, dtype=f64, device=self.dev)
        self.S = torch.empty((B, n), dtype=f64, device=self.dev)
        # [ profiles | vw ] per system in one pinned block -> one H2D copy per batch
        self.inp_host = torch.zeros((B, self.nprof + 1), dtype=f64).pin_memory()
        self.inp = torch.empty((B, self.nprof + 1), dtype=f64, device=self.dev)
        self.prof = torch.empty((B, self.nprof), dtype=f64, device=self.dev)
        self.vw = torch.empty(B, dtype=f64, device=self.dev)
        self.spectra = torch.empty((B, self.nspec), dtype=f64, device=self.dev)
        self.spectra_host = torch.empty((B, self.nspec), dtype=f64).pin_memory()
        self.keep_host = torch.zeros((B, 3), dtype=torch.int32).pin_memory()
        self.keep = torch.zeros((B, 3), dtype=torch.int32, device=self.dev)
        pm, pe = 4 * parent.P * parent.M1, 4 * parent.P
        self.dF = torch.empty((B, n), dtype=f64, device=self.dev)
        self.deltas = torch.empty((B, pm), dtype=f64, device=self.dev)
        self.err = torch.empty((B, pe), dtype=f64, device=self.dev)
        self.dF_host = torch.empty((B, n), dtype=f64).pin_memory()
        self.deltas_host = torch.empty((B, pm), dtype=f64).pin_memory()
        self.err_host = torch.empty((B, pe), dtype=f64).pin_memory()
        self.refstats = torch.zeros((B, 2), dtype=f64, device=self.dev)
        self.refstats_host = torch.zeros((B, 2), dtype=f64).pin_memory()
        self.info_host = (C.c_int * B)()
        self.stream = torch.cuda.Stream(device=self.dev)
        self._sp = C.c_void_p(self.stream.cuda_stream)

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.wg_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_system(self, i, T, v, msq, dxidchi, vw):
        """Host staging of system i's background (pinned memory; nothing is enqueued)."""
        L = self.parent.M + 1
        P = self.parent.P
        row = self.inp_host.numpy()[i]
        row[0:L] = T
        row[L:2 * L] = v
        row[2 * L:(2 + P) * L] = np.asarray(msq, dtype=np.float64).reshape(-1)
        row[(2 + P) * L:self.nprof] = dxidchi
        row[self.nprof] = vw

    def solve_async(self, collisionMultiplier: float, refine: bool = True):
        """One H2D copy, then assemble + LU + triangular solves (+ one refinement step) + spectra of all systems,
        spectra D2H; returns an event that fires when the spectra are on the host."""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            self.inp.copy_(self.inp_host, non_blocking=True)
            self.prof.copy_(self.inp[:, :self.nprof])
            self.vw.copy_(self.inp[:, self.nprof])
            n, sA = self.n, self.strideA
            check(self.lib.wg_batch_assemble(self.h, self.prof.data_ptr(), self.vw.data_ptr(), float(collisionMultiplier),
                                             3, self.A.data_ptr(), n, sA, self.S.data_ptr(), self._sp), "wg_batch_assemble")
            check(self.lib.wg_batch_factor(self.h, self.A.data_ptr(), n, sA, self._sp), "wg_batch_factor")
            check(self.lib.wg_batch_solve(self.h, self.A.data_ptr(), n, sA, self.S.data_ptr(), self._sp), "wg_batch_solve")
            if refine:
                check(self.lib.wg_batch_refine(self.h, self.prof.data_ptr(), self.vw.data_ptr(), float(collisionMultiplier),
                                               self.A.data_ptr(), n, sA, self.S.data_ptr(), self.refstats.data_ptr(),
                                               self._sp), "wg_batch_refine")
            check(self.lib.wg_batch_spectra(self.h, self.S.data_ptr(), self.spectra.data_ptr(), self._sp), "wg_batch_spectra")
            self.spectra_host.copy_(self.spectra, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return ev

    def moments_async(self, keeps, roundtrip: bool):
        """keeps: per system the three kept-coefficient counts; truncation + back-transform + moments of all
        systems, results D2H; returns an event."""
        torch = self.torch
        kh = self.keep_host.numpy()
        kh[:len(keeps)] = np.asarray(keeps, dtype=np.int32).reshape(-1, 3)
        with torch.cuda.stream(self.stream):
            self.keep.copy_(self.keep_host, non_blocking=True)
            check(self.lib.wg_batch_moments(self.h, self.S.data_ptr(), self.prof.data_ptr(), self.keep.data_ptr(),
                                            1 if roundtrip else 0, self.dF.data_ptr(), self.deltas.data_ptr(),
                                            self.err.data_ptr(), self._sp), "wg_batch_moments")
            self.dF_host.copy_(self.dF, non_blocking=True)
            self.deltas_host.copy_(self.deltas, non_blocking=True)
            self.err_host.copy_(self.err, non_blocking=True)
            self.refstats_host.copy_(self.refstats, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return ev

    def infos(self):
        check(self.lib.wg_batch_info(self.h, self._sp, self.info_host), "wg_batch_info")
        return list(self.info_host)
