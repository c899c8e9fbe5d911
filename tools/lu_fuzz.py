"""Developer fuzz: LU on structured random matrices vs LAPACK (scipy.linalg.lu_factor)."""
import ctypes as C, os, sys
import numpy as np, scipy.linalg, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from wallgo_b200 import _lib
lib = _lib.load()
ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rng0 = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)

def gpu_lu(A, lda):
    n = A.shape[0]
    Ah = np.zeros((n, lda)); Ah[:, :n] = A
    Ad = torch.from_numpy(Ah).cuda()
    ipiv = torch.zeros(n, dtype=torch.int32, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    plan = C.c_void_p(); _lib.check(lib.wg_lu_create(n, 0, C.byref(plan)))
    try:
        _lib.check(lib.wg_lu_factor(plan, Ad.data_ptr(), lda, ipiv.data_ptr(), info.data_ptr(), st()))
        torch.cuda.synchronize()
    finally:
        lib.wg_lu_destroy(plan)
    return Ad.cpu().numpy()[:, :n], ipiv.cpu().numpy(), int(info.item())

kinds = ["dense", "sparse", "lowrank", "duprows", "range", "integer", "zeroblock", "permdiag", "tinydiag", "zerocols", "nonfinite"]
big = len(sys.argv) > 3
bad = 0
for case in range(ncases):
    seed = int(rng0.integers(1 << 30)); rng = np.random.default_rng(seed)
    kind = kinds[case % len(kinds)]
    n = int(rng.choice([1, 2, 3, 7, 15, 16, 17, 31, 33, 64, 100, 255, 256, 257, 300, 511, 513, 700, 1023, 1500, 2100]))
    if big: n = int(rng.choice([4100, 5000, 8200, 8448, 9100]))
    lda = n + int(rng.integers(0, 4))
    A = rng.standard_normal((n, n))
    if kind == "sparse": A *= rng.random((n, n)) < 0.15
    elif kind == "lowrank" and n > 3: r = max(1, n // 3); A = rng.standard_normal((n, r)) @ rng.standard_normal((r, n))
    elif kind == "duprows" and n > 4: A[n // 2] = A[1]; A[n - 1] = -A[2]
    elif kind == "range": A *= 10.0 ** rng.uniform(-150, 150, size=(n, 1)) * 10.0 ** rng.uniform(-100, 100, size=(1, n)) * 1e-50
    elif kind == "integer": A = rng.integers(-2, 3, size=(n, n)).astype(float)
    elif kind == "zeroblock" and n > 4: A[: n // 2, : n // 2] = 0.0
    elif kind == "permdiag": A = np.diag(rng.uniform(0.5, 2, n))[rng.permutation(n)]
    elif kind == "tinydiag": A[np.arange(n), np.arange(n)] *= 1e-14
    elif kind == "zerocols" and n > 2: A[:, rng.integers(0, n, size=max(1, n // 50))] = 0.0
    elif kind == "nonfinite":
        for _ in range(int(rng.integers(1, 4))):
            A[rng.integers(0, n), rng.integers(0, n)] = rng.choice([np.inf, -np.inf, np.nan])
    print(f"case {case} kind={kind} n={n} lda={lda} seed={seed}", flush=True)
    LU, ipiv, info = gpu_lu(A, lda)
    if ipiv.min() < 0 or ipiv.max() >= n or np.any(ipiv < np.arange(n)):
        bad += 1; print("  BAD pivots out of range"); continue
    if kind == "nonfinite" or (big and kind not in ("dense", "tinydiag", "zerocols", "permdiag")):
        continue
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            LUr, pr = scipy.linalg.lu_factor(A, check_finite=False)
    zr = np.where(np.diag(LUr) == 0)[0]
    info_ref = int(zr[0]) + 1 if zr.size else 0
    same_piv = np.array_equal(ipiv, pr)
    # P A = L U whenever the factorisation ran through without an exact zero pivot
    ok_fact = True
    if info == 0 and np.all(np.isfinite(LU)):
        perm = np.arange(n)
        for i, pv in enumerate(ipiv): perm[i], perm[pv] = perm[pv], perm[i]
        L = np.tril(LU, -1) + np.eye(n); U = np.triu(LU)
        scale = np.abs(L) @ np.abs(U) + 1e-300
        ok_fact = bool(np.max(np.abs(L @ U - A[perm]) / scale.max()) < 1e-12) if kind != "range" else bool(np.max(np.abs(L @ U - A[perm]) / scale) < 1e-9)
        if np.max(np.abs(L)) > 1.0 + 1e-12: ok_fact = False
    structural = kind in ("zerocols", "permdiag", "zeroblock")
    if not ok_fact or (structural and info != info_ref) or (kind in ("dense", "permdiag", "tinydiag") and not same_piv):
        bad += 1
        print(f"  MISMATCH ok_fact={ok_fact} info={info} ref={info_ref} same_piv={same_piv}")
print(f"{ncases} cases, {bad} bad")
sys.exit(1 if bad else 0)
