"""Developer probe: LU on matrices containing NaN / inf must not fault (separate process per case)."""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    import torch
    from wallgo_b200 import _lib
    kind, n, graph = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    lib = _lib.load(); lib.wg_set_tuning(256, graph)
    rng = np.random.default_rng(0)
    A = rng.standard_normal((n, n))
    if kind == "allnan": A[:] = np.nan
    elif kind == "onenan": A[n // 2, n // 3] = np.nan
    elif kind == "oneinf": A[n // 2, n // 3] = np.inf
    elif kind == "nanrow": A[n // 2, :] = np.nan
    elif kind == "nancol": A[:, n // 3] = np.nan
    elif kind == "clean": pass
    elif kind.startswith("val:"):
        _, x, r, c = kind.split(":"); A[int(r), int(c)] = float(x)
    Ad = torch.from_numpy(A).cuda()
    ipiv = torch.zeros(n, dtype=torch.int32, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    plan = C.c_void_p(); _lib.check(lib.wg_lu_create(n, 0, C.byref(plan)))
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    try:
        _lib.check(lib.wg_lu_factor(plan, Ad.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), st))
        torch.cuda.synchronize()
    finally:
        out = (C.c_int * 56)()
        rc = lib.wg_debug_leaf_fail(out)
        print("leaf_fail rc", rc, list(out), flush=True)
    p = ipiv.cpu().numpy()
    print("ok", kind, n, "info", int(info.item()), "ipiv range", int(p.min()), int(p.max()))
else:
    cases = [("val:inf:150:100", 300), ("val:inf:40:20", 48), ("val:inf:100:20", 128), ("val:-inf:150:250", 300),
             ("allnan", 300), ("nanrow", 1900), ("nancol", 1900), ("onenan", 1900), ("oneinf", 1900), ("clean", 300)]
    for kind, n in cases:
        if True:
            r = subprocess.run([sys.executable, __file__, kind, str(n), "0"], capture_output=True, text=True, timeout=120,
                               env=dict(os.environ, CUDA_LAUNCH_BLOCKING="1"))
            print([l for l in r.stdout.splitlines() if l.startswith("leaf_fail")])
            line = [l for l in r.stdout.splitlines() if l.startswith("ok")]
            err = [l for l in r.stderr.splitlines() if "wg_lu.cu" in l or "illegal" in l][:1]
            print(kind, n, line[0] if line else "FAULT", err)
