import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if len(sys.argv) > 1:
    n, mode, B = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
    import numpy as np, torch
    from test_formats import _shipped_spec
    from engine_util import make_engine
    name = {2: "empty-8-8", 3: "room-32-32-4", 4: "room-32-32-4", 5: "empty-16-16", 8: "empty-8-8"}[n]
    eng = make_engine(_shipped_spec(name, 1, n, 0.2, -1000.0, 100.0, -1.0, True))
    rng = np.random.default_rng(1)
    cells = torch.from_numpy(rng.integers(0, eng.L, (B, n)).astype(np.int32)).cuda()
    st = eng.encode(cells)
    ac = torch.from_numpy(rng.integers(0, 5 ** n, B).astype(np.int32)).cuda()
    un = torch.from_numpy(rng.random((B, n))).cuda() if mode == "tape" else None
    a = eng.step(st, ac, uniforms=un, seed=3)
    torch.cuda.synchronize()
    b = eng.step(st, ac, uniforms=un, seed=3, mapping="lanes")
    torch.cuda.synchronize()
    print("n", n, mode, B, "equal:", [bool(torch.equal(x, y)) for x, y in zip(a, b)])
    if not all(torch.equal(x, y) for x, y in zip(a, b)):
        bad = (a[0] != b[0]).nonzero().flatten()[:5].tolist()
        print(" first mismatching envs", bad, "thread", a[0][bad].tolist(), "lanes", b[0][bad].tolist(), "cells", cells[bad].tolist())
else:
    for n in (2, 3, 4, 5, 8):
        for mode in ("tape", "philox"):
            for B in (1, 4099):
                try:
                    r = subprocess.run([sys.executable, __file__, str(n), mode, str(B)], capture_output=True, text=True,
                                       env=dict(os.environ, CUDA_LAUNCH_BLOCKING="1"), timeout=40)
                except subprocess.TimeoutExpired:
                    print("n %d %s %d TIMEOUT" % (n, mode, B), flush=True)
                    sys.exit(1)
                print(r.stdout.strip() or ("n %d %s %d FAILED: " % (n, mode, B)) + r.stderr.strip().splitlines()[-1][:200], flush=True)
