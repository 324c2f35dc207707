"""Randomised parity sweep: many (grid family, d, B, M, request, transformations, nvec) configurations at small sizes,
CUDA library vs CPU oracle — exact mode bit for bit, fast mode within TAU·S.  Complements tests/ (fixed cases) by
walking the dispatch table: full-size and 128-thread leaf variants of every depth, the generic nested kernel (M < B),
the general-∂ kernel, the block kernel and the per-vector path; and basis_at (chunked tables, TMA stores) bit for bit.
usage: python profiles/fuzz_parity.py [n_configs] [seed] > profiles/r1_fuzz_parity.jsonl"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
from oracle import oracle as orc
sk = ge.load_package()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 150
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
TAU = 1e-13
GR = {0: sk.InteriorGrid(), 1: sk.EndpointGrid(), 2: sk.InteriorGrid2()}
same = lambda a, b: a.shape == b.shape and bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())
bad, worst, paths = 0, 0.0, {}
t0 = time.time()
for it in range(N):
    code = int(rng.integers(0, 3))
    d = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 16]))
    Bmax = {0: 4, 1: 5, 2: 5}[code] if d <= 6 else (4 if d <= 8 else 3 if d <= 12 else 2)      # K <= 6000 filters below
    B = int(rng.integers(0, Bmax + 1))
    M = B if rng.random() < 0.8 else int(rng.integers(0, B + 1))
    if orc.smolyak_length(code, d, B, M) > 6000:
        continue
    # transformations: none, or a random mix (orders >= 2 only through BoundedLinear, the reference's behaviour)
    kind = rng.choice(["values", "gradient", "general", "block"], p=[0.2, 0.4, 0.2, 0.2])
    use_tr = rng.random() < 0.5
    trs_sk = trs_or = None
    if use_tr:
        ks = [int(rng.integers(1, 4)) if kind != "general" else 1 for _ in range(d)]
        par = [(float(rng.uniform(-2, 0)), float(rng.uniform(0.5, 3))) for _ in range(d)]
        mk = {1: (sk.BoundedLinear, orc.BoundedLinear), 2: (sk.SemiInfRational, orc.SemiInfRational), 3: (sk.InfRational, orc.InfRational)}
        trs_sk = [mk[k][0](a, b) for k, (a, b) in zip(ks, par)]
        trs_or = [mk[k][1](a, b) for k, (a, b) in zip(ks, par)]
    basis0 = sk.smolyak_basis(sk.Chebyshev, GR[code], sk.SmolyakParameters(B, M), d)
    basis = basis0 @ sk.coordinate_transformations(*trs_sk) if use_tr else basis0
    K = sk.dimension(basis0)
    n = int(rng.integers(1, 300))
    u = np.clip((rng.random((n, d)) - 0.5) * 2.2, -0.97, 0.97) if use_tr else np.clip((rng.random((n, d)) - 0.5) * 2.5, -1, 1)
    y = orc.transform_from_batch(trs_or, u) if use_tr else u
    spec = ospec = None
    parts = None
    nvec = 1
    if kind == "gradient":
        spec, ospec = sk.gradient(d), orc.gradient_spec(d)
    elif kind == "general":
        parts = []
        for _ in range(int(rng.integers(1, 4))):
            p = [0] * d
            for j in rng.choice(d, size=min(d, int(rng.integers(1, 3))), replace=False):
                p[j] = int(rng.integers(1, 3))
            while p and p[-1] == 0:
                p.pop()
            if p:
                parts.append(tuple(p))
        if not parts:
            parts = [(1,)]
        spec = sk.partial(*parts[0])
        for q in parts[1:]:
            spec = spec | sk.partial(*q)
        ospec = orc.deriv_spec(parts, d)
    elif kind == "block":
        nvec = int(rng.choice([2, 3, 5, 8, 17, 30, 40, 70, 130]))
    theta = rng.standard_normal(K if nvec == 1 else (K, nvec))
    if os.environ.get("SK_FUZZ_VERBOSE"):
        print(json.dumps({"it": it, "code": code, "d": d, "B": B, "M": M, "kind": str(kind), "nvec": nvec, "tr": bool(use_tr), "n": n}), file=sys.stderr, flush=True)
    pts = spec(y) if spec is not None else y
    want = orc.smolyak_lincomb(code, d, B, M, theta, y, trs=trs_or, spec=ospec)
    scale = orc.smolyak_lincomb(code, d, B, M, theta, y, trs=trs_or, spec=ospec, absmode=True)
    fp = int(sk.get_plan(basis, 0, spec).info.fast_path)
    paths[(kind, fp)] = paths.get((kind, fp), 0) + 1
    try:
        ex = np.asarray(sk.linear_combination(basis, theta, pts, exact=True)).reshape(want.shape)
    except sk.UnsupportedError:      # tables of the bit-identical direct kernel exceed shared memory (large ℓ(B)·d): fast mode only
        ex = want
        paths[("exact-unsupported", fp)] = paths.get(("exact-unsupported", fp), 0) + 1
    fa = np.asarray(sk.linear_combination(basis, theta, pts)).reshape(want.shape)
    ok_exact = same(ex, want)
    # basis_at / collocation_matrix (round 2: chunked-table kernel with TMA stores): bit for bit against the oracle
    ok_basis = True
    if kind != "block":
        m = min(n, 24)
        bw = orc.smolyak_basis_at(code, d, B, M, y[:m], trs=trs_or, spec=ospec)                 # (K, m, E)
        try:
            bg = np.asarray(sk.basis_at(basis, spec(y[:m]) if spec is not None else y[:m]))
            bg = bg[:, :, None] if bg.ndim == 2 else bg
            ok_basis = same(bg, bw.transpose(1, 0, 2))
        except Exception as ex_:      # a refused or failed launch is a failure of the sweep, reported with its configuration
            ok_basis = False
            print(json.dumps({"basis_at_error": str(ex_)[-160:], "E": int(bw.shape[2]), "parts": parts if kind == "general" else None}), flush=True)
        paths[("basis_at", int(bw.shape[2]))] = paths.get(("basis_at", int(bw.shape[2])), 0) + 1
    with np.errstate(all="ignore"):
        err = float(np.nanmax(np.where(scale > 0, np.abs(fa - want) / scale, np.where(fa == want, 0.0, np.inf)))) if fa.size else 0.0
    worst = max(worst, err)
    if not ok_exact or not ok_basis or not err <= TAU:
        bad += 1
        print(json.dumps({"FAIL": True, "code": code, "d": d, "B": B, "M": M, "kind": kind, "nvec": nvec, "tr": use_tr, "n": n, "fast_path": fp,
                          "exact_bitwise": ok_exact, "basis_at_bitwise": ok_basis, "fast_scaled_err": err}), flush=True)
print(json.dumps({"configs": N, "seed": seed, "failures": bad, "worst_fast_scaled_error": worst, "tolerance": TAU,
                  "paths (kind, fast_path)": {f"{k[0]}:{k[1]}": v for k, v in sorted(paths.items())}, "seconds": time.time() - t0}))
