"""Development tool (not part of the product): numpy/scipy prototype of the aggregation-multigrid preconditioner used
to choose cycle / smoother / over-correction before writing mg.cu.  Reports PCG iteration counts on the cfg-5
hydrostatic problem.  Usage: python tools/mg_prototype.py N"""
import sys, time
import numpy as np
import scipy.sparse as sp

WATER, AIR, SOLID = 0, 1, 2

def build(n, kind="dam"):
    t = np.full((n, n, n), AIR, np.uint8)
    t[0] = t[-1] = SOLID; t[:, 0] = SOLID; t[:, :, 0] = t[:, :, -1] = SOLID
    if kind == "dam":
        t[1:n // 2, 1:n - 1, 1:n - 1] = WATER
    elif kind == "splash":
        rng = np.random.default_rng(1)
        x, y, z = np.meshgrid(*[np.arange(n)] * 3, indexing="ij")
        h = n * (0.3 + 0.15 * np.sin(x * 6.0 / n) * np.cos(z * 4.0 / n))
        m = (y < h)
        m |= ((x - 0.6 * n) ** 2 + (y - 0.7 * n) ** 2 + (z - 0.5 * n) ** 2 < (0.12 * n) ** 2)
        t[1:-1, 1:-1, 1:-1][m[1:-1, 1:-1, 1:-1]] = WATER
        # a solid pillar
        t[int(.3*n):int(.4*n), 1:int(.6*n), int(.3*n):int(.5*n)] = SOLID
    return t

def assemble(t):
    n = t.shape
    idx = -np.ones(n, np.int64)
    w = t == WATER
    idx[w] = np.arange(w.sum())
    N = int(w.sum())
    diag = np.zeros(N)
    rows, cols, vals = [], [], []
    for ax in range(3):
        for sgn in (-1, 1):
            nb_t = np.roll(t, -sgn, axis=ax)
            nb_i = np.roll(idx, -sgn, axis=ax)
            diag += (nb_t != SOLID)[w]
            m = w & (nb_t == WATER)
            rows.append(idx[m]); cols.append(nb_i[m]); vals.append(-np.ones(m.sum()))
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N)).tocsr()
    A = A + sp.diags(diag)
    return A, idx

def hierarchy(t, A, idx, min_cells=64):
    levels = [dict(A=A, shape=t.shape, idx=idx)]
    while levels[-1]["A"].shape[0] > min_cells:
        L = levels[-1]
        shp = L["shape"]
        cshape = tuple((s + 1) // 2 for s in shp)
        fi = L["idx"]
        coords = np.argwhere(fi >= 0)
        cc = coords // 2
        key = (cc[:, 0] * cshape[1] + cc[:, 1]) * cshape[2] + cc[:, 2]
        uk, inv = np.unique(key, return_inverse=True)
        P = sp.coo_matrix((np.ones(len(inv)), (fi[fi >= 0][np.argsort(np.argsort(fi[fi>=0]))*0 + np.arange(len(inv))] if False else fi[tuple(coords.T)], inv)), shape=(L["A"].shape[0], len(uk))).tocsr()
        Ac = (P.T @ L["A"] @ P).tocsr()
        cidx = -np.ones(cshape, np.int64)
        cidx.ravel()[uk] = np.arange(len(uk))
        L["P"] = P
        levels.append(dict(A=Ac, shape=cshape, idx=cidx))
    for L in levels:
        L["D"] = L["A"].diagonal()
        c = np.argwhere(L["idx"] >= 0)
        col = np.zeros(L["A"].shape[0], bool)
        col[L["idx"][tuple(c.T)]] = (c.sum(1) % 2 == 0)
        L["red"] = col
    return levels

def smooth(L, x, b, kind, sweeps, reverse=False, omega=0.8):
    A, D = L["A"], L["D"]
    for _ in range(sweeps):
        if kind == "jacobi":
            x = x + omega * (b - A @ x) / D
        else:  # red-black GS
            order = (L["red"], ~L["red"]) if not reverse else (~L["red"], L["red"])
            for m in order:
                r = b - A @ x
                x = x.copy(); x[m] += r[m] / D[m]
    return x

def vcycle(levels, l, b, cfg):
    L = levels[l]
    if l == len(levels) - 1:
        x = np.zeros_like(b)
        return smooth(L, x, b, "rbgs", 10)  # coarsest: a few sweeps
    x = smooth(L, np.zeros_like(b), b, cfg["smoother"], cfg["pre"])
    r = b - L["A"] @ x
    rc = L["P"].T @ r
    ec = vcycle(levels, l + 1, rc, cfg)
    if cfg.get("w") and l + 1 < len(levels) - 1:
        rc2 = rc - levels[l + 1]["A"] @ ec
        ec = ec + vcycle(levels, l + 1, rc2, cfg)
    x = x + cfg["over"] * (L["P"] @ ec)
    x = smooth(L, x, b, cfg["smoother"], cfg["post"], reverse=True)
    return x

def pcg(A, b, M, tol, maxit=500):
    x = np.zeros_like(b); r = b.copy(); z = M(r); s = z.copy(); sigma = z @ r
    for it in range(maxit):
        q = A @ s; alpha = sigma / (s @ q)
        x += alpha * s; r -= alpha * q
        if np.abs(r).max() < tol: return x, it
        z = M(r); sn = z @ r; s = z + (sn / sigma) * s; sigma = sn
    return x, maxit

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    kind = sys.argv[2] if len(sys.argv) > 2 else "dam"
    t = build(n, kind)
    A, idx = assemble(t)
    scale = 0.005
    A = A * scale
    # hydrostatic RHS: -div of v2 = g*dt on y faces where neither cell solid
    gdt = -39.24 * 0.005
    vy = np.zeros(t.shape); up = np.roll(t, -1, axis=1)
    vy[(t != SOLID) & (up != SOLID)] = gdt
    div = vy - np.roll(vy, 1, axis=1)
    b = -div[t == WATER]
    rng = np.random.default_rng(0)
    b2 = b + rng.normal(0, 1.0, b.shape)   # FLIP-like noisy rhs (density drift term)
    t0 = time.time(); levels = hierarchy(t, A, idx); print("levels", [L["A"].shape[0] for L in levels], "build %.1fs" % (time.time() - t0))
    for name, rhs in (("hydro", b), ("noisy", b2)):
        x, it = pcg(A, rhs, lambda r: r / A.diagonal(), 1e-6, 3000); print(name, "jacobi its", it)
        for cfg in [dict(smoother="jacobi", pre=2, post=2, over=1.0), dict(smoother="jacobi", pre=2, post=2, over=1.8),
                    dict(smoother="rbgs", pre=1, post=1, over=1.0), dict(smoother="rbgs", pre=1, post=1, over=1.8),
                    dict(smoother="rbgs", pre=1, post=1, over=2.0), dict(smoother="rbgs", pre=2, post=2, over=1.8),
                    dict(smoother="rbgs", pre=1, post=1, over=1.8, w=True), dict(smoother="rbgs", pre=1, post=1, over=1.5, w=True)]:
            t0 = time.time()
            x, it = pcg(A, rhs, lambda r: vcycle(levels, 0, r, cfg), 1e-6)
            print(name, cfg, "its", it, "%.1fs" % (time.time() - t0), "true res", np.abs(rhs - A @ x).max())

def sweep_configs(n, kind="dam"):
    t = build(n, kind); A, idx = assemble(t); A = A * 0.005
    gdt = -39.24 * 0.005
    vy = np.zeros(t.shape); up = np.roll(t, -1, axis=1)
    vy[(t != SOLID) & (up != SOLID)] = gdt
    b = -(vy - np.roll(vy, 1, axis=1))[t == WATER]
    b = b + np.random.default_rng(0).normal(0, 1.0, b.shape)
    levels = hierarchy(t, A, idx)
    print(n, kind, "levels", [L["A"].shape[0] for L in levels])
    for cfg in [dict(smoother="jacobi", pre=1, post=1, over=1.8, omega=0.8), dict(smoother="jacobi", pre=2, post=2, over=1.8, omega=0.8),
                dict(smoother="jacobi", pre=2, post=2, over=1.8, omega=0.67), dict(smoother="jacobi", pre=2, post=2, over=1.6, omega=0.8),
                dict(smoother="jacobi", pre=3, post=3, over=1.8, omega=0.8), dict(smoother="jacobi", pre=2, post=2, over=1.8, omega=0.8, w=True),
                dict(smoother="rbgs", pre=1, post=1, over=1.8, w=True)]:
        global _omega
        om = cfg.get("omega", 0.8)
        import functools
        sm = smooth
        def M(r, cfg=cfg, om=om):
            return vcycle_om(levels, 0, r, cfg, om)
        t0 = time.time(); x, it = pcg(A, b, M, 1e-6); print(cfg, "its", it, "%.1fs" % (time.time() - t0))

def vcycle_om(levels, l, b, cfg, om):
    L = levels[l]
    if l == len(levels) - 1:
        return smooth(L, np.zeros_like(b), b, "rbgs", 10)
    x = smooth(L, np.zeros_like(b), b, cfg["smoother"], cfg["pre"], omega=om)
    rc = L["P"].T @ (b - L["A"] @ x)
    ec = vcycle_om(levels, l + 1, rc, cfg, om)
    if cfg.get("w") and l + 1 < len(levels) - 1:
        ec = ec + vcycle_om(levels, l + 1, rc - levels[l + 1]["A"] @ ec, cfg, om)
    x = x + cfg["over"] * (L["P"] @ ec)
    return smooth(L, x, b, cfg["smoother"], cfg["post"], reverse=True, omega=om)

def vcycle2(levels, l, b, cfg):
    """V / partial-W cycle with Jacobi smoothing and a Jacobi-iterated coarsest level (what mg.cu implements)."""
    L = levels[l]
    om = cfg["omega"]
    if l == len(levels) - 1:
        return smooth(L, np.zeros_like(b), b, "jacobi", cfg["coarse_sweeps"], omega=om)
    x = smooth(L, np.zeros_like(b), b, "jacobi", cfg["pre"], omega=om)
    rc = L["P"].T @ (b - L["A"] @ x)
    ec = vcycle2(levels, l + 1, rc, cfg)
    if l + 1 <= cfg.get("wlevels", 0) and l + 1 < len(levels) - 1:
        ec = ec + vcycle2(levels, l + 1, rc - levels[l + 1]["A"] @ ec, cfg)
    x = x + cfg["over"] * (L["P"] @ ec)
    return smooth(L, x, b, "jacobi", cfg["post"], omega=om)

def sweep2(n, kind="dam", min_cells=600):
    t = build(n, kind); A, idx = assemble(t); A = A * 0.005
    gdt = -39.24 * 0.005
    vy = np.zeros(t.shape); up = np.roll(t, -1, axis=1)
    vy[(t != SOLID) & (up != SOLID)] = gdt
    b = -(vy - np.roll(vy, 1, axis=1))[t == WATER]
    b = b + np.random.default_rng(0).normal(0, 1.0, b.shape)
    levels = hierarchy(t, A, idx, min_cells=min_cells)
    print(n, kind, "levels", [L["A"].shape[0] for L in levels])
    base = dict(pre=2, post=2, over=1.8, omega=0.8, coarse_sweeps=30)
    for extra in [dict(), dict(wlevels=1), dict(wlevels=2), dict(wlevels=3), dict(wlevels=9), dict(pre=3, post=3), dict(pre=3, post=3, wlevels=2),
                  dict(over=2.0, wlevels=2), dict(coarse_sweeps=100, wlevels=2)]:
        cfg = dict(base); cfg.update(extra)
        t0 = time.time(); x, it = pcg(A, b, lambda r: vcycle2(levels, 0, r, cfg), 1e-6); print(extra, "its", it, "%.1fs" % (time.time() - t0))

def vcycle3(levels, l, b, cfg):
    """cycle with a per-level visit multiplier cfg['visits'][l] (how often level l is visited per visit of level l-1)"""
    L = levels[l]; om = cfg["omega"]
    if l == len(levels) - 1:
        return smooth(L, np.zeros_like(b), b, "jacobi", cfg["coarse_sweeps"], omega=om)
    pre = cfg.get("pre_l", {}).get(l, cfg["pre"]); post = cfg.get("post_l", {}).get(l, cfg["post"])
    x = smooth(L, np.zeros_like(b), b, "jacobi", pre, omega=om)
    rc = L["P"].T @ (b - L["A"] @ x)
    ec = np.zeros(levels[l + 1]["A"].shape[0])
    for v in range(cfg["visits"].get(l + 1, 1) if l + 1 < len(levels) - 1 else 1):
        ec = ec + vcycle3(levels, l + 1, rc - levels[l + 1]["A"] @ ec, cfg)
    x = x + cfg["over"] * (L["P"] @ ec)
    return smooth(L, x, b, "jacobi", post, omega=om)

def sweep3(n, kind="dam"):
    t = build(n, kind); A, idx = assemble(t); A = A * 0.005
    gdt = -39.24 * 0.005
    vy = np.zeros(t.shape); up = np.roll(t, -1, axis=1)
    vy[(t != SOLID) & (up != SOLID)] = gdt
    b = -(vy - np.roll(vy, 1, axis=1))[t == WATER]
    b = b + np.random.default_rng(0).normal(0, 1.0, b.shape)
    levels = hierarchy(t, A, idx, min_cells=600)
    print(n, kind, "levels", [L["A"].shape[0] for L in levels])
    base = dict(pre=2, post=2, over=1.8, omega=0.8, coarse_sweeps=30)
    for extra in [dict(visits={1: 2, 2: 2}), dict(visits={2: 2}), dict(visits={2: 3}), dict(visits={2: 2, 3: 2}), dict(visits={2: 4}),
                  dict(visits={1: 2}), dict(visits={3: 2}), dict(visits={2: 2}, pre_l={1: 3}, post_l={1: 3}),
                  dict(visits={2: 2}, pre_l={2: 4, 3: 4}, post_l={2: 4, 3: 4}), dict(visits={}, pre_l={1: 4, 2: 8, 3: 8}, post_l={1: 4, 2: 8, 3: 8})]:
        cfg = dict(base); cfg.update(extra)
        t0 = time.time(); x, it = pcg(A, b, lambda r: vcycle3(levels, 0, r, cfg), 1e-6); print(extra, "its", it, "%.1fs" % (time.time() - t0))

def smooth_seq(L, x, b, omegas):
    A, D = L["A"], L["D"]
    for om in omegas:
        x = x + om * (b - A @ x) / D
    return x

def vcycle4(levels, l, b, cfg):
    L = levels[l]
    if l == len(levels) - 1:
        return smooth(L, np.zeros_like(b), b, "jacobi", cfg["coarse_sweeps"], omega=0.8)
    oms = cfg["omegas_l"].get(l, cfg["omegas"])
    x = smooth_seq(L, np.zeros_like(b), b, oms)
    rc = L["P"].T @ (b - L["A"] @ x)
    ec = np.zeros(levels[l + 1]["A"].shape[0])
    for v in range(cfg["visits"].get(l + 1, 1) if l + 1 < len(levels) - 1 else 1):
        ec = ec + vcycle4(levels, l + 1, rc - levels[l + 1]["A"] @ ec, cfg)
    x = x + cfg["over"] * (L["P"] @ ec)
    return smooth_seq(L, x, b, oms[::-1])

def sweep4(n, kind="dam"):
    t = build(n, kind); A, idx = assemble(t); A = A * 0.005
    gdt = -39.24 * 0.005
    vy = np.zeros(t.shape); up = np.roll(t, -1, axis=1)
    vy[(t != SOLID) & (up != SOLID)] = gdt
    b = -(vy - np.roll(vy, 1, axis=1))[t == WATER]
    b = b + np.random.default_rng(0).normal(0, 1.0, b.shape)
    levels = hierarchy(t, A, idx, min_cells=600)
    print(n, kind, "levels", [L["A"].shape[0] for L in levels])
    for oms, over, extra in [((0.8, 0.8), 1.8, {}), ((1.389, 0.5617), 1.8, {}), ((1.2, 0.6), 1.8, {}), ((1.0, 0.67), 1.8, {}), ((1.389, 0.5617), 2.0, {}),
                      ((1.1, 0.62), 1.8, {}), ((0.9, 0.9), 1.8, {}), ((1.0, 1.0), 1.8, {}), ((1.5, 0.75, 0.5), 1.8, {}), ((0.8, 0.8, 0.8), 1.8, {}),
                      ((0.8, 0.8), 1.8, {1: (0.8, 0.8, 0.8), 2: (0.8,) * 4, 3: (0.8,) * 4}),
                      ((1.389, 0.5617), 1.8, {1: (1.5, 0.75, 0.5), 2: (1.5, 0.75, 0.5), 3: (1.5, 0.75, 0.5)})]:
        cfg = dict(omegas=oms, over=over, coarse_sweeps=30, visits={2: 2}, omegas_l=extra)
        t0 = time.time(); x, it = pcg(A, b, lambda r: vcycle4(levels, 0, r, cfg), 1e-6); print(oms, over, extra, "its", it, "%.1fs" % (time.time() - t0))
