"""Scratch timing of the fill kernels on a structured cube (lexicographic node ids, not the
reference numbering) -- development aid, not the contract bench (see bench.py)."""
import sys, time, json
import torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from tianxin_b200 import capi


def cube(n, dev, perturb=0.0):
    i = torch.arange(n, device=dev)
    ez, ey, ex = torch.meshgrid(i, i, i, indexing="ij")
    n0 = (ex + ey * (n + 1) + ez * (n + 1) ** 2).reshape(-1)
    s = n + 1
    lids = torch.stack([n0, n0 + 1, n0 + 1 + s, n0 + s, n0 + s * s, n0 + 1 + s * s, n0 + 1 + s + s * s, n0 + s + s * s], 1).to(torch.int32).contiguous()
    j = torch.arange(s, device=dev, dtype=torch.float64) / n
    zz, yy, xx = torch.meshgrid(j, j, j, indexing="ij")
    xyz = torch.stack([xx, yy, zz], -1).reshape(-1, 3).contiguous()
    if perturb:
        g = torch.Generator(device=dev); g.manual_seed(1)
        inner = ((xyz > 1e-9) & (xyz < 1 - 1e-9)).all(1, keepdim=True)
        xyz = xyz + inner * perturb / n * (torch.rand(xyz.shape, generator=g, device=dev, dtype=torch.float64) - 0.5)
    return lids, xyz


def run(n, mode, perturb=0.0, reps=5):
    dev = torch.device("cuda:0")
    lids, xyz = cube(n, dev, perturb)
    nrows = (n + 1) ** 3
    h = capi.Handle(scatter_mode=mode)
    h.block_add(lids, node_coords=xyz, n_rows=nrows)
    t0 = time.time(); nnz = h.graph_build(); torch.cuda.synchronize(); tg = time.time() - t0
    h.terms_set(capi.poisson_terms())
    t0 = time.time(); h.setup(); torch.cuda.synchronize(); ts = time.time() - t0
    h.option_set("stage_timers", 1)
    x = torch.sin(0.37 * torch.arange(nrows, device=dev, dtype=torch.float64))
    f = torch.empty(nrows, device=dev, dtype=torch.float64)
    A = torch.empty(nnz, device=dev, dtype=torch.float64)
    ms = []
    for r in range(reps):
        h.evaluate(capi.JACOBIAN, x, f, A, flags=capi.FLAG_VOLUMETRIC_FILL)
        ms.append(h.last_fill_ms())   # (needs option stage_timers = 1, set below)
    A.fill_(float("nan")); f.fill_(float("nan"))
    h.evaluate(capi.JACOBIAN, x, f, A, flags=capi.FLAG_VOLUMETRIC_FILL); h.sync()
    nan_A, nan_f = int(torch.isnan(A).sum()), int(torch.isnan(f).sum())
    vol = h.timers().evaluate_volume
    info = h.info()
    ne = n ** 3
    best = min(ms[1:]) if len(ms) > 1 else ms[0]
    out = dict(n=n, mode=info.scatter_mode, perturb=perturb, nnz=nnz, graph_s=round(tg, 3), setup_s=round(ts, 3),
               fill_ms=[round(m, 3) for m in ms], volume_ms=round(vol, 3), melem_s=round(ne / best / 1e3, 1),
               gbs=round(288 * ne / best / 1e6, 1), frac_hbm=round(288 * ne / best / 1e6 / 6468.6, 4),
               affine=info.n_affine_cells, te_max=info.tile_cells_max, smem=info.smem_bytes, ctas=info.ctas_per_sm, nan_A=nan_A, nan_f=nan_f, fsum=float(f.sum()), asum=float(A.abs().sum()))
    print(json.dumps(out), flush=True)
    h.close()


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("sizes", nargs="*", type=int, default=[128])
    ap.add_argument("--modes", default="2,3,1")
    ap.add_argument("--perturb", default="0.0,0.2")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    for n in a.sizes:
        for mode in [int(m) for m in a.modes.split(",")]:
            for p in [float(v) for v in a.perturb.split(",")]:
                try:
                    run(n, mode, p, a.reps)
                except Exception as e:
                    print("FAIL", n, mode, p, e, flush=True)
