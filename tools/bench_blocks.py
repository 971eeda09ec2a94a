"""Throughput of the general element blocks (BASELINE.json configs 3-5) on one B200: one Jacobian-type volume fill per
step, CUDA events on the handle's stream.  Algorithmic bytes per element from SURVEY.md section 8d.
  python tools/bench_blocks.py [--n-elastic 128] [--n-hcurl 128] [--n-mixed 48] [--steps 10]"""
import argparse, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tianxin_b200 import capi, blocks


def _time(h, n_rows, nnz, steps, need=(True, False, False), **kw):
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1)
    vec = [torch.randn(n_rows, dtype=torch.float64, generator=g).to(dev) if nd else None for nd in need]
    f = torch.empty(n_rows, dtype=torch.float64, device=dev); A = torch.empty(nnz, dtype=torch.float64, device=dev)
    run = lambda: h.evaluate(capi.JACOBIAN, vec[0], f, A, xdot=vec[1], xdotdot=vec[2], flags=capi.FLAG_VOLUMETRIC_FILL, **kw)
    for _ in range(3):
        run()
    h.sync(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, float(A.abs().sum()), float(f.abs().sum())


def run_blocks(n_elastic=128, n_hcurl=128, n_mixed=48, steps=10, peak_gbs=6468.6):
    dev = torch.device("cuda:0")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = {}
    # ---- config 4: transient linear elastodynamics, 3 DOF per node, Q1 hexahedra
    n = n_elastic
    ijk, dims = blocks.hex_cells(n)
    lids = blocks.three_field_lids(blocks.q1_node_lids(ijk, dims)); n_rows = 3 * (n + 1) ** 3
    h = capi.Handle()
    t0 = time.time()
    b = h.gblock_add(capi.TOPO_HEX8, capi.BASIS_HGRAD_C1, 2, T(blocks.hex_vertex_coords(ijk, dims)), T(lids), n_rows, n_fields=3)
    h.gblock_terms_set(b, capi.OP_ELASTICITY, [1.3, 0.7, 2.5, 0.0])
    nnz = h.graph_build(); h.setup()
    ts = time.time() - t0
    ms, cs, _ = _time(h, n_rows, nnz, steps, need=(True, False, True), beta=1.0, gamma=4.0)
    ne = len(ijk)
    out["elastodynamics_q1hex"] = {"workload": f"elastodynamics_q1hex_{n}^3 (3 DOF/node, rho M d2u/dt2 + K u; J = gamma rho M + beta K)", "elements": ne,
                                   "rows": n_rows, "nnz": nnz, "ms_per_step": ms, "Melem_per_s": ne / ms / 1e3, "bytes_per_element": 2160,
                                   "hbm_frac": 2160 * ne / (ms * 1e-3) / 1e9 / peak_gbs, "setup_s": round(ts, 2), "parity": "unpinned (SURVEY appendix B); tests/test_blocks_gpu.py vs oracle",
                                   "kernel": "k_gblock<ELASTICITY,8,8,3> (thread per element DOF row, planned atomic sumInto)"}
    h.close(); del h
    torch.cuda.empty_cache()
    # ---- config 5: HCURL edge elements, curl-curl + mass
    n = n_hcurl
    ijk, dims = blocks.hex_cells(n)
    el, sg, n_rows = blocks.hcurl_hex_lids(ijk, dims)
    h = capi.Handle()
    t0 = time.time()
    b = h.gblock_add(capi.TOPO_HEX8, capi.BASIS_HCURL_I1, 2, T(blocks.hex_vertex_coords(ijk, dims)), T(el), n_rows, orientation_signs=T(sg))
    h.gblock_terms_set(b, capi.OP_CURLCURL, [1.0, 1.0, 0.0])
    nnz = h.graph_build(); h.setup()
    ts = time.time() - t0
    ms, cs, _ = _time(h, n_rows, nnz, steps)
    ne = len(ijk)
    out["maxwell_hcurl_hex"] = {"workload": f"hcurl_i1_hex_{n}^3 curl-curl + mass with edge orientations", "elements": ne, "rows": n_rows, "nnz": nnz,
                                "ms_per_step": ms, "Melem_per_s": ne / ms / 1e3, "bytes_per_element": 912,
                                "hbm_frac": 912 * ne / (ms * 1e-3) / 1e9 / peak_gbs, "setup_s": round(ts, 2),
                                "parity": "unpinned; tests/test_blocks_gpu.py vs oracle", "kernel": "k_gblock<CURLCURL,12,8,1>"}
    h.close(); del h
    torch.cuda.empty_cache()
    # ---- config 3: mixed mesh, Q2 hexahedra (x < 1/2) + P2 tetrahedra (x > 1/2), conforming across the interface
    n = n_mixed
    ijk_h, dims = blocks.hex_cells(n, x_range=(0, n // 2))
    ijk_t, _ = blocks.hex_cells(n, x_range=(n // 2, n))
    tv, txyz = blocks.cube_tets(ijk_t, dims)
    # one LID space: Q2 nodes live on the (2n+1)^3 lattice, P2 tet nodes on the (8n+1)^3 one (vertices at quarter points doubled)
    (lh, lt), n_rows = blocks.compress(4 * (2 * ijk_h[:, None, :] + 1 + blocks.HEX27[None, :, :]), blocks.p2_tet_positions(tv), dims=dims)
    res = {}
    for dmma in (1, 0):
        h = capi.Handle()
        t0 = time.time()
        b0 = h.gblock_add(capi.TOPO_HEX27, capi.BASIS_HGRAD_C2, 4, T(blocks.hex_vertex_coords(ijk_h, dims)), T(lh), n_rows)
        b1 = h.gblock_add(capi.TOPO_TET10, capi.BASIS_HGRAD_C2, 2, T(txyz), T(lt), n_rows)
        h.gblock_terms_set(b0, capi.OP_DIFFUSION, [1.0]); h.gblock_terms_set(b1, capi.OP_DIFFUSION, [1.0])
        nnz = h.graph_build(); h.setup()
        h.option_set("dmma", dmma)
        ts = time.time() - t0
        ms, cs, fs = _time(h, n_rows, nnz, steps)
        res[dmma] = (ms, cs, fs, nnz, ts)
        h.close(); del h
        torch.cuda.empty_cache()
    ne_h, ne_t = len(ijk_h), len(tv)
    out["mixed_q2hex_p2tet"] = {"workload": f"scalar diffusion, Q2 hexahedra ({n // 2}x{n}x{n}) + P2 tetrahedra (12 per hexahedron of the other half), conforming",
                                "hex_elements": ne_h, "tet_elements": ne_t, "rows": n_rows, "nnz": res[1][3],
                                "ms_per_step_dmma": res[1][0], "ms_per_step_dfma": res[0][0], "Melem_per_s": (ne_h + ne_t) / res[1][0] / 1e3,
                                "bytes_per_q2_hex": 4360, "hbm_frac_hex_bytes_only": 4360 * ne_h / (res[1][0] * 1e-3) / 1e9 / peak_gbs,
                                "checksum_rel_diff_dmma_vs_dfma": abs(res[1][1] - res[0][1]) / abs(res[0][1]), "setup_s": round(res[1][4], 2),
                                "parity": "unpinned (the reference throws on mixed topologies); tests/test_blocks_gpu.py vs oracle",
                                "kernel": "k_gblock_q2_dmma (mma.sync m8n8k4 f64) + k_gblock<DIFFUSION,10,4,1>"}
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-elastic", type=int, default=128)
    ap.add_argument("--n-hcurl", type=int, default=128)
    ap.add_argument("--n-mixed", type=int, default=48)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    print(json.dumps(run_blocks(a.n_elastic, a.n_hcurl, a.n_mixed, a.steps)), flush=True)
