"""One setup of the 256^3 workload, then the fill under several run-time options (CUDA events on the handle's stream).
  python tools/sweep_fill.py [--n 256] [--iters 20]"""
import argparse, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tianxin_b200 import capi, host
from tianxin_b200.assembly_engine import build_poisson_problem


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--perturb", type=float, default=0.0)
    ap.add_argument("--only", default="", help="comma-separated combo names")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    t0 = time.time()
    prob = build_poisson_problem(a.n, stream=stream.cuda_stream, perturb=a.perturb)
    h = prob.handle
    i = h.info()
    print(f"setup {time.time() - t0:.1f}s (txasm_setup {i.setup_ms:.0f} ms) tiles={i.n_tiles} uniform={i.n_uniform_tiles} brick={i.n_brick_tiles} edge={i.n_edge_tiles}", flush=True)
    x = torch.from_numpy(host.state_by_gid(prob.dof.getOwnedAndGhostedIndices())).to(dev)
    f = torch.empty(prob.n_local, dtype=torch.float64, device=dev)
    A = torch.empty(prob.nnz, dtype=torch.float64, device=dev)

    def timed(flags):
        for _ in range(3):
            h.evaluate(capi.JACOBIAN, x, f, A, flags=flags)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(a.iters):
                h.evaluate(capi.JACOBIAN, x, f, A, flags=flags)
            e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters

    combos = [
        ("default", {}),
        ("no concurrent", {"concurrent_fill": 0}),
        ("no edge kernel", {"edge_kernel": 0}),
        ("brick 3 + rest 2", {"brick_ctas_per_sm": 3, "rest_ctas_per_sm": 2}),
        ("brick 3 + rest 1", {"brick_ctas_per_sm": 3, "rest_ctas_per_sm": 1}),
        ("brick 2 + rest 2", {"brick_ctas_per_sm": 2, "rest_ctas_per_sm": 2}),
        ("brick 3 CTA/SM", {"brick_ctas_per_sm": 3}),
        ("brick 3 CTA/SM no concurrent", {"brick_ctas_per_sm": 3, "concurrent_fill": 0}),
        ("no fuse dirichlet", {"fuse_dirichlet": 0}),
        ("brick 2 CTA/SM", {"brick_ctas_per_sm": 2}),
        ("brick 3 CTA/SM", {"brick_ctas_per_sm": 3}),
        ("brick 4 CTA/SM", {"brick_ctas_per_sm": 4}),
        ("brick 3 CTA/SM no concurrent", {"brick_ctas_per_sm": 3, "concurrent_fill": 0}),
        ("brick 2 CTA/SM + rest 1 CTA/SM concurrent", {"brick_ctas_per_sm": 2, "rest_ctas_per_sm": 1}),
        ("brick 3 CTA/SM + rest 1 CTA/SM concurrent", {"brick_ctas_per_sm": 3, "rest_ctas_per_sm": 1}),
        ("uniform kernel (no brick)", {"brick_kernel": 0}),
        ("uniform kernel (no brick) no concurrent", {"brick_kernel": 0, "concurrent_fill": 0}),
        ("rowtile only", {"uniform_kernel": 0}),
    ]
    defaults = {"concurrent_fill": 1, "fuse_dirichlet": 1, "brick_ctas_per_sm": 0, "brick_kernel": 1, "uniform_kernel": 1, "rest_ctas_per_sm": 0, "edge_kernel": 1}
    only = [o for o in a.only.split(",") if o]
    for name, opts in combos:
        if only and name not in only:
            continue
        for k, v in defaults.items():
            h.option_set(k, opts.get(k, v))
        vol, full = timed(2), timed(15)
        if os.environ.get("TXASM_TIMELINE") == "1":
            h.evaluate(capi.JACOBIAN, x, f, A, flags=2); h.sync()
            tl = h.debug_timeline()
            print("   timeline (us): brick %.0f..%.0f  edge %.0f..%.0f  rowtile %.0f..%.0f" % tuple(tl),
                  {k: h.option_get(k) for k in ("concurrent_fill", "edge_kernel", "brick_ctas_per_sm", "rest_ctas_per_sm")}, flush=True)
        print(f"{name:45s} volume {vol:.3f} ms  evaluate(All) {full:.3f} ms  -> {prob.n_cells / full / 1e3:.0f} Melem/s, "
              f"{288 * prob.n_cells / vol / 1e6 / 6468.6:.3f} of HBM (volume)", flush=True)
    h.close()


if __name__ == "__main__":
    main()
