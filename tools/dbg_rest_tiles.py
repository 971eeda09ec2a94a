"""Which tiles are neither brick nor edge lattice tiles?  python tools/dbg_rest_tiles.py [n]"""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from tools.quick_bench import cube
from tianxin_b200 import capi
dev = torch.device('cuda:0'); n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lids, xyz = cube(n, dev)
h = capi.Handle(scatter_mode=1); h.block_add(lids, node_coords=xyz, n_rows=(n + 1) ** 3); h.graph_build(); h.terms_set(capi.poisson_terms()); h.setup()
i = h.info(); print("tiles", i.n_tiles, "brick", i.n_brick_tiles, "edge", i.n_edge_tiles, "te_max", i.tile_cells_max)
s = n + 1
for t in range(i.n_brick_tiles + i.n_edge_tiles, i.n_tiles):
    rows, cells, adjl = h.tile_get(t)
    rows = rows[rows >= 0]
    ijk = np.stack([rows % s, (rows // s) % s, rows // (s * s)], 1)
    print("tile", t, "nrows", len(rows), "ncells", len(cells), "rows box", ijk.min(0), ijk.max(0))
