import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from tools.quick_bench import cube
from tianxin_b200 import capi
dev=torch.device('cuda:0'); n=32
lids,xyz=cube(n,dev)
h=capi.Handle(scatter_mode=1); h.block_add(lids,node_coords=xyz,n_rows=(n+1)**3); h.graph_build(); h.terms_set(capi.poisson_terms()); h.setup()
i=h.info(); print("tiles",i.n_tiles,"te_max",i.tile_cells_max)
for t in (0, 21, 40):
    rows,cells,adjl=h.tile_get(t)
    s=n+1
    ijk=np.stack([rows%s,(rows//s)%s,rows//(s*s)],1)
    print("tile",t,"ncells",len(cells),"rows box",ijk.min(0),ijk.max(0))
    print(" first 20 rows ijk", ijk[:20].tolist())
    for a in range(8):
        print("  a",a, adjl[:32,a].tolist())
