import sys; sys.path.insert(0,".")
import numpy as np, torch
from tools.quick_bench import cube
from tianxin_b200 import capi
dev=torch.device("cuda:0")
n=int(sys.argv[1]); pert=float(sys.argv[2])
lids,xyz=cube(n,dev,pert)
nr=(n+1)**3
res={}
for mode in (1,3):
    h=capi.Handle(scatter_mode=mode); h.block_add(lids,node_coords=xyz,n_rows=nr); nnz=h.graph_build(); h.terms_set(capi.poisson_terms()); h.setup()
    x=torch.sin(0.37*torch.arange(nr,device=dev,dtype=torch.float64))
    f=torch.empty(nr,device=dev,dtype=torch.float64); A=torch.full((nnz,),float("nan"),device=dev,dtype=torch.float64)
    for rep in range(2):
        h.evaluate(capi.JACOBIAN,x,f,A,flags=2); h.sync()
    rp=np.empty(nr+1,np.int64); ci=np.empty(nnz,np.int32); h.graph_get(rp,ci)
    res[mode]=(f.clone(),A.clone())
    print("mode",mode,"nan",int(torch.isnan(A).sum()),"asum",float(A.abs().sum()))
    h.close()
d=(res[1][1]-res[3][1]).abs()
bad=(d>1e-9)|torch.isnan(d)
print("mismatch entries",int(bad.sum()))
if bad.any():
    idx=bad.nonzero().flatten().cpu().numpy()
    rows=np.searchsorted(rp,idx,side="right")-1
    ur=np.unique(rows)
    print("rows affected",len(ur),ur[:40])
    s=n+1
    print("ijk",[(int(r%s),int((r//s)%s),int(r//(s*s))) for r in ur[:20]])
    print("idx", idx[:30])
    print("vals tile", res[1][1][idx[:10]].cpu().numpy(), "ref", res[3][1][idx[:10]].cpu().numpy())
