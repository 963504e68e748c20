import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
from scipy.sparse import csr_array
from scipy.sparse.linalg import spsolve
from test_gpu_parity import _path_case, _raw_plan
from util import make_cm
# why tests/test_gpu_parity.py::test_pcg_matches_direct_solve compares with the reduced system: SuperLU on the row-only-BC matrix
coords, conn, states, thick, loads, bcDOF, bcVal = _path_case("quad", 40)
thick=thick*100
plan=_raw_plan("quad", coords, conn)
K,R=plan.assemble(coords, states, thick, make_cm(0).material(), bcDOF, bcVal, True, 1e6*loads)
indptr,indices=plan.pattern()
A=csr_array((K,indices,indptr),shape=(R.size,R.size))
want=spsolve(A.tocsc(),-R)
x,it,res=plan.solvePCG(K,-R,relTol=1e-12)
u=np.unique(bcDOF); cnt=np.bincount(bcDOF,minlength=R.size)
print("iters",it,res)
print("bc x   ", x[u][:6]); print("bc want", want[u][:6]); print("-R/cnt ", (-R/np.maximum(cnt,1))[u][:6])
print("row0 of A", A[[0]].toarray()[0][:6], "nnz in row", A[[0]].nnz, "sum abs offdiag", abs(A[[0]].toarray()[0]).sum()-abs(A[0,0]))
d=np.abs(x-want); print("max diff at", d.argmax(), d.max(), "is bc:", d.argmax() in set(u.tolist()))
free=np.setdiff1d(np.arange(R.size),u); print("free diff", d[free].max()/abs(want).max(), "bc diff", d[u].max()/abs(want).max())
print("res want", abs(A@want+R).max(), "res x", abs(A@x+R).max(), abs((A@x+R)[u]).max())
