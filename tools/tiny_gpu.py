import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from fempy_b200 import AssemblyPlan, meshes
from util import ELEMENT_FOR, make_cm
coords, conn = meshes.quad4_grid(6, 6)
N=coords.shape[0]; E=36
plan=AssemblyPlan(N); plan.setLocalityHint(coords)
Nn,dN,w=ELEMENT_FOR["quad"]().tables(); plan.addBlock(Nn,dN,w,conn["quad"]); plan.finalize()
plan.setOption("debug_flags", int(sys.argv[1]) if len(sys.argv)>1 else 0)
K,R=plan.assemble(coords,np.zeros((N,2)),np.ones(E),make_cm(0).material(),applyBCs=False)
print(K[:4], R[:4])
