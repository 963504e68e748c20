import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'./tests')
from fempy_b200 import AssemblyPlan, meshes
from util import ELEMENT_FOR, make_cm, oracle_blocks, relerr, MAT
from oracle import fem_oracle as fo
for elType,n in (("quad",6),("quad",40),("triangle",33),("quad9",12),("quad",200)):
    coords, conn = meshes.GENERATORS[elType](n, n)
    N=coords.shape[0]; E=conn[elType].shape[0]
    rng=np.random.default_rng(0); states=1e-3*rng.random((N,2)); thick=rng.random(E)+0.5; loads=rng.random(2*N)
    left=meshes.edge_nodes(coords,x=0.0); bc=np.concatenate((2*left,2*left+1)); bv=rng.random(bc.size)*1e-4
    plan=AssemblyPlan(N); plan.setLocalityHint(coords); 
    Nn,dN,w=ELEMENT_FOR[elType]().tables(); plan.addBlock(Nn,dN,w,conn[elType]); plan.finalize()
    print(elType,n,'path',plan.info('assembly_path'),'rows',plan.info('rows_path'),'maxdeg',plan.info('max_degree'),'patch_path',plan.info("patch_path"),'patches',plan.info("num_patches"),'maxE',plan.info("patch_elems_max"),'blob',plan.info("patch_blob_bytes"),'PN',plan.info("patch_nodes"),'colours',plan.info("patch_colours_max"), flush=True)
    mat=make_cm(0).material()
    K,R=plan.assemble(coords,states,thick,mat,bc,bv,True,loads)
    print(' launches',plan.lastLaunches, flush=True)
    bl=oracle_blocks(conn)
    Ko=fo.assemble_matrix(bl,coords,states,thick,0,MAT["E"],MAT["nu"],bc,dropZeros=False); Ko.sort_indices()
    Ro=fo.assemble_residual(bl,coords,states,thick,0,MAT["E"],MAT["nu"],loads,bc,bv)
    print(' K',relerr(K,Ko.data),'R',relerr(R,Ro), flush=True)
