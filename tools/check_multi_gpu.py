"""Launched with torchrun on N GPUs: the stacked rank row blocks of the element-partitioned assembly must equal
the single-GPU assembly (pattern bit for bit, values to 1e-12).  Prints MULTI_GPU_PARITY_OK on rank 0."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from fempy_b200 import AssemblyPlan, Constitutive, Elements, meshes
from fempy_b200.dist import DistributedAssembly, GpuLocalAssembler, Partition

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for mode, elType, n in (("halo", "quad", 96), ("halo", "triangle", 70), ("exchange", "quad", 96), ("exchange", "triangle", 70)):
    coords, conn = meshes.GENERATORS[elType](n, n + 13)
    N = coords.shape[0]
    numEl = conn[elType].shape[0]
    rng = np.random.default_rng(3)
    states = 1e-3 * rng.random((N, 2))
    thick = rng.random(numEl) * 1e-2 + 1e-3
    loads = rng.random(2 * N)
    left = meshes.edge_nodes(coords, x=0.0)
    bcDOF = np.concatenate((2 * left, 2 * left + 1, [2 * (N - 1)]))
    bcVal = rng.random(bcDOF.size) * 1e-4
    el = Elements.QuadElement2D(1) if elType == "quad" else Elements.TriElement2D(1)
    mat = Constitutive.IsoPlaneStress(70e9, 0.3, 2700.0, 1.0).material()
    part = Partition(N, conn, rank, world, mode=mode)
    asm = GpuLocalAssembler(part, {elType: el.tables()}, coords, mat, local)
    da = DistributedAssembly(part, asm, dev)
    da.set_bcs(bcDOF, bcVal)
    loadsLocal = np.zeros(2 * part.numLocalNodes)
    loadsLocal[part.owned_dof_slice()] = loads[2 * part.n0 : 2 * part.n1]
    Kb, Rb = da.assemble(part.local_thickness(thick, {elType: 0}), states[part.localNodes], loadsLocal, True)
    gathered = [None] * world
    dist.all_gather_object(gathered, (da.blockIndptr, da.blockIndices, Kb.cpu().numpy(), Rb.cpu().numpy()))
    if rank == 0:
        plan = AssemblyPlan(N, device=local)
        plan.addBlock(*el.tables(), conn[elType])
        plan.finalize()
        Kref, Rref = plan.assemble(coords, states, thick, mat, bcDOF, bcVal, True, loads)
        ip, ix = plan.pattern()
        indptr = np.concatenate([[0]] + [g[0][1:] + sum(len(h[1]) for h in gathered[:i]) for i, g in enumerate(gathered)])
        assert np.array_equal(indptr, ip) and np.array_equal(np.concatenate([g[1] for g in gathered]), ix), "pattern differs"
        K = np.concatenate([g[2] for g in gathered])
        R = np.concatenate([g[3] for g in gathered])
        ek, er = np.abs(K - Kref).max() / np.abs(Kref).max(), np.abs(R - Rref).max() / np.abs(Rref).max()
        assert ek < 1e-12 and er < 1e-12, (ek, er)
        print(f"{elType}: {world} ranks, K err {ek:.2e}, R err {er:.2e}")
if rank == 0:
    print("MULTI_GPU_PARITY_OK")
dist.destroy_process_group()
