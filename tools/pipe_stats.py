"""Measurement helper: per-role cycle counters of the assembly pipeline on the config-2 mesh."""
import sys
import numpy as np
sys.path.insert(0, '.')
import torch
from fempy_b200 import AssemblyPlan, Constitutive, Elements, meshes, _lib
from fempy_b200._lib import iptr, check
pn = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
coords, conn = meshes.quad4_grid(n, n)
N = coords.shape[0]; E = n * n
plan = AssemblyPlan(N); plan.setLocalityHint(coords); team = int(sys.argv[4]) if len(sys.argv) > 4 else 2
plan.setOption("assembly_path", 1)
plan.setOption("team_warps", team)
if pn: plan.setOption("patch_nodes", pn)
dbg = int(sys.argv[3]) if len(sys.argv) > 3 else 0
Nn, dN, w = Elements.QuadElement2D(1).tables(); plan.addBlock(Nn, dN, w, conn["quad"]); plan.finalize()
plan.setOption("debug_flags", dbg)
left = meshes.edge_nodes(coords, x=0.0); bc = np.concatenate((2 * left, 2 * left + 1)); plan.setBCs(bc, np.zeros(bc.size))
dev = torch.device("cuda")
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
dc, ds, dt, dl = d(coords), d(1e-3 * np.random.default_rng(0).random((N, 2))), d(np.ones(E)), d(np.zeros(2 * N))
dK = torch.empty(plan.nnz, dtype=torch.float64, device=dev); dR = torch.empty(2 * N, dtype=torch.float64, device=dev)
mat = Constitutive.IsoPlaneStress(70e9, 0.3, 2700.0, 1.0).material()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
lib = _lib.load()
for _ in range(3):
    plan.assembleDev(dc, ds, dt, mat, True, dl, dK, dR)
torch.cuda.synchronize()
check(lib.femb_plan_pipe_stats(plan.handle, 1, None, 0))
flush.zero_(); torch.cuda.synchronize()
plan.assembleDev(dc, ds, dt, mat, True, dl, dK, dR); plan.sync(); torch.cuda.synchronize()
out = np.zeros(16 * 256, dtype=np.int64)
check(lib.femb_plan_pipe_stats(plan.handle, 1, iptr(out), out.size))
s = out.reshape(256, 16)[:148]
names = ["T0 total", "T0 patches", "T0 wait+prefetch", "T0 zero+compute", "T0 scatter", "T0 writeout", "T1 total", "T1 patches", "T1 wait+prefetch", "T1 zero+compute", "T1 scatter", "T1 writeout"]
print("debug", dbg, "team", team, "pn", plan.info("patch_nodes"), "patches", plan.info("num_patches"), "elems", plan.info("patch_elems_total"), "maxE", plan.info("patch_elems_max"))
for i, nm in enumerate(names):
    print(f"{nm:16s} mean {s[:, i].mean():10.0f}  min {s[:, i].min():8d}  max {s[:, i].max():8d}")
