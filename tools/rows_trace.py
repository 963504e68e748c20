"""Measurement helper: warp lifetimes / start times of the row-owner kernel.  Needs the trace build:
tools/knock_variants.sh trace && cp tools/micro/lib_trace.so fempy_b200/csrc/libfemb200.so (on the GPU box), then
python tools/rows_trace.py [nx]"""
import sys
import numpy as np
sys.path.insert(0, '.')
import torch
from fempy_b200 import AssemblyPlan, Constitutive, Elements, meshes, _lib
from fempy_b200._lib import iptr, check
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
coords, conn = meshes.quad4_grid(n, n)
N = coords.shape[0]; E = n * n
plan = AssemblyPlan(N)
Nn, dN, w = Elements.QuadElement2D(1).tables(); plan.addBlock(Nn, dN, w, conn["quad"]); plan.finalize()
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
check(lib.femb_plan_row_stats(plan.handle, 1, None, 0))
flush.zero_(); flush.sum(); torch.cuda.synchronize()
plan.assembleDev(dc, ds, dt, mat, True, dl, dK, dR); plan.sync(); torch.cuda.synchronize()
out = np.zeros(16 * 256, dtype=np.int64)
check(lib.femb_plan_row_stats(plan.handle, 1, iptr(out), out.size))
warps = out[1]
print(f"nx {n}: warps {warps}, lifetime cycles mean {out[0] / max(warps, 1):.0f} min {out[5]} max {out[2]}, first start -> last end {1e-3 * (out[4] - out[3]):.1f} us")
print("starts per us:", out[16:80][: int(1e-3 * (out[4] - out[3])) + 2].tolist())
print("lifetimes per 2000 cycles:", out[96:160][: out[2] // 2000 + 1].tolist())
