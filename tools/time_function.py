"""Time the element-function kernel (von Mises / mass recovery) with CUDA events, L2 flushed between launches.
usage: python tools/time_function.py [quad|quad9|triangle] [nx]"""
import sys
import numpy as np
sys.path.insert(0, '.')
import torch
from fempy_b200 import AssemblyPlan, Constitutive, Elements, meshes
elType = sys.argv[1] if len(sys.argv) > 1 else "quad"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
coords, conn = meshes.GENERATORS[elType](n, n)
N = coords.shape[0]; E = conn[elType].shape[0]
el = {"quad": lambda: Elements.QuadElement2D(1), "quad9": lambda: Elements.QuadElement2D(2), "triangle": lambda: Elements.TriElement2D(1)}[elType]()
plan = AssemblyPlan(N); plan.addBlock(*el.tables(), conn[elType]); plan.finalize()
dev = torch.device("cuda")
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
dc, ds = d(coords), d(1e-3 * np.random.default_rng(0).random((N, 2)))
out = torch.empty(E, dtype=torch.float64, device=dev)
mat = Constitutive.IsoPlaneStress(70e9, 0.3, 2700.0, 1.0).material()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.Stream(); plan.setStream(st.cuda_stream)
nn = conn[elType].shape[1]
algBytes = 4 * nn * E + 32 * N + 8 * E
for name, func, red in (("von Mises, mean", 1, 2), ("mass, integrate", 0, 5)):
    ts = []
    for it in range(13):
        flush.zero_(); flush[: 256 << 20].sum(); torch.cuda.synchronize()
        with torch.cuda.stream(st):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(200000)  # keep the GPU busy while the host enqueues: the interval holds no launch latency
            e0.record(st); plan.functionDev(dc, ds, mat, func, red, out); e1.record(st)
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts[3:]))
    print(f"{elType} nx={n} E={E} {name}: {1e3 * t:.1f} us, {E / t / 1e6:.2f} G elements/s, algorithmic {algBytes / 1e6:.1f} MB -> {algBytes / t / 1e6:.0f} GB/s ({100 * algBytes / t / 1e6 / 6506.6:.1f}% of 6506.6)")
