#!/usr/bin/env python
"""Benchmark of the FEMpy assembly hot path on B200 (metric of BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the CPU port of the reference's algorithm

A step is ONE fused assembly of the CSR stiffness values + residual vector (BCs and loads applied) of the
workload mesh.  At N=1 the workload is BASELINE config 2: structured 4-node quad mesh, 512 x 512 elements
(526 338 DOF), isotropic plane stress.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAT = dict(E=70e9, nu=0.3, rho=2700.0, t=1.0)
WORKLOADS = {  # name -> (element type, nx = ny, description)
    "c2": ("quad", 512, "config2: structured Q4 512x512, 526338 DOF, iso plane stress"),
    "c3": ("quad9", 500, "config3: structured Q9 500x500, 2004002 DOF, iso plane stress"),
    "c4": ("triangle", 1580, "config4: structured T3 2x1580x1580, 4999122 DOF, iso plane stress"),
    "c5": ("quad", 5000, "config5: structured Q4 5000x5000, 50020002 DOF, iso plane stress"),
}
METRIC = "elements/s assembled (CSR K + residual)"


def make_workload(name, nx=None, ny=None):
    from fempy_b200 import meshes

    elType, n, desc = WORKLOADS[name]
    n = nx or n
    coords, conn = meshes.GENERATORS[elType](n, ny or n)
    if nx or ny:
        desc += f" [mesh overridden: {n} x {ny or n} cells]"
    N = coords.shape[0]
    numEl = conn[elType].shape[0]
    rng = np.random.default_rng(0)
    states = 1e-3 * rng.random((N, 2))
    thick = np.full(numEl, MAT["t"])
    left, right = meshes.edge_nodes(coords, x=0.0), meshes.edge_nodes(coords, x=2.0)
    bcDOF = np.stack((2 * left, 2 * left + 1), axis=1).reshape(-1).astype(np.int64)
    bcVal = np.zeros(bcDOF.size)
    loads = np.zeros(2 * N)
    loads[2 * right] += 1e4
    return dict(elType=elType, n=n, desc=desc, coords=coords, conn=conn, states=states, thick=thick, bcDOF=bcDOF, bcVal=bcVal,
                loads=loads, numNodes=N, numElems=numEl)


def algorithmic_bytes(wl, nnz):
    """Compulsory HBM traffic of one K+R assembly: int32 connectivity, coords, states,
    thickness read once; K values and R written once."""
    nn = wl["conn"][wl["elType"]].shape[1]
    return 4 * nn * wl["numElems"] + 16 * wl["numNodes"] * 2 + 8 * wl["numElems"] + 8 * nnz + 16 * wl["numNodes"]


class ClockSampler:
    """SM clock and throttle reasons of one GPU through NVML, sampled DURING the timed region: `sample()` is called from
    the timing loop every few steps, right after the L2 flush has been queued and before the step's start event -- the GPU
    is under load, but no NVML query lands inside a timed interval (a sampling thread did: at 2 GPUs it inflated single
    steps of one rank by milliseconds)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index, every=8):
        self.index, self.samples, self.reasons, self.maxMhz, self.every, self.calls = index, [], set(), None, every, 0
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.maxMhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def start(self):
        self.sample(force=True)

    def sample(self, force=False):
        self.calls += 1
        if self.nv is None or not (force or self.calls % self.every == 0):
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def stop(self):
        self.sample(force=True)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.maxMhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class CpuPort:
    """The CPU port (oracle/fem_oracle.py: the reference's stages, FEMpy/Problem.py:565-661) on all host cores:
    one worker process per core (oracle/cpu_port.py), each assembling K (COO -> CSR) + R with BCs and loads of its
    own strip of `rows` element rows of the workload mesh."""

    def __init__(self, wl, procs=None):
        from oracle import cpu_port

        self.cp, self.wl = cpu_port, wl
        self.procs = min(procs or host_cores(), 64)
        self.pool = cpu_port.StripPool(self.procs)

    def run(self, rows):
        """One round: `procs` strips of `rows` rows each (wrapping around the mesh); returns (elements, seconds)."""
        n = self.wl["n"]
        rows = max(1, min(rows, n))
        starts = [(k * rows) % max(1, n - rows + 1) for k in range(self.procs)]
        strips = [self.cp.strip_inputs(self.wl, st, rows, MAT) for st in starts]
        return self.pool.run(strips)

    def close(self):
        self.pool.close()


def reference_installed():
    """oracle/_ref holds the pip-installed, unmodified reference (oracle/make_ref.sh)."""
    return os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "FEMpy", "Problem.py")) or os.path.isdir("/root/reference/FEMpy")


# element rows of the workload mesh the reference assembles per timed pass: config 2 runs whole; the larger configs
# are sampled by a strip of full-width element rows so that a pass stays at a few seconds of CPU work (the reference's
# Python gathers and (EP, n, 2, n, 2) temporaries grow with E: config 3 whole needs 5.8 GB and 40 s, config 5 51 GB)
REF_SAMPLE_ROWS = {"c2": None, "c3": 80, "c4": 130, "c5": 52}


def run_ref_timing(config, reps, nx=None, ny=None, threads=None, breakdown=False, vm=False, timeout=900):
    """oracle/ref_timing.py in its own process (its Numba thread pool and caches stay out of this one)."""
    import subprocess

    cmd = [sys.executable, "-m", "oracle.ref_timing", "--config", config, "--reps", str(reps)]
    for flag, val in (("--nx", nx), ("--ny", ny), ("--threads", threads)):
        if val:
            cmd += [flag, str(val)]
    cmd += (["--breakdown"] if breakdown else []) + (["--vm"] if vm else [])
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    if res.returncode != 0:
        raise RuntimeError("oracle.ref_timing failed:\n" + res.stderr[-2000:])
    return json.loads(res.stdout.strip().splitlines()[-1])


def port_baseline(wl, budget_s=6.0):
    """The numpy port of the reference's stages on all host cores (a second, stronger CPU number)."""
    port = CpuPort(wl)
    rows = max(1, wl["n"] // 16)
    port.run(max(1, rows // 4))  # start the workers, warm numpy
    numEl, dt = port.run(rows)  # calibration
    rows = int(min(wl["n"], max(1, rows * budget_s / max(dt, 1e-3))))
    numEl, dt = port.run(rows)
    port.close()
    return {"value": numEl / dt, "unit": "elements/s", "cores": port.procs, "kind": "port",
            "sample": f"oracle/fem_oracle.py (numpy port of the reference's stages), {port.procs} worker processes each assembling "
                      f"its own strip of {rows} of the workload's {wl['n']} element rows ({numEl} elements in {dt:.2f} s wall)"}


def cpu_baseline(wl, config, nx=None):
    """The reference's own Numba path (FEMpyProblem._assembleMatrix + _assembleResidual, FEMpy/Problem.py:565-661) on
    the host cores: protocol of Examples/Benchmarks/benchmarkQuadMeshScaling.py:56-102 (one pass to compile, then the
    best of 3), all Numba threads, plus one single-thread pass and the stage breakdown; the numpy port beside it."""
    port = port_baseline(wl)
    if not reference_installed():
        port["note"] = "oracle/_ref absent (oracle/make_ref.sh needs /root/reference): numpy port only"
        return port
    ny = None if nx else REF_SAMPLE_ROWS.get(config)
    full = run_ref_timing(config, 3, nx=nx, ny=ny, breakdown=True)
    one = run_ref_timing(config, 1, nx=nx, ny=ny, threads=1)
    what = "whole workload mesh" if ny is None else f"strip of {ny} of the workload's {wl['n']} element rows"
    extra = {}
    if config == "c3":  # the configuration BASELINE.json quotes with stress recovery: the same pass + computeFunction("Von-Mises-Stress", "mean")
        vm = run_ref_timing(config, 2, nx=nx, ny=ny, vm=True)
        extra["with_von_mises_mean"] = {"seconds": vm["min_s"], "elements_per_s": vm["elements"] / vm["min_s"],
                                        "von_mises_seconds": vm["min_s"] - full["min_s"]}
    return {**extra, "value": full["elements"] / full["min_s"], "unit": "elements/s", "cores": full["numba_threads"], "kind": "reference",
            "sample": f"unmodified FEMpy 1.0.1 from oracle/_ref: _assembleMatrix + _assembleResidual (applyBCs) on the {what} "
                      f"({full['elements']} elements), one compile pass then best of 3 ({full['min_s']:.2f} s), numba "
                      f"{full['numba_threads']} threads ({full['threading_layer']}) on {full['cores']} host cores",
            "times_s": full["times_s"], "threading_layer": full["threading_layer"], "host_cores": full["cores"],
            "stage_breakdown_s": full.get("breakdown_s"), "first_pass_with_jit_s": full["first_pass_s"],
            "one_thread": {"value": one["elements"] / one["min_s"], "seconds": one["min_s"]},
            "residual_checksum": full["residual_checksum"], "K_abs_checksum": full["K_checksum"], "sampled_rows": ny,
            "port": port}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (rank 0 only): the
    unmodified FEMpy from oracle/_ref, every step one _assembleMatrix + _assembleResidual pass (FEMpy/Problem.py:565-661)
    with all Numba threads.  Config 2 is assembled whole every step; the larger workloads (and the N-GPU weak-scaling
    mesh, whose per-GPU share is config 2) by a bounded strip of full-width element rows, the metric being a rate.
    Without oracle/_ref the numpy port runs instead (kind "port")."""
    if rank != 0:
        return
    strong = args.gpus > 1 and args.config == "c5"
    config = "c5" if strong else (args.config if args.gpus == 1 else "c2")
    if not reference_installed():
        return run_reference_port(args, config)
    import contextlib
    import io

    sys.path.insert(0, ROOT)
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(ROOT, "oracle", "_ref", ".numba_cache"))
    from oracle import ref_timing

    ny = None if args.nx else REF_SAMPLE_ROWS.get(config)
    wl = make_workload(config, args.nx, ny)
    desc = WORKLOADS[config][2] + (f" [mesh overridden: {args.nx} x {args.nx} cells]" if args.nx else "")
    if args.gpus > 1 and not strong:
        desc = (f"config2 per GPU: structured Q4 {wl['n']}x{wl['n'] * args.gpus} = {args.gpus} x ({wl['n']}x{wl['n']}), "
                "iso plane stress, element-partitioned (weak scaling)")
    with contextlib.redirect_stdout(io.StringIO()):
        fp, model, prob = ref_timing.build_reference_problem(wl, MAT)
        ref_timing.one_pass(prob)  # every Numba kernel is compiled here
        for _ in range(max(0, min(args.warmup, 3) - 1)):
            ref_timing.one_pass(prob)
    import numba

    times = []
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(args.steps):
            t0 = time.perf_counter()
            K, R, _ = ref_timing.one_pass(prob)
            times.append(time.perf_counter() - t0)
    numEl = wl["numElems"]
    value = numEl * len(times) / sum(times)
    sample = (f"unmodified FEMpy 1.0.1 (oracle/_ref): _assembleMatrix + _assembleResidual, applyBCs, "
              + ("whole workload mesh" if ny is None else f"strip of {ny} full-width element rows of the workload mesh")
              + f" = {numEl} elements per step, numba {numba.get_num_threads()} threads ({numba.threading_layer()})")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "elements/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "elements_per_step": int(numEl), "sampled_rows": ny, "nnz_per_step": int(K.nnz)},
        "cpu_baseline": {"value": value, "unit": "elements/s", "cores": int(numba.get_num_threads()), "kind": "reference", "sample": sample,
                         "threading_layer": numba.threading_layer(), "host_cores": host_cores(), "best_step_s": min(times)},
        "e2e": {"value": value, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "residual_checksum": float(R.sum()),
    }  # fmt: skip
    print(json.dumps(line))


def run_reference_port(args, config):
    """Fallback of --impl reference when oracle/_ref is absent: the numpy port on all host cores."""
    strong = args.gpus > 1 and args.config == "c5"
    wl = make_workload(config, args.nx)
    if args.gpus > 1 and not strong:
        wl["desc"] = (f"config2 per GPU: structured Q4 {wl['n']}x{wl['n'] * args.gpus} = {args.gpus} x ({wl['n']}x{wl['n']}), "
                      "iso plane stress, element-partitioned (weak scaling)")
    port = CpuPort(wl, args.cpu_procs)
    rows = max(1, 32768 // wl["n"])
    port.run(max(1, rows // 4))  # starts the workers, warms numpy
    for _ in range(max(0, min(args.warmup, 3) - 1)):
        port.run(max(1, rows // 4))
    # bounded run: the timed steps together take about two minutes at most (the strips shrink for large --steps)
    _, dt = port.run(rows)
    if dt * args.steps > 120.0:
        rows = max(2, int(rows * 120.0 / (dt * args.steps)))
    total_el, total_t = 0, 0.0
    for _ in range(args.steps):
        numEl, dt = port.run(rows)
        total_el += numEl
        total_t += dt
    port.close()
    value = total_el / total_t
    sample = (f"{port.procs} worker processes, each assembling its own strip of {rows} of {wl['n']} element rows per step "
              f"({total_el // args.steps} elements per step), numpy port of the reference's stages")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "elements/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "sample_rows_per_worker": rows, "workers": port.procs},
        "cpu_baseline": {"value": value, "unit": "elements/s", "cores": port.procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line))


def ncu_traffic(config):
    """DRAM bytes per launch of the dominant kernel of a configuration, from the committed ncu --set full capture
    (profiles/r02_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep; None if absent)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            return json.load(f).get(config, {}).get("dram_bytes")
    except OSError:
        return None


def fp64_peak():
    """FP64 peak measured on a B200 of this pool with tools/micro/fp64_peak.cu (profiles/r02_fp64_peak.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_fp64_peak.json")) as f:
            return json.load(f)["fp64_tflops"]
    except (OSError, KeyError, ValueError):
        return None


# minimal FP64 operations per element (symmetric, B-sparse count of SURVEY.md section 8d)
MIN_FLOPS = {"quad": 1170.0, "quad9": 9900.0, "triangle": 240.0}


class _PluginProblem:
    """What bench.py needs of a load case to drive B200AssemblyMixin the way FEMpyProblem.updateJacobian /
    updateResidual do (FEMpy/Problem.py:239-264): the attributes SURVEY.md section 8b lists, nothing else."""

    def __init__(self, wl, el, cm):
        from types import SimpleNamespace

        elType = wl["elType"]
        self.constitutiveModel = cm
        self.model = SimpleNamespace(
            numNodes=wl["numNodes"], numDim=2, constitutiveModel=cm, nodeCoords=wl["coords"], dvs={"Thickness": wl["thick"]},
            elements={elType: {"elementObject": el, "connectivity": wl["conn"][elType], "numElements": wl["numElems"]}})
        self.numDOF = 2 * wl["numNodes"]
        self.states = wl["states"]
        loaded = np.flatnonzero(wl["loads"])
        self.loads = {"Load": {"DOF": loaded.tolist(), "Value": wl["loads"][loaded].tolist()}}
        self._bcs = {"Fixed": {"DOF": wl["bcDOF"].tolist(), "Value": wl["bcVal"].tolist()}}

    def getBCs(self):
        return self._bcs


def run_ours(args, rank, world, local_rank):
    import torch

    from fempy_b200 import AssemblyPlan, Constitutive, Elements

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: fempy_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        return run_ours_multi(args, rank, world, local_rank, dev)
    wl = make_workload(args.config, args.nx)
    elType = wl["elType"]
    el = {"quad": Elements.QuadElement2D(1), "quad9": Elements.QuadElement2D(2), "triangle": Elements.TriElement2D(1)}[elType]
    cm = Constitutive.IsoPlaneStress(MAT["E"], MAT["nu"], MAT["rho"], MAT["t"])
    mat = cm.material()

    plan = AssemblyPlan(wl["numNodes"], 2, 2, device=local_rank)
    N_, dN, w = el.tables()
    t0 = time.perf_counter()
    plan.setOption("assembly_path", {"scatter": 0, "rows": 2}[args.path])
    if args.balance_rows is not None:
        plan.setOption("balance_rows", args.balance_rows)
    if args.no_fan:
        plan.setOption("fan_rows", 0)
    for kv in args.plan_opt:
        name, val = kv.split("=")
        plan.setOption(name, int(val))
    plan.addBlock(N_, dN, w, wl["conn"][elType])
    plan.finalize()
    plan_wall_ms = 1e3 * (time.perf_counter() - t0)
    plan.setBCs(wl["bcDOF"], wl["bcVal"])
    nnz, numDOF, numEl = plan.nnz, plan.numDOF, plan.numElements

    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    plan.setStream(stream.cuda_stream)
    d_coords = torch.from_numpy(wl["coords"]).to(dev)
    d_states = torch.from_numpy(wl["states"]).to(dev)
    d_thick = torch.from_numpy(wl["thick"]).to(dev)
    d_loads = torch.from_numpy(wl["loads"]).to(dev)
    d_K = torch.empty(nnz, dtype=torch.float64, device=dev)
    d_R = torch.empty(numDOF, dtype=torch.float64, device=dev)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    flushRead = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)  # 256 MB, read after the write

    def flush_l2():
        """Evict the previous step's working set: write 512 MB (> L2), then read 256 MB of other data so the dirty
        lines of that write are written back BEFORE the timed region (otherwise the timed kernel pays for them)."""
        flush.zero_()
        if not args.flush_write_only:
            flushRead.sum()

    def step():
        plan.assembleDev(d_coords, d_states, d_thick, mat, not args.no_bcs, d_loads, d_K, d_R)

    for _ in range(args.warmup):
        flush_l2()
        step()
    torch.cuda.synchronize()
    launches_per_step = plan.lastLaunches

    sampler = ClockSampler(local_rank)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    torch.cuda.synchronize()
    sampler.start()
    for i in range(args.steps):
        flush_l2()  # evict the previous step's working set from L2 (untimed)
        sampler.sample()
        starts[i].record(stream)
        step()
        ends[i].record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    times = np.array([s.elapsed_time(e) for s, e in zip(starts, ends)])
    # second pass with the in-library events around the assembly kernel (roofline leg): kept out of the timed steps
    # above, where the two extra event records would sit inside the measured interval
    plan.profile(True)
    for i in range(args.steps):  # as many launches as the timed region had
        flush_l2()
        step()
    torch.cuda.synchronize()
    kernel_ms_sum, kernel_calls = plan.profileRead()
    plan.profile(False)
    ms_per_step = float(times.mean())
    value = numEl / (ms_per_step * 1e-3)

    # --- roofline of the dominant kernel -------------------------------------------------------------
    peak, peak_src = measured_peaks()
    alg_bytes = algorithmic_bytes(wl, nnz)
    kernel_ms = kernel_ms_sum / max(kernel_calls, 1)
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    nn = wl["conn"][elType].shape[1]
    if args.path == "scatter" or not plan.info("rows_path"):
        kernel = f"k_assemble_sym<{nn}> (fused K_e / R_e evaluation + RED.F64 scatter through the slot map)"
    elif plan.info("fan_path"):
        kernel = (f"k_assemble_fan<{nn}> (one thread per node row pair: corners walked as a fan, closed-form element rows, every 2x2 block "
                  "finished in registers and staged once, coalesced warp write-out; BCs, loads and residual fused)")
    else:
        kernel = f"k_assemble_rows<{nn}> (one thread per node row pair, accumulator rows in shared memory, coalesced warp write-out; BCs, loads, residual fused)"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(args.config) if (args.path == "rows" and args.nx is None and not args.no_fan) else None,
                "kernel": kernel, "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_bytes_per_element": alg_bytes / numEl,
                "peak_source": peak_src, "step_frac_of_roofline": (alg_bytes / (ms_per_step * 1e-3) / 1e9) / peak}
    f64 = fp64_peak()
    if f64:
        minTf = MIN_FLOPS[elType] * numEl / (kernel_ms * 1e-3) / 1e12
        roofline["fp64"] = {"peak_tflops": f64, "peak_source": "profiles/r02_fp64_peak.json (tools/micro/fp64_peak.cu on a B200 of this pool)",
                            "min_flops_per_element": MIN_FLOPS[elType], "achieved_tflops_min_count": minTf, "frac": minTf / f64}

    # --- stress recovery: von Mises, mean over the element's Gauss points, device-resident ---------
    d_vm = torch.empty(numEl, dtype=torch.float64, device=dev)
    for _ in range(3):
        plan.functionDev(d_coords, d_states, mat, 1, 2, d_vm)
    fs = [torch.cuda.Event(enable_timing=True) for _ in range(20)]
    fe = [torch.cuda.Event(enable_timing=True) for _ in range(20)]
    for i in range(20):
        flush_l2()
        fs[i].record(stream)
        plan.functionDev(d_coords, d_states, mat, 1, 2, d_vm)
        fe[i].record(stream)
    torch.cuda.synchronize()
    vm_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(fs, fe)]))
    vm_bytes = 4 * nn * numEl + 32 * wl["numNodes"] + 8 * numEl  # connectivity + coords + states read, one value per element written
    function = {"von_mises_mean": {"ms": vm_ms, "elements_per_s": numEl / (vm_ms * 1e-3), "algorithmic_bytes": vm_bytes,
                                   "frac_of_hbm_roofline": vm_bytes / (vm_ms * 1e-3) / 1e9 / peak, "checksum": float(d_vm.sum().item())}}

    # --- end to end through the host-pointer C ABI (pinned host buffers) ---------------------------
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64).pin_memory()
        t.numpy()[...] = a
        return t.numpy()

    h_coords, h_states, h_thick, h_loads = pinned(wl["coords"]), pinned(wl["states"]), pinned(wl["thick"]), pinned(wl["loads"])
    h_K, h_R = pinned(np.zeros(nnz)), pinned(np.zeros(numDOF))
    plan.setStream(None)
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        plan.assemble(h_coords, h_states, h_thick, mat, wl["bcDOF"], wl["bcVal"], True, h_loads, Kout=h_K, Rout=h_R)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.assemble(h_coords, h_states, h_thick, mat, wl["bcDOF"], wl["bcVal"], True, h_loads, Kout=h_K, Rout=h_R)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    checksum = float(h_R.sum())
    e2e = {"value": numEl / e2e_s, "unit": "elements/s", "ms_per_step": 1e3 * e2e_s,
           "h2d_bytes_per_step": int(h_coords.nbytes + h_states.nbytes + h_thick.nbytes + h_loads.nbytes),
           "d2h_bytes_per_step": int(h_K.nbytes + h_R.nbytes), "api": "femb_assemble (host pointers, pinned)",
           "residual_checksum": checksum}
    # the same call with the model's constants (node coordinates, thickness) resident on the device
    # (femb_plan_set_geometry): per step only states + loads go up -- what a Newton / optimisation loop pays
    plan.setGeometry(h_coords, h_thick)
    for _ in range(2):
        plan.assemble(None, h_states, None, mat, wl["bcDOF"], wl["bcVal"], True, h_loads, Kout=h_K, Rout=h_R)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.assemble(None, h_states, None, mat, wl["bcDOF"], wl["bcVal"], True, h_loads, Kout=h_K, Rout=h_R)
    res_s = (time.perf_counter() - t0) / e2e_steps
    assert float(h_R.sum()) == checksum
    e2e["resident_geometry"] = {"value": numEl / res_s, "ms_per_step": 1e3 * res_s,
                                "h2d_bytes_per_step": int(h_states.nbytes + h_loads.nbytes)}
    e2e["K_abs_checksum"] = float(np.abs(h_K).sum())
    e2e["residual_abs_checksum"] = float(np.abs(h_R).sum())
    # the plugin seam itself: B200AssemblyMixin._assembleMatrix + _assembleResidual as FEMpyProblem.updateJacobian /
    # updateResidual call them (FEMpy/Problem.py:239-264): pageable numpy arrays, a scipy csr_array built per call,
    # two kernel passes (K, then R)
    from fempy_b200.problem import B200AssemblyMixin

    plug = type("BenchProblem", (B200AssemblyMixin, _PluginProblem), {})(wl, el, cm)
    plug.model._b200_plan = plan
    plan.setGeometry(None, None)
    for _ in range(2):
        Kp, Rp = plug._assembleMatrix(plug.states, applyBCs=True), plug._assembleResidual(plug.states, applyBCs=True)
    t0 = time.perf_counter()
    psteps = max(3, min(args.steps, 10))
    for _ in range(psteps):
        Kp, Rp = plug._assembleMatrix(plug.states, applyBCs=True), plug._assembleResidual(plug.states, applyBCs=True)
    plug_s = (time.perf_counter() - t0) / psteps
    assert Kp.nnz == nnz and float(Rp.sum()) == checksum, "plugin path and C-ABI path disagree"
    e2e["plugin"] = {"value": numEl / plug_s, "ms_per_step": 1e3 * plug_s,
                     "api": "B200AssemblyMixin._assembleMatrix + _assembleResidual (pageable numpy in, scipy csr_array + ndarray out)"}

    line = {
        "metric": METRIC, "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": wl["desc"], "elements": numEl, "dof": numDOF, "nnz": nnz, "l2": "flushed between timed steps (512 MB write" + ("" if args.flush_write_only else " + 256 MB read so its dirty lines are written back before the timed region") + ")",
                   "states": "1e-3*rng(0)", "bcs": int(wl["bcDOF"].size)},
        "nnz_per_s": nnz / (ms_per_step * 1e-3), "ms_per_step_min": float(times.min()), "ms_per_step_median": float(np.median(times)),
        "ms_per_step_p90": float(np.percentile(times, 90)), "ms_per_step_max": float(times.max()),
        "roofline": roofline, "e2e": e2e, "function": function, "gpu_launches": int(launches_per_step * args.steps),
        "launches_per_step": int(launches_per_step), "clocks": clocks,
        "plan_build": {"device_ms": plan.buildMs, "wall_ms": plan_wall_ms},
        "path": {"assembly_path": args.path, "rows_path": plan.info("rows_path"), "max_degree": plan.info("max_degree"), "rows_balanced": plan.info("rows_balanced"),
                 "row_iters": [plan.info("row_iters_node_order"), plan.info("row_iters_balanced")], "closed_form_q4": plan.info("closed_form_q4"), "fan_path": plan.info("fan_path")},
    }  # fmt: skip
    if not args.no_cpu:
        cb = line["cpu_baseline"] = cpu_baseline(wl, args.config, args.nx)
        # the GPU result against the reference's own, where the reference assembled the same (whole) mesh in this run
        if cb.get("kind") == "reference" and cb.get("sampled_rows") is None and args.nx is None:
            dR = abs(cb["residual_checksum"] - checksum) / e2e["residual_abs_checksum"]
            dK = abs(cb["K_abs_checksum"] - e2e["K_abs_checksum"]) / cb["K_abs_checksum"]
            line["parity_vs_reference_in_this_run"] = {"residual_sum_rel_diff": dR, "K_abs_sum_rel_diff": dK, "tolerance": 1e-12}
            assert dR < 1e-12 and dK < 1e-12, f"GPU result differs from the reference's: R {dR:.2e}, K {dK:.2e}"
    print(json.dumps(line))



def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads to the CPUs NVML reports as closest to its GPU, so that the pinned staging buffers
    (first touch) land on that NUMA node and the per-step H2D / D2H copies do not all cross one socket.  Returns the
    resulting CPU set (or the reason nothing was done)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        before = sorted(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = sorted(os.sched_getaffinity(0))
        if not after:
            os.sched_setaffinity(0, before)
            return "unchanged (empty NVML set)"
        return f"{len(after)} cpus ({after[0]}-{after[-1]})"
    except Exception as exc:  # containers often hide the topology: keep the inherited affinity
        return f"unchanged ({type(exc).__name__})"


def run_ours_multi(args, rank, world, local_rank, dev):
    """N > 1: the mesh is element-partitioned over the ranks (fempy_b200/dist.py): every rank assembles the CSR row
    block of the nodes it owns, interface partial sums travel over NCCL.  Weak scaling: config 2's per-GPU share
    (512 x 512 quads) times the number of GPUs, i.e. a 512 x 512N strip mesh; `--config c5` runs the fixed
    50M-DOF mesh instead (strong scaling)."""
    import torch
    import torch.distributed as dist

    from fempy_b200 import Constitutive, Elements, meshes
    from fempy_b200.dist import DistributedAssembly, GpuLocalAssembler, Partition

    # stdout carries exactly one JSON line: NCCL's own "NCCL version ..." banner (printed on stdout at communicator
    # creation when NCCL_DEBUG=VERSION) is sent to stderr by pointing fd 1 at fd 2 until the first collective is through
    sys.stdout.flush()
    savedStdout = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(savedStdout, 1)
        os.close(savedStdout)
    affinity = bind_to_gpu_numa_node(local_rank)
    strong = args.config == "c5"
    if strong:
        nx = ny = args.nx or 5000
        desc = f"config5: structured Q4 {nx}x{ny}, iso plane stress, element-partitioned over {world} GPUs (strong scaling)"
    else:
        nx, ny = args.nx or 512, (args.nx or 512) * world
        desc = f"config2 per GPU: structured Q4 {nx}x{ny} = {world} x ({nx}x{nx}), iso plane stress, element-partitioned (weak scaling)"

    def measure(nx, ny, desc, strong, steps, warmup):
        """one partitioned workload: plan, device-timed steps (max over ranks), roofline leg, end-to-end leg; the JSON line
        on rank 0 (None elsewhere)"""
        coords, conn = meshes.quad4_grid(nx, ny)
        N = coords.shape[0]
        numEl = conn["quad"].shape[0]
        rng = np.random.default_rng(0)
        states = 1e-3 * rng.random((N, 2))
        left, right = meshes.edge_nodes(coords, x=0.0), meshes.edge_nodes(coords, x=2.0)
        bcDOF = np.stack((2 * left, 2 * left + 1), axis=1).reshape(-1).astype(np.int64)
        bcVal = np.zeros(bcDOF.size)
        cm = Constitutive.IsoPlaneStress(MAT["E"], MAT["nu"], MAT["rho"], MAT["t"])

        t0 = time.perf_counter()
        part = Partition(N, conn, rank, world, mode=args.dist_mode)
        asm = GpuLocalAssembler(part, {"quad": Elements.QuadElement2D(1).tables()}, coords, cm.material(), local_rank)
        da = DistributedAssembly(part, asm, dev)
        da.set_bcs(bcDOF, bcVal)
        plan_wall_ms = 1e3 * (time.perf_counter() - t0)

        stream = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(stream)
        own = part.owned_dof_slice()
        loadsLocal = np.zeros(2 * part.numLocalNodes)
        ownRight = right[(right >= part.n0) & (right < part.n1)]
        loadsLocal[2 * np.searchsorted(part.localNodes, ownRight)] += 1e4
        thickLocal = part.local_thickness(np.full(numEl, MAT["t"]), {"quad": 0})
        h_thick, h_states, h_loads = (torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (thickLocal, states[part.localNodes], loadsLocal))
        d_thick, d_states, d_loads = (h.to(dev) for h in (h_thick, h_states, h_loads))
        flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
        flushRead = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)

        def flush_l2():
            """same flush as the single-GPU arm: 512 MB write, then a 256 MB read so that the write's dirty lines are written
            back before the timed region"""
            flush.zero_()
            if not args.flush_write_only:
                flushRead.sum()

        def step():
            return da.assemble(d_thick, d_states, d_loads, True)

        for _ in range(warmup):
            flush_l2()
            step()
        torch.cuda.synchronize()
        launches_per_step = asm.plan.lastLaunches

        sampler = ClockSampler(local_rank)
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        torch.cuda.synchronize()
        dist.barrier()
        sampler.start()
        for i in range(steps):
            flush_l2()
            sampler.sample()
            starts[i].record(stream)
            step()
            ends[i].record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        clocks = sampler.stop()
        asm.plan.profile(True)  # roofline leg: in-library events, outside the timed steps
        for i in range(steps):
            flush_l2()
            step()
        torch.cuda.synchronize()
        kernel_ms_sum, kernel_calls = asm.plan.profileRead()
        asm.plan.profile(False)
        total_ms = torch.tensor([sum(s.elapsed_time(e) for s, e in zip(starts, ends))], dtype=torch.float64, device=dev)
        perRank = torch.zeros(world, 2, dtype=torch.float64, device=dev)  # every rank's own step / kernel time
        perRank[rank, 0] = total_ms[0] / steps
        perRank[rank, 1] = kernel_ms_sum / max(kernel_calls, 1)
        dist.all_reduce(perRank, op=dist.ReduceOp.SUM)
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)  # the slowest rank sets the pace
        ms_per_step = float(total_ms.item()) / steps
        value = numEl / (ms_per_step * 1e-3)

        # end to end: pinned host inputs -> device, assemble + exchange, owned row block -> pinned host, every step
        Kb, Rb = step()
        h_K = torch.empty(Kb.shape, dtype=torch.float64).pin_memory()
        h_R = torch.empty(Rb.shape, dtype=torch.float64).pin_memory()
        e2e_steps = max(3, min(steps, 20))
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            d_thick.copy_(h_thick, non_blocking=True)
            d_states.copy_(h_states, non_blocking=True)
            d_loads.copy_(h_loads, non_blocking=True)
            Kb, Rb = step()
            h_K.copy_(Kb, non_blocking=True)
            h_R.copy_(Rb, non_blocking=True)
            torch.cuda.synchronize()
        dist.barrier()
        e2e_t = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        counts = torch.tensor([h_thick.nbytes + h_states.nbytes + h_loads.nbytes, h_K.nbytes + h_R.nbytes, da.interface_bytes(),
                               int(part.numActiveElems), int(Kb.numel())], dtype=torch.float64, device=dev)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)

        if rank == 0:
            peak, peak_src = measured_peaks()
            nnzLocal = asm.plan.nnz
            localEl = asm.plan.numElements
            alg_bytes = 4 * 4 * localEl + 16 * part.numLocalNodes * 2 + 8 * localEl + 8 * nnzLocal + 16 * part.numLocalNodes
            kernel_ms = kernel_ms_sum / max(kernel_calls, 1)
            achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
            line = {
                "metric": METRIC, "value": value, "unit": "elements/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "elements": numEl, "dof": 2 * N, "nnz": int(counts[4].item()),
                           "l2": "flushed between timed steps (512 MB write" + ("" if args.flush_write_only else " + 256 MB read, as at N=1") + ")", "partition": "contiguous node ranges; element -> owner of its lowest node",
                           "collective": ("none: boundary elements are evaluated by every rank owning one of their nodes (owner-computes rows)" if args.dist_mode == "halo"
                                          else "NCCL batched send/recv of interface-row partial sums"), "interface_bytes_per_step": int(counts[2].item())},
                "nnz_per_s": counts[4].item() / (ms_per_step * 1e-3),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                             "kernel": "k_assemble_fan on rank 0's local mesh", "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": alg_bytes,
                             "peak_source": peak_src},
                "e2e": {"value": numEl / float(e2e_t.item()), "unit": "elements/s", "ms_per_step": 1e3 * float(e2e_t.item()),
                        "h2d_bytes_per_step": int(counts[0].item()), "d2h_bytes_per_step": int(counts[1].item()),
                        "api": "DistributedAssembly.assemble (pinned host inputs, owned row blocks back to pinned host)",
                        "host_gbs_aggregate": (counts[0].item() + counts[1].item()) / float(e2e_t.item()) / 1e9},
                "gpu_launches": int(launches_per_step * steps * world), "launches_per_step": int(launches_per_step), "clocks": clocks,
                "plan_build": {"wall_ms": plan_wall_ms, "device_ms": asm.plan.buildMs},
                "elements_evaluated": int(counts[3].item()),
                "per_rank": {"ms_per_step": [round(v, 5) for v in perRank[:, 0].tolist()], "kernel_ms": [round(v, 5) for v in perRank[:, 1].tolist()]},
                "host_affinity": affinity,
            }  # fmt: skip
            return line
        return None

    line = measure(nx, ny, desc, strong, args.steps, args.warmup)
    if not strong and args.nx is None and not args.no_strong:
        # the multi-GPU configuration BASELINE.json names (config 5, 50 M DOF) beside the weak-scaling line: same ranks,
        # fixed mesh (strong scaling), fewer steps
        import gc

        gc.collect()
        torch.cuda.empty_cache()
        c5 = measure(5000, 5000, f"config5: structured Q4 5000x5000, iso plane stress, element-partitioned over {world} GPUs (strong scaling)",
                     True, max(3, min(args.steps, 20)), 3)
        if rank == 0:
            line["config5_strong"] = {k: c5[k] for k in ("value", "unit", "ms_per_step", "scaling", "config", "nnz_per_s", "roofline", "e2e", "per_rank",
                                                       "elements_evaluated", "plan_build")}
    if rank == 0:
        print(json.dumps(line))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--nx", type=int, default=None, help="override the mesh size (exploration only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--balance-rows", type=int, default=None, help="row schedule of the row-owner kernel: 0 node order, 1 balanced when it pays (default), 2 always")
    ap.add_argument("--cpu-procs", type=int, default=None, help="worker processes of the CPU legs (default: all host cores)")
    ap.add_argument("--flush-write-only", action="store_true", help="L2 flush by the 512 MB write alone (leaves dirty lines for the timed kernel to evict)")
    ap.add_argument("--dist-mode", default="halo", choices=["halo", "exchange"], help="multi-GPU: redundant boundary elements (no exchange) or unique elements + NCCL interface exchange")
    ap.add_argument("--no-fan", action="store_true", help="accumulating row-owner kernel instead of the write-once fan kernel (exploration only)")
    ap.add_argument("--plan-opt", action="append", default=[], metavar="NAME=INT", help="femb_plan_set_option before finalize (exploration only)")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the config-5 strong-scaling leg of the default line")
    ap.add_argument("--no-bcs", action="store_true", help="device-timed steps without boundary conditions (exploration only)")
    ap.add_argument("--path", default="rows", choices=["rows", "scatter"],
                    help="assembly path: row-owner kernel (default) or slot-map scatter with FP64 atomics")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
