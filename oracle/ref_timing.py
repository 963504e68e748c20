"""Timed CPU baseline: the UNMODIFIED reference's own assembly drivers on the host cores.

TEST / BENCH INFRASTRUCTURE -- never imported by fempy_b200.  Used by bench.py (`cpu_baseline` leg and `--impl reference`).

Runs FEMpyProblem._assembleMatrix(states, applyBCs=True) + _assembleResidual(states, applyBCs=True)
(reference FEMpy/Problem.py:565-661) -- and computeFunction("Von-Mises-Stress", "mean") on request -- on a BASELINE
workload mesh with the protocol of the reference's own benchmark (Examples/Benchmarks/benchmarkQuadMeshScaling.py:56-102):
one full pass first so that every Numba kernel is compiled, then timed repetitions.  The reference comes from
oracle/_ref (installed by oracle/make_ref.sh; /root/reference itself when that is present), its un-installed
dependencies from oracle/stubs.

  python -m oracle.ref_timing --config c2 --reps 3 [--threads N] [--breakdown] [--vm]     -> one JSON line on stdout
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_reference_problem(wl, mat):
    """The reference's FEMpyModel / FEMpyProblem for a bench.py workload dict (plane stress, BCs, loads, states)."""
    from oracle import ref_harness as rh

    fp = rh.import_reference()
    left = wl["bcDOF"][::2] // 2
    right = (wl["loads"].reshape(-1, 2)[:, 0] != 0).nonzero()[0]
    model, prob = rh.build_problem(fp, wl["coords"], {wl["elType"]: wl["conn"][wl["elType"]]}, 0, mat["E"], mat["nu"], mat["rho"], mat["t"],
                                   thickness=wl["thick"], bcs=[("Fixed", left, [0, 1], 0.0)], loads=[("Load", right, [0], 1e4, False)],
                                   states=wl["states"])
    return fp, model, prob


def one_pass(prob, vm=False):
    K = prob._assembleMatrix(prob.states, applyBCs=True)
    R = prob._assembleResidual(prob.states, applyBCs=True)
    f = prob.computeFunction("Von-Mises-Stress", "mean") if vm else None
    return K, R, f


def stage_breakdown(fp, model, prob):
    """One instrumented K + R pass: seconds spent in the stages of SURVEY.md section 6 (timers wrapped around the
    reference's own functions from outside; the reference is not edited)."""
    import FEMpy.Problem as P
    from FEMpy.Utils import AssemblyUtils as AU

    acc = {}

    def timed(name, fn):
        def w(*a, **k):
            t0 = time.perf_counter()
            try:
                return fn(*a, **k)
            finally:
                acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
        return w

    saved = []

    def patch(obj, attr, name):
        saved.append((obj, attr, getattr(obj, attr)))
        setattr(obj, attr, timed(name, getattr(obj, attr)))

    patch(model, "getElementCoordinates", "python_gather")
    patch(model, "getElementDVs", "python_gather")
    patch(prob, "getElementStates", "python_gather")
    for elData in model.elements.values():
        patch(elData["elementObject"], "computeResidualJacobians", "numba_kernels")
        patch(elData["elementObject"], "computeResiduals", "numba_kernels")
    patch(AU, "localMatricesToCOOArrays", "coo_expand")
    patch(AU, "scatterLocalResiduals", "residual_scatter")
    patch(AU, "applyBCsToMatrix", "bc")
    patch(AU, "applyBCsToVector", "bc")
    patch(P, "coo_array", "scipy_coo_to_csr")
    patch(P, "csr_array", "scipy_coo_to_csr")
    t0 = time.perf_counter()
    try:
        one_pass(prob)
    finally:
        for obj, attr, fn in reversed(saved):
            setattr(obj, attr, fn)
    acc["total"] = time.perf_counter() - t0
    acc["other"] = acc["total"] - sum(v for k, v in acc.items() if k != "total")
    return acc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--nx", type=int, default=None)
    ap.add_argument("--ny", type=int, default=None, help="element rows (default nx): a strip of the workload")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--warm", type=int, default=1)
    ap.add_argument("--threads", type=int, default=None)
    ap.add_argument("--breakdown", action="store_true")
    ap.add_argument("--vm", action="store_true", help="add computeFunction('Von-Mises-Stress', 'mean') to every pass")
    args = ap.parse_args()
    if args.threads:
        os.environ["NUMBA_NUM_THREADS"] = str(args.threads)
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(ROOT, "oracle", "_ref", ".numba_cache"))
    sys.path.insert(0, ROOT)
    import contextlib
    import io

    import numba

    import bench

    wl = bench.make_workload(args.config, args.nx, args.ny)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        fp, model, prob = build_reference_problem(wl, bench.MAT)
    setup_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(max(1, args.warm)):
            K, R, f = one_pass(prob, args.vm)  # JIT compilation happens here
    warm_s = time.perf_counter() - t0
    times = []
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(args.reps):
            t0 = time.perf_counter()
            K, R, f = one_pass(prob, args.vm)
            times.append(time.perf_counter() - t0)
    out = {
        "elements": int(wl["numElems"]), "dof": int(2 * wl["numNodes"]), "nnz": int(K.nnz), "times_s": times, "min_s": min(times),
        "mean_s": sum(times) / len(times), "setup_s": setup_s, "first_pass_s": warm_s, "numba_threads": int(numba.get_num_threads()),
        "threading_layer": numba.threading_layer(), "cores": len(os.sched_getaffinity(0)), "workload": wl["desc"],
        "residual_checksum": float(R.sum()), "K_checksum": float(abs(K.data).sum()), "vm": bool(args.vm),
        "reference": os.path.dirname(fp.__file__),
    }  # fmt: skip
    if f is not None:
        out["vm_checksum"] = float(sum(v.sum() for v in f.values()))
    if args.breakdown:
        with contextlib.redirect_stdout(io.StringIO()):
            out["breakdown_s"] = stage_breakdown(fp, model, prob)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
