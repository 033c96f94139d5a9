#!/usr/bin/env python
"""bench.py -- headline benchmark: dense fp64 lasso solved by FASTA's forward-backward splitting.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): lasso FBS iterations/sec (fp64) on config 2, dense lasso M=40000 N=100000
(32 GB A), row-sharded over the N GPUs of one box (strong scaling: the problem is fixed).

A "step" is ONE full solve to tolerance 1e-5 (adaptive mode, hybrid stop rule, verbose off)
through the public API ``fasta.fasta(A, loss.f, loss.gradf, pen.g, pen.prox, x0)``.
  value   = iterations performed in the K timed solves / device time of the K solves, with A, b,
            x0 already resident in HBM (max over ranks, barrier + synchronize on both sides).
            The time includes each solve's prologue (Lipschitz estimate + initial gradient = 6
            passes over A), so it is a lower bound on the in-loop rate (also reported).
  e2e     = the same metric when every step also uploads A, b, x0 from pinned host memory and
            downloads the solution (the call a user with host arrays makes).
  roofline= dominant kernel, timed with CUDA events around every launch DURING the timed solves:
            the single-pass dense_sweep_kernel (z = A x, loss, g = A^T r in one read of A;
            algorithmic bytes = the reference's two contractions = 2*M*N*8 per launch, DRAM traffic
            = M*N*8), or the two-pass dense_stream_kernel (M*N*8 per launch) when the sweep is off.
  cpu_baseline = the numpy oracle (oracle/fasta_oracle.py, a port of the reference loop; numpy's
            OpenBLAS dgemv is the same arithmetic the reference runs) on the host cores, on a
            bounded sample (first iterations of the same problem).
``--impl reference`` prints the CPU arm alone (rank 0 only under torchrun).
"""

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "fasta-python_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: M, N, K (non-zeros of the true signal), sigma, mu, row chunks for seeding
    "lasso_40000x100000": dict(M=40000, N=100000, K=5000, sigma=0.01, mu=0.02, chunks=64),
    "lasso_8000x20000": dict(M=8000, N=20000, K=1000, sigma=0.01, mu=0.02, chunks=64),   # dev / small boxes
}
SOLVER_OPTS = dict(adaptive=True, accelerate=False, verbose=False, tolerance=1e-5, max_iters=1000)
CPU_SAMPLE_ITERS = 3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lasso_40000x100000", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--debug-phases", action="store_true", help="print host-side phase times of every timed solve to stderr")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# synthetic data: identical for every world size (seeded per fixed row chunk)
# --------------------------------------------------------------------------------------------------
def make_local_problem(w, rank, world, device):
    """Rows [lo, hi) of the config-2 recipe: A = randn / (sqrt(M)+sqrt(N)),
    x_true with K ones, b = A x_true + sigma * randn.  Generated on the device."""
    import torch
    M, N, K = w["M"], w["N"], w["K"]
    assert M % w["chunks"] == 0 and w["chunks"] % world == 0
    rows_per_chunk = M // w["chunks"]
    lo, hi = (M * rank) // world, (M * (rank + 1)) // world
    A = torch.empty(hi - lo, N, dtype=torch.float64, device=device)
    noise = torch.empty(hi - lo, dtype=torch.float64, device=device)
    gen = torch.Generator(device=device)
    for c in range(lo // rows_per_chunk, hi // rows_per_chunk):
        gen.manual_seed(1000 + c)
        r0 = c * rows_per_chunk - lo
        A[r0:r0 + rows_per_chunk].normal_(generator=gen)
        noise[r0:r0 + rows_per_chunk].normal_(generator=gen)
    A /= (np.sqrt(M) + np.sqrt(N))
    cpu_gen = torch.Generator().manual_seed(7)
    support = torch.randperm(N, generator=cpu_gen)[:K]
    x_true = torch.zeros(N, dtype=torch.float64)
    x_true[support] = 1.0
    x_true = x_true.to(device)
    b = torch.mv(A, x_true) + w["sigma"] * noise      # setup only, untimed
    return A, b, x_true


def clocks_sampler():
    """nvidia-smi clocks line of the profiling recipe, sampled every 200 ms in the background."""
    try:
        f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                             stdout=f, stderr=subprocess.DEVNULL)
        return p, f
    except Exception:
        return None, None


def _stamp(text):
    from datetime import datetime
    try:
        return datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
    except ValueError:
        return None


def clocks_summary(proc, f, device_index, t_begin, t_end):
    """Median SM clock / throttle reasons over the samples taken INSIDE [t_begin, t_end]."""
    out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    if proc is None:
        return out
    proc.terminate()
    try:
        proc.wait(timeout=5)
    except Exception:
        proc.kill()
    f.flush()
    f.seek(0)
    sm, smmax, power, reasons = [], [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for line in f.read().splitlines():
        parts = [p.strip() for p in line.split(",")]
        if len(parts) < 10:
            continue
        ts = _stamp(parts[0])
        try:
            if int(parts[1]) != device_index or ts is None or ts < t_begin - 0.2 or ts > t_end + 0.2:
                continue
            sm.append(float(parts[2]))
            smmax.append(float(parts[3]))
            power.append(float(parts[4]))
        except ValueError:
            continue
        for name, val in zip(names, parts[6:10]):
            if val.lower().startswith("active"):
                reasons.add(name)
    f.close()
    try:
        os.unlink(f.name)
    except OSError:
        pass
    if sm:
        # median over samples taken under load (the upper half of the observed clocks)
        out["sm_mhz"] = float(np.median(sm))
        out["sm_max_mhz"] = float(max(smmax))
        out["samples"] = len(sm)
        out["power_w_max"] = float(max(power))
    out["reasons"] = sorted(reasons)
    return out


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_arm(w, A_host, b_host, sample_iters=CPU_SAMPLE_ITERS, gpu_solve=None):
    """Bounded sample: Lipschitz prologue + `sample_iters` iterations of the same problem with numpy
    (all host threads).  Rate = iterations / in-loop time, the reference's own definition
    (examples/__init__.py:59-60), which EXCLUDES the prologue -- the most favourable reading."""
    from oracle import fasta_oracle
    try:
        from threadpoolctl import threadpool_info
        pools = [f"{p.get('internal_api')}:{p.get('num_threads')}" for p in threadpool_info()]
    except Exception:
        pools = []
    mu = w["mu"]
    la = np.linalg
    f = lambda z: .5 * la.norm((z - b_host).ravel()) ** 2
    gradf = lambda z: z - b_host
    g = lambda x: mu * la.norm(x.ravel(), 1)
    proxg = lambda x, t: fasta_oracle.shrink(x, t * mu)
    op = lambda x: A_host @ x
    adj = lambda y: A_host.T @ y
    np.random.seed(0)
    t0 = time.time()
    opts = dict(SOLVER_OPTS)
    opts["max_iters"] = sample_iters
    res = fasta_oracle.solve(op, adj, f, gradf, g, proxg, np.zeros(w["N"]), **opts)
    wall = time.time() - t0
    n = res.iteration_count
    loop = res.times[n] - res.times[0]
    cores = len(os.sched_getaffinity(0))
    out = dict(value=n / loop, unit="iterations/s", cores=cores, kind="port",
               sample=(f"oracle/fasta_oracle.py (numpy {np.__version__}, BLAS pools {pools}) on the full "
                       f"{w['M']}x{w['N']} problem, first {n} iterations (in-loop time {loop:.2f} s; whole call "
                       f"incl. 6-pass prologue {wall:.2f} s); os.cpu_count()={os.cpu_count()}"))
    if gpu_solve is not None:
        # the oracle as the checker (not the thing measured): the same first iterations on the GPU, same seed
        try:
            np.random.seed(0)
            got = gpu_solve(opts)
            rel = lambda a, b_: float(np.max(np.abs(np.asarray(a) - np.asarray(b_)) / np.maximum(np.abs(np.asarray(b_)), 1e-300)))
            sol = got.solution.detach().cpu().numpy() if hasattr(got.solution, "detach") else np.asarray(got.solution)
            out["parity_full_size"] = dict(
                iterations=[int(got.iteration_count), int(n)], backtracks=[int(got.backtracks), int(res.backtracks)],
                stepsizes_rel_err=rel(got.stepsizes[:n], res.stepsizes[:n]),
                residuals_rel_err=rel(got.residuals[:n], res.residuals[:n]),
                iterate_rel_err=float(la.norm(sol - res.solution) / max(la.norm(res.solution), 1e-300)))
        except Exception as exc:                             # never lose the bench line over the cross-check
            out["parity_full_size"] = dict(error=repr(exc))
    return out


def reference_main(args):
    """--impl reference: time the reference's CPU path (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    w = WORKLOADS[args.workload]
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if torch.cuda.is_available() else torch.device("cpu")
    A, b, _ = make_local_problem(w, 0, 1, dev)
    A_host, b_host = A.cpu().numpy(), b.cpu().numpy()
    del A
    steps = max(1, args.steps)
    vals = []
    for _ in range(min(args.warmup, 1)):
        cpu_arm(w, A_host, b_host, 1)
    for _ in range(min(steps, 2)):
        base = cpu_arm(w, A_host, b_host)
        vals.append(base["value"])
    base["value"] = float(np.mean(vals))
    line = dict(impl="reference", metric="lasso_fbs_iterations_per_sec", value=base["value"], unit="iterations/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 / base["value"],
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                config=config_of(w, args.gpus), cpu_baseline=base,
                e2e=dict(value=base["value"], unit="iterations/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def config_of(w, world):
    return dict(workload=f"dense lasso M={w['M']} N={w['N']} K={w['K']} fp64 ({w['M'] * w['N'] * 8 / 1e9:.1f} GB A), "
                         f"adaptive FASTA to tolerance 1e-5, A row-sharded over {world} GPU(s)",
                step="one full solve (Lipschitz prologue + iterations to tolerance)",
                l2="inputs (A) are far larger than the 126 MB L2; no flush needed",
                sigma=w["sigma"], mu=w["mu"], scaling_recipe="A = randn/(sqrt(M)+sqrt(N)) (SURVEY.md 8d)")


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        reference_main(args)
        return

    import torch
    import torch.distributed as dist
    import fasta
    from fasta import _backends

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run --nproc-per-node N")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    w = WORKLOADS[args.workload]
    M, N = w["M"], w["N"]
    A, b, _ = make_local_problem(w, rank, world, device)
    m_local = A.shape[0]
    x0 = torch.zeros(N, dtype=torch.float64, device=device)

    # ---- operator / loss / penalty (public API objects) -------------------------------------------
    def build_objects(A_dev, b_dev):
        op = fasta.distributed.RowShardedMatrix(A_dev) if world > 1 else fasta.linalg.LinearMap.from_matrix(A_dev)
        return op, fasta.losses.LeastSquares(b_dev), fasta.proximal.L1Norm(w["mu"])

    op, loss, pen = build_objects(A, b)

    # live kernel timing: CUDA events around every dense contraction launched in the timed region
    kernel_events = []
    orig_forward, orig_adjoint = _backends.DenseDriver.forward, _backends.DenseDriver.adjoint
    orig_sweep = _backends.DenseDriver.sweep
    record = {"on": False}
    sweep_events = []

    def timed(fn, sink):
        def wrapper(self, *a, **k):
            if not record["on"]:
                return fn(self, *a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(self, *a, **k)
            e1.record()
            sink.append((e0, e1))
            return out
        return wrapper

    _backends.DenseDriver.forward = timed(orig_forward, kernel_events)
    _backends.DenseDriver.adjoint = timed(orig_adjoint, kernel_events)
    _backends.DenseDriver.sweep = timed(orig_sweep, sweep_events)

    marks = []
    if args.debug_phases:
        def phase(cls, name):
            orig = getattr(cls, name)

            def f(self, *a, **k):
                t0 = time.perf_counter()
                out = orig(self, *a, **k)
                marks.append((name, time.perf_counter() - t0))
                return out
            setattr(cls, name, f)
        for name in ("__init__", "load", "lipschitz", "start", "solution", "close"):
            phase(_backends.FusedBackend, name)

    def solve(xstart):
        np.random.seed(0)          # identical tau0 probes in every step and on every rank
        t0 = time.perf_counter()
        res = fasta.fasta(op, loss.f, loss.gradf, pen.g, pen.prox, xstart, **SOLVER_OPTS)
        if args.debug_phases and rank == 0:
            loop = res.times[res.iteration_count] - res.times[0]
            print(f"[phases] call {1e3 * (time.perf_counter() - t0):.1f} ms, loop {1e3 * loop:.1f} ms :: " +
                  ", ".join(f"{n} {1e3 * t:.2f}" for n, t in marks), file=sys.stderr, flush=True)
            marks.clear()
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(fn, steps):
        """K steps bracketed by barrier + synchronize; device time by CUDA events; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        outs = []
        for _ in range(steps):
            if outs:
                outs[-1].solution = None     # drop the previous solution buffer, as a caller's loop would: the caching
            outs.append(fn())                # allocator then stays in steady state (no cudaMalloc inside the timed region)
        e1.record()
        barrier()
        wall = time.time() - t0
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return outs, float(ms.item()), wall

    # ---- warm-up, then the resident-input measurement ---------------------------------------------
    sampler, sfile = clocks_sampler() if rank == 0 else (None, None)
    for _ in range(max(args.warmup, 3)):
        res = solve(x0)
    record["on"] = True
    t_begin = time.time()
    outs, ms_total, wall = timed_region(lambda: solve(x0), args.steps)
    t_end = time.time()
    record["on"] = False
    torch.cuda.synchronize()
    clocks = clocks_summary(sampler, sfile, local, t_begin, t_end) if rank == 0 else None

    iters = sum(r.iteration_count for r in outs)
    peer_reductions = sum(getattr(r, "peer_reductions", 0) for r in outs)
    backtracks = sum(r.backtracks for r in outs)
    launches = sum(r.kernel_launches for r in outs)
    loop_s = sum(r.times[r.iteration_count] - r.times[0] for r in outs)
    value = iters / (ms_total / 1e3)
    single_pass = len(sweep_events) > 0
    kern_ms_all = [a.elapsed_time(b_) for a, b_ in (sweep_events if single_pass else kernel_events)]
    # a speculative launch whose predecessor was rejected / final returns at once (fb200_trial_decide): not a pass over A
    kern_ms = [t for t in kern_ms_all if t > 0.2 * float(np.median(kern_ms_all))]
    kern_avg_ms = float(np.mean(kern_ms))
    dram_bytes = m_local * N * 8                       # one read of the local rows of A
    alg_bytes = (2 if single_pass else 1) * dram_bytes  # reference contractions covered by one launch
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (kern_avg_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath))
        if t.get("workload") == args.workload and world == 1:
            traffic = t.get("sweep_dram_bytes_per_launch" if single_pass else "dram_bytes_per_launch")
    kernel_name = ("dense_sweep_kernel (single pass: z = A x, loss, g = A^T r; one launch per iteration, covers the "
                   "reference's two contractions = 2*M*N*8 algorithmic bytes while reading A from HBM once)"
                   if single_pass else "dense_stream_kernel (A x and A^T r, one launch each per iteration)")
    roofline = dict(bound="hbm", kernel=kernel_name,
                    achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                    peak_source=peak_src, algorithmic_bytes_per_launch=alg_bytes,
                    avg_launch_ms=kern_avg_ms, launches_timed=len(kern_ms),
                    speculative_launches_returned_at_once=len(kern_ms_all) - len(kern_ms),
                    frac_of_nominal_8TBs=achieved / 8000.0,
                    dram_GBs=dram_bytes / (kern_avg_ms * 1e-3) / 1e9,
                    dram_frac_of_measured_peak=dram_bytes / (kern_avg_ms * 1e-3) / 1e9 / peak,
                    whole_iteration_algorithmic_GBs=(2 * iters + backtracks) * dram_bytes / loop_s / 1e9,
                    whole_iteration_frac_of_nominal_8TBs=(2 * iters + backtracks) * dram_bytes / loop_s / 1e9 / 8000.0)

    # ---- end to end: host buffers in, host result out ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        pinned = True
        try:
            A_host = torch.empty(A.shape, dtype=torch.float64, pin_memory=True)
        except Exception:
            A_host = torch.empty(A.shape, dtype=torch.float64)
            pinned = False
        A_host.copy_(A)
        b_host = b.cpu().pin_memory()
        x0_host = torch.zeros(N, dtype=torch.float64).pin_memory()
        sol_host = torch.empty(N, dtype=torch.float64).pin_memory()
        torch.cuda.synchronize()

        def e2e_step():
            A.copy_(A_host, non_blocking=True)          # reuse the resident buffers as upload targets
            b_dev = b_host.to(device, non_blocking=True)
            x_dev = x0_host.to(device, non_blocking=True)
            op_, loss_, pen_ = build_objects(A, b_dev)
            np.random.seed(0)
            r = fasta.fasta(op_, loss_.f, loss_.gradf, pen_.g, pen_.prox, x_dev, **SOLVER_OPTS)
            sol_host.copy_(r.solution, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return r

        e2e_step()
        e_outs, e_ms, _ = timed_region(e2e_step, args.steps)
        e_iters = sum(r.iteration_count for r in e_outs)
        e2e = dict(value=e_iters / (e_ms / 1e3), unit="iterations/s",
                   h2d_bytes_per_step=int((A.numel() + b.numel() + N) * 8 * world),
                   d2h_bytes_per_step=int(N * 8 * world), ms_per_step=e_ms / args.steps, pinned_host=pinned)

    # ---- CPU baseline on the same problem (rank 0, N=1 only) --------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        A_np = A_host.numpy() if e2e is not None else A.cpu().numpy()
        cpu = cpu_arm(w, A_np, b.cpu().numpy(), gpu_solve=lambda o: fasta.fasta(op, loss.f, loss.gradf, pen.g, pen.prox, x0, **o))

    if rank == 0:
        line = dict(metric="lasso_fbs_iterations_per_sec", value=value, unit="iterations/s", n_gpus=world,
                    steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_total / args.steps,
                    higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                    config=config_of(w, world), roofline=roofline, cpu_baseline=cpu, e2e=e2e,
                    gpu_launches=int(launches), clocks=clocks,
                    iterations_per_solve=iters / args.steps, backtracks_per_solve=backtracks / args.steps,
                    time_to_tol_ms=1e3 * loop_s / args.steps, iters_per_sec_in_loop=iters / loop_s,
                    wall_ms_per_step=1e3 * wall / args.steps)
        line["final_residual"] = float(outs[-1].residuals[outs[-1].iteration_count - 1])
        line["speculation"] = getattr(outs[-1], "speculation", None)
        if world > 1:
            line["collective"] = ("fused peer-memory all-reduce + BB epilogue kernel over NVLink (fb200_peer_allreduce_bb), "
                                  f"{peer_reductions} calls in the timed region" if peer_reductions else "ncclAllReduce + bb kernel")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
