#!/usr/bin/env python
"""bench.py -- Gibbs sweeps/s and datum-component log_post_pred evals/s of the B200 engine (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4|c5] [--impl ours|reference]

A *step* is one Gibbs sweep (one pass of crpmm.py:57-88 / pcrpmm.py:93-131 over all N data).  Default workload
= BASELINE.json's target configuration C3: PCRPMM, NIW full covariance, N=1e6, D=16, K_true=100, r=1.5, random scan
(SURVEY.md 8d), synthetic data from the demos' generator, `rand` initial assignments with K=K_true.

Protocol ("rand init, one warm-up sweep, time sweeps 2..S"): the W warm-up steps are W sweeps of the
chain from the rand initial state (the chain's cold sweeps: they warm the device and exercise every engine regime);
the chain is then put back to the SAME initial state, sweep 0 runs untimed (8d's warm-up sweep), and EXACTLY K sweeps
-- sweeps 1..K of the chain, cold sweeps included -- are timed.  Nothing converged is hidden in the warm-up: the timed
region is the run a user pays for.  `regimes` splits the timed sweeps by their share of movers.

Prints ONE JSON line (rank 0).  `value` = whole-job evals/s with the per-step inputs (scan order + uniforms)
already resident in HBM; `e2e` = the same K sweeps replayed from the same state through the host C-ABI call
(bgmm_sweep) with pinned HOST buffers -> H2D of order+uniforms and D2H of the assignments inside the timed region.
Under torchrun (N>1) every rank runs its own chain: independent chains (seed 1+rank) on the same data for c2/c3/c4
(BASELINE.json configs[3]), disjoint 1e6-row shards of one 8e6-row data set for c5 (configs[4]); weak scaling, no
data-path collective; the assignments are all-gathered with NCCL after the timed region.  `--replicas` makes every
rank run the same seeded chain instead (identical per-GPU work).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (sampler, N, D, K_true, power, cov)                                   BASELINE.json configs[i]
    "c2": ("CRPMM", 100000, 2, 30, 1.0, "full"),      # configs[1]
    "c3": ("PCRPMM", 1000000, 16, 100, 1.5, "full"),  # configs[2]  <- north_star target, default
    "c4": ("CRPMM", 1000000, 64, 100, 1.0, "full"),   # configs[3] (one chain per GPU)
    "c5": ("PCRPMM", 1000000, 8, 100, 1.5, "full"),   # configs[4] (one 1e6-row shard per GPU)
    "tiny": ("PCRPMM", 20000, 16, 20, 1.5, "full"),   # smoke-sized
}
METRIC = "Gibbs sweeps/sec (N x K log_post_pred evals)"
UNIT = "evals/s"


def gen_data(N, D, K_true, seed, shard=0, chain_seed=None):
    """examples/crpmm_2d_demo.py:41-55 scaled.  The component means depend on `seed` alone; the rows of
    shard `shard` come from their own stream, so shards 0..G-1 are contiguous row blocks of one G*N-row data set.  The
    initial assignments (igmm.py:86-94 "rand") come from `chain_seed` (default: seed)."""
    mu = np.random.RandomState(seed).randn(D, K_true) * 4.0
    rs = np.random.RandomState([seed, 104729 + shard])
    z_true = rs.randint(0, K_true, N)
    X = np.ascontiguousarray((mu[:, z_true] + rs.randn(D, N) * 0.7).T)
    rz = np.random.RandomState([seed if chain_seed is None else chain_seed, 15485863])
    drawn = rz.randint(0, K_true, N)
    z0 = np.unique(drawn, return_inverse=True)[1].astype(np.int64)   # consecutive labels (igmm.py:89-94)
    return X, z_true, z0


def prior_for(D, cov):
    v_0 = D + 3
    return np.zeros(D), 0.7 ** 2 / 4.0 ** 2, v_0, 0.7 ** 2 * v_0 * (np.eye(D) if cov == "full" else np.ones(D))


def step_input(N, s, power, seed):
    """Scan order (pcrpmm.py:89) and uniforms (utils.py:15) of sweep `s` of chain `seed`, generated on the host like
    the reference does; a function of (seed, s) only, so sweep s is the same whatever --steps / --warmup are."""
    rs = np.random.RandomState([seed, 7919, s])
    order = rs.permutation(N).astype(np.int64) if power > 1 else None
    return order, rs.random_sample(N)


def step_inputs(N, n_steps, power, seed):
    pairs = [step_input(N, s, power, seed) for s in range(n_steps)]
    return [o for o, _ in pairs], [u for _, u in pairs]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super(ClockSampler, self).__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([time.time()] + [c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self, t0=None, t1=None):
        """Summary over the samples taken in [t0, t1] (the timed region); the sampler itself runs from the start of
        the warm-up sweeps so that a short timed region still sees the clocks the device was running at under load."""
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)

        def summarise(rows):
            sm, mx, reasons = [], [], set()
            for r in rows:
                try:
                    sm.append(float(r[2])); mx.append(float(r[3]))
                except Exception:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[6:10]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons
        timed = [r for r in self.rows if t0 is not None and t0 - 0.05 <= r[0] <= t1 + 0.05]
        sm_t, mx_t, re_t = summarise(timed)
        sm_a, mx_a, re_a = summarise(self.rows)
        sm, mx = (sm_t, mx_t) if sm_t else (sm_a, mx_a)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(re_t | re_a), "samples": len(sm_t), "samples_incl_warmup_under_load": len(sm_a)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


FP64_PEAK_TFLOPS = 40.0   # nominal vector FP64 of one B200 (no measured figure in MEASURED_PEAKS.json)


def ncu_profile(workload, regime, N):
    """Per-sweep figures of the kernel that runs `regime` (cold / window / converged) from the committed `ncu --set full`
    capture, profiles/ncu_traffic.json: DRAM bytes, FP64 pipe %, issue-active %.  A capture taken on one launch that
    covers part of a sweep (the cluster step engine: a launch is a span of the scan) carries `dram_bytes_per_datum`;
    the sweep's traffic is that times N.  None when no capture exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            prof = json.load(fh).get(workload, {}).get(regime)
    except Exception:
        return None
    if prof and "dram_bytes" not in prof and "dram_bytes_per_datum" in prof:
        prof = dict(prof, dram_bytes=prof["dram_bytes_per_datum"] * N)
    return prof


def kernel_of(cov, D, regime):
    """The kernel a timed sweep of this regime spends its time in."""
    DP = next(d for d in (1, 2, 4, 8, 16, 32, 64) if d >= D)
    if cov == "full" and DP <= 16:
        if regime == "cold":
            return ("k_clu_sweep<%d>: thread-block cluster of 16 CTAs, one warp per component, weights and draws over DSMEM; "
                    "a sweep is a few dozen launches (spans of the scan)" % DP)
        return "k_fast_sweep<%d> (148 replicated CTAs, speculative windows; one launch = one sweep)" % DP
    if cov == "full" and DP in (32, 64):
        if regime == "cold":
            return "k_big_sweep<%d> (thread-block cluster, records over DSMEM; a sweep is a few launches)" % DP
        return "k_big_window<%d> (data-parallel stay test) + k_big_sweep<%d> around the movers" % (DP, DP)
    return "k_sweep (generic engine)"


def regime_of(moves, N):
    """cold: >= 10 % of the data move in the sweep (sequential regime); window: some move; converged: (almost) none."""
    if moves >= 0.1 * N:
        return "cold"
    return "window" if moves > 1e-5 * N else "converged"


# ------------------------------------------------------------------------------------------------------------
# CPU arms: the reference's own implementation (oracle/_ref, mechanically shimmed Python) or the C oracle port
# ------------------------------------------------------------------------------------------------------------
def _cpu_chain(kind, sampler, X, z0, cov, K_max, power, seed):
    """Returns (step, reset): step() runs the next sweep of the chain on the CPU and returns its evals; reset() puts
    the chain back to the initial state (sweep index 0).  Same per-sweep inputs protocol as the GPU arm: sweep 0 is
    plain CRP (pcrpmm.py:105: the power applies once i_iter > power_burnin = 0), the scan is random from sweep 0."""
    import random
    m_0, k_0, v_0, S_0 = prior_for(X.shape[1], cov)
    state = {"i": 0, "model": None}
    N = X.shape[0]
    if kind == "reference":
        from oracle.make_ref import import_ref
        NIW, CRPMM, PCRPMM, _, _ = import_ref()
        cls = CRPMM if sampler == "CRPMM" else PCRPMM

        def reset():
            random.seed(seed)
            np.random.seed(seed)
            model = cls(X, NIW(m_0, k_0, v_0, S_0), 1.0, None, assignments=z0.tolist(), K_max=K_max, covariance_type=cov)
            # the reference's update_record_dict computes O(K_true K N) Python metrics; the sweep alone is timed
            model.update_record_dict = lambda rec, i, z, t: rec
            state["model"], state["i"] = model, 0

        def step():
            model = state["model"]
            K_before = model.components.K
            if sampler == "CRPMM":
                model.collapsed_gibbs_sampler(1, None, num_saved=0)
            else:  # one sweep through the reference's own public method, continuing the chain
                model.collapsed_gibbs_sampler(1, None, n_power=power, power_burnin=(-1 if state["i"] > 0 else 0),
                                              num_saved=0)
            state["i"] += 1
            return N * 0.5 * (K_before + model.components.K)
        reset()
        return step, reset
    from oracle import oracle as O
    tab = O.logcount_table(N, power) if power > 1 else None

    def reset():
        orc = O.Oracle(X, m_0, k_0, v_0, S_0, K_max=K_max, covariance_type=cov)
        orc.set_assignments(z0)
        state["model"], state["i"] = orc, 0

    def step():
        order, u = step_input(N, state["i"], power, seed)
        st = state["model"].sweep(u, 1.0, order=order, logcount_tab=tab if state["i"] > 0 else None)
        state["i"] += 1
        return st.evals
    reset()
    return step, reset


def cpu_kind():
    try:
        from oracle.make_ref import import_ref
        import_ref()
        return "reference"
    except Exception:
        return "port"


N_CPU = 20000   # the CPU arm runs the first N_cpu = 2e4 rows (the reference at full N is ~30 min/sweep)


def cpu_rows(kind, wl, rows=0):
    if rows:
        return rows
    D = WORKLOADS[wl][2]
    if kind == "reference":
        return {64: 4000}.get(D, N_CPU)     # D = 64: ~1e3 data/s per core and an N x D x D outer-product cache
    return {64: 20000}.get(D, 10 * N_CPU)   # the C port is ~10x the reference per core


def _cpu_worker(args):
    kind, wl, n_rows, seed, conn = args
    sampler, N, D, K_true, power, cov = WORKLOADS[wl]
    X, _, z0 = gen_data(n_rows, D, K_true, 1, chain_seed=seed)
    step, reset = _cpu_chain(kind, sampler, X, z0, cov, 4 * K_true + 64, power, seed)
    conn.send("ready")
    while True:
        msg = conn.recv()
        if msg == "stop":
            break
        if msg == "reset":
            reset()
            conn.send("ok")
            continue
        t = time.perf_counter()
        ev = step()
        conn.send((ev, time.perf_counter() - t))


def run_reference_arm(a):
    """bench.py --impl reference: the reference's CPU path on all host cores (independent seeded chains, one per
    process -- the reference itself is single threaded), same workload / metric / unit / protocol (W warm-up sweeps
    from the rand initial state, back to the initial state, sweep 0 untimed, K timed sweeps), bounded row sample."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = cpu_kind()
    wl = a.workload
    sampler, N, D, K_true, power, cov = WORKLOADS[wl]
    cores = max(1, min(os.cpu_count() or 1, 64))
    n_rows = cpu_rows(kind, wl, a.rows)
    ctx = mp.get_context("spawn")
    pipes, procs = [], []
    for w in range(cores):
        pa, pb = ctx.Pipe()
        p = ctx.Process(target=_cpu_worker, args=((kind, wl, n_rows, 1 + w, pb),), daemon=True)
        p.start()
        pipes.append(pa); procs.append(p)
    for pa in pipes:
        assert pa.recv() == "ready"

    def one_step():
        for pa in pipes:
            pa.send("go")
        res = [pa.recv() for pa in pipes]
        return sum(r[0] for r in res), max(r[1] for r in res)
    for _ in range(a.warmup):
        one_step()
    for pa in pipes:
        pa.send("reset")
    for pa in pipes:
        assert pa.recv() == "ok"
    one_step()      # sweep 0 of the chain, untimed
    evals, t0 = 0.0, time.perf_counter()
    for _ in range(a.steps):
        ev, _ = one_step()
        evals += ev
    wall = time.perf_counter() - t0
    for pa in pipes:
        pa.send("stop")
    value = evals / wall
    sample = ("%d independent seeded chains (one per core) x first-%d-row sample of %s, %s sampler, K_true=%d; timed: "
              "sweeps 1..%d from the rand initial state" % (cores, n_rows, wl, sampler, K_true, a.steps))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * wall / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s D=%d %s K_true=%d r=%s (CPU: rows=%d per chain)" % (
            wl, sampler, D, cov, K_true, power, n_rows),
            "protocol": "W warm-up sweeps from the rand initial state, reset, sweep 0 untimed, sweeps 1..K timed"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sweeps_per_s_per_chain": a.steps / wall,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(wl, budget_s=25.0):
    """1-core CPU baseline timed inside the default run (rank 0, N=1): the reference if importable, else the port.
    Same protocol on a bounded sample: sweep 0 untimed, then sweeps 1.. until the budget is spent (at most 5)."""
    kind = cpu_kind()
    sampler, N, D, K_true, power, cov = WORKLOADS[wl]
    n_rows = cpu_rows(kind, wl)
    X, _, z0 = gen_data(n_rows, D, K_true, 1)
    step, _ = _cpu_chain(kind, sampler, X, z0, cov, 4 * K_true + 64, power, 1)
    step()  # sweep 0, untimed
    evals, t0, n = 0.0, time.perf_counter(), 0
    while n < 1 or (time.perf_counter() - t0 < budget_s * 0.6 and n < 5):
        evals += step()
        n += 1
    wall = time.perf_counter() - t0
    return {"value": evals / wall, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "first %d rows of %s, sweeps 1..%d timed after sweep 0 from the rand initial state, single "
                      "process (the reference is single threaded)" % (n_rows, wl, n)}


# ------------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    from pybgmm_b200 import _lib, fanout

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    saved_stdout = None
    if world > 1:
        # stdout carries the ONE JSON line: anything a library prints there at communicator set-up goes to stderr
        # instead -- file descriptor 1 is pointed at stderr until the line is printed (NCCL_DEBUG is left alone)
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available() or _lib.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    wl = a.workload
    sampler, N, D, K_true, power, cov = WORKLOADS[wl]
    if a.rows:
        N = a.rows
    K_max = 4 * K_true + 64
    W, K = a.warmup, a.steps

    # Multi-GPU: a single exact chain does not shard; the two fan-outs BASELINE.json names are independent
    # chains on the same data (configs[3]: seeds 1..G) and disjoint contiguous shards of one data set (configs[4]).
    if a.replicas or world == 1:
        shard, chain_seed, fan = 0, 1, ("replicas: every GPU runs the same seeded chain" if world > 1 else "one chain")
    elif wl == "c5":
        shard, chain_seed, fan = rank, 1 + rank, "disjoint shards: rank g owns rows [g N, (g+1) N) of one %d-row data set" % (world * N)
    else:
        shard, chain_seed, fan = 0, 1 + rank, "independent chains: rank g runs chain seed 1+g on the same data"
    X, z_true, z0 = gen_data(N, D, K_true, 1, shard=shard, chain_seed=chain_seed)
    m_0, k_0, v_0, S_0 = prior_for(D, cov)
    n_sw = max(W, K + 1)
    orders, unis = step_inputs(N, n_sw, power, chain_seed)
    chain = _lib.Chain(X, m_0, k_0, v_0, S_0, K_max, covariance_type=cov, device=local_rank)
    stream = torch.cuda.current_stream(dev)
    chain.set_stream(stream.cuda_stream)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def reduce_max(x):
        return reduce(x, dist.ReduceOp.MAX if world > 1 else None)

    def reduce_sum(x):
        return reduce(x, dist.ReduceOp.SUM if world > 1 else None)

    def pw(s):  # pcrpmm.py:105: the power applies once i_iter > power_burnin (= 0)
        return power if (power > 1 and s > 0) else 1.0

    # ---------------- device-resident arm: inputs already in HBM -------------------------------------------------
    d_o = [torch.from_numpy(o).to(dev) if o is not None else None for o in orders]
    d_u = [torch.from_numpy(u).to(dev) for u in unis]

    def dev_sweep(s):
        return chain.sweep_dev(1.0, pw(s), 0 if d_o[s] is None else d_o[s].data_ptr(), d_u[s].data_ptr())
    sampler_thread = ClockSampler(local_rank)
    if rank == 0:
        sampler_thread.start()
    chain.set_assignments(z0)
    warm = [dev_sweep(s) for s in range(W)]          # W warm-up steps: the chain's own cold sweeps
    chain.set_assignments(z0)                        # back to the initial state
    sweep0 = dev_sweep(0)                            # one untimed sweep, then time sweeps 1..K
    sync_all()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_region0 = time.time()
    e0.record(stream)
    stats = [dev_sweep(s) for s in range(1, K + 1)]
    e1.record(stream)
    sync_all()
    t_region1 = time.time()
    clocks = sampler_thread.stop(t_region0, t_region1) if rank == 0 else None
    ms_rank = e0.elapsed_time(e1)
    ms = reduce_max(ms_rank)
    evals = float(sum(st.evals for st in stats))
    total_evals = reduce_sum(evals)
    rate_sum = reduce_sum(evals / (ms_rank * 1e-3))
    per_rank = None
    if world > 1:   # every rank's own time and movers in the timed region (what the max is taken over)
        t = torch.zeros(world, 3, dtype=torch.float64, device=dev)
        t[rank, 0] = ms_rank
        t[rank, 1] = float(sum(st.moves for st in stats))
        t[rank, 2] = evals
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        per_rank = {"ms": [round(v, 3) for v in t[:, 0].tolist()], "moves": [int(v) for v in t[:, 1].tolist()],
                    "evals_per_s": [e / (m * 1e-3) for m, e in zip(t[:, 0].tolist(), t[:, 2].tolist())]}
    kernel_ms = sum(st.sweep_kernel_ms for st in stats)    # CUDA events around the sweep kernel alone, on its stream
    launches = int(sum(st.launches for st in stats))
    z_value_arm = chain.assignments()

    # ---------------- e2e arm: the same sweeps from the same state, host buffers through the C-ABI call ---------
    pin_o = [torch.from_numpy(o).pin_memory() if o is not None else None for o in orders]
    pin_u = [torch.from_numpy(u).pin_memory() for u in unis]
    z_host = torch.empty(N, dtype=torch.int64).pin_memory()
    z_host_np = z_host.numpy()

    def host_sweep(s):
        return chain.sweep(1.0, pw(s), None if pin_o[s] is None else pin_o[s].numpy(), pin_u[s].numpy())
    chain.set_assignments(z0)
    host_sweep(0)                                    # sweep 0 untimed (also the host path's warm-up)
    sync_all()
    t_wall = time.perf_counter()
    e0.record(stream)
    e2e_evals = 0
    for s in range(1, K + 1):
        st = host_sweep(s)
        e2e_evals += st.evals
        # device -> host read of the step's result: the relabelled assignments
        _lib._check(_lib.lib().bgmm_get_state(chain._h, _lib._ip(z_host_np), None, None, None, None, None, None))
    e1.record(stream)
    sync_all()
    e2e_ms = max(reduce_max(e0.elapsed_time(e1)), reduce_max(1e3 * (time.perf_counter() - t_wall)))
    e2e_total_evals = reduce_sum(float(e2e_evals))
    same = bool((z_value_arm == z_host_np).all())          # both arms walked the same chain
    h2d = (8 * N if power > 1 else 0) + 8 * N
    d2h = 8 * N

    # gather of per-rank assignments at the end (NCCL over NVLink), outside the timed steps; the first call carries
    # NCCL's lazy set-up of the all_gather channels, the second is the steady-state cost
    gather_ms = None
    if world > 1:
        zt = fanout.chain_assignments_tensor(chain, dev)
        gather_ms = []
        for _ in range(2):
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sync_all()
            g0.record()
            outs, ks = fanout.gather_assignments(zt, chain.K)
            g1.record()
            sync_all()
            gather_ms.append(reduce_max(g0.elapsed_time(g1)))

    # ---------------- optional: many independent chains on this GPU (bgmm_sweep_many) ---------------------------
    multi = None
    if a.chains > 1 and world == 1:
        multi = run_multi_chain(a, _lib, torch, dev, X, z0, (m_0, k_0, v_0, S_0), K_max, cov, power, N, K)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = total_evals / (ms * 1e-3)
    peak, peak_src = measured_peak()
    b_eval = 8 * (D * D + D + 2) if cov == "full" else 8 * (2 * D + 2)
    b_datum = 8 * D + 32
    alg_bytes = evals * b_eval + K * N * b_datum                            # this rank, K launches
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    flops_eval = (D * D + 3 * D + 30) if cov == "full" else (8 * D + 30)    # ~fp64 FMAs*2 per eval, DESIGN.md
    # regimes of the timed sweeps (by share of movers) and the one that dominates the timed region's time
    regimes = {}
    for st in stats:
        r = regimes.setdefault(regime_of(st.moves, N), {"sweeps": 0, "ms": 0.0, "evals": 0.0, "moves": 0})
        r["sweeps"] += 1; r["ms"] += st.sweep_kernel_ms; r["evals"] += st.evals; r["moves"] += int(st.moves)
    for r in regimes.values():
        r["evals_per_s"] = r["evals"] / (r["ms"] * 1e-3) if r["ms"] > 0 else None
        r["us_per_mover"] = 1e3 * r["ms"] / r["moves"] if r["moves"] else None
    dominant = max(regimes, key=lambda k: regimes[k]["ms"])
    prof = ncu_profile(wl, dominant, N) or {}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "%s: %s D=%d %s N=%d per GPU, K_true=%d, r=%s, rand init K=%d, K_max=%d" % (
            wl, sampler, D, cov, N, K_true, power, K_true, K_max),
            "protocol": "W warm-up sweeps from the rand initial state, chain put back to that state, sweep 0 untimed, "
                        "sweeps 1..K timed (SURVEY.md 8d): the cold sweeps are inside the timed region",
            "chains": fan,
            "l2": "inputs_larger_than_l2 (X %.0f MB + per-step order/uniform buffers %.0f MB, never reused; no "
                  "explicit flush)" % (8e-6 * N * D, 16e-6 * N),
            "K_live_mean": evals / (K * N), "moves_per_sweep": [int(st.moves) for st in stats],
            "ms_per_sweep": [round(st.device_ms, 3) for st in stats], "K_live": [int(st.K) for st in stats]},
        "sweeps_per_s": world * K / (ms * 1e-3),
        "sum_of_rank_rates": rate_sum,
        "e2e": {"value": e2e_total_evals / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / K, "same_chain_as_value_arm": same},
        "gpu_launches": launches,
        "clocks": clocks,
        "regimes": regimes,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": prof.get("dram_bytes"), "traffic_regime": dominant if prof else None,
                     "peak_source": peak_src,
                     "kernel": kernel_of(cov, D, dominant),
                     "algorithmic_bytes_per_eval": b_eval, "algorithmic_bytes_per_datum": b_datum,
                     "kernel_ms_per_launch": kernel_ms / K, "launches_per_step": launches / K,
                     "true_bound": "latency of the sequential dependency: every datum is one serial step of the cluster "
                                   "(cold regime: two DSMEM hops + one scan of the K + 1 weights) or every mover one "
                                   "window round (window regime); HBM and the FP64 pipe are both far from saturated",
                     "fp64_tflops": evals * 2 * flops_eval / (kernel_ms * 1e-3) / 1e12,
                     "ncu": prof or None,
                     "note": "secondary figure: ALGORITHMIC bytes per SURVEY.md 8(d) (one sufficient-statistic record "
                             "per eval) over the timed kernel time; the records live in shared memory / registers "
                             "and are reused across data, so DRAM traffic (`traffic`, ncu capture of the regime that "
                             "dominates the timed region) is far below it"},
        "warmup_chain": {"note": "the %d untimed warm-up sweeps from the rand initial state" % W,
                         "ms": [round(st.device_ms, 3) for st in warm], "moves": [int(st.moves) for st in warm]},
        "sweep0": {"ms": round(sweep0.device_ms, 3), "moves": int(sweep0.moves), "K_live": int(sweep0.K)},
        "engine": {"windows": [int(st.windows) for st in stats], "seq_data": [int(st.seq_data) for st in stats],
                   "fast_steps": [int(st.fast_steps) for st in stats],
                   "wasted": [int(st.wasted) for st in stats], "min_margin": min(st.min_margin for st in stats),
                   "guard_hits": int(sum(st.guard_hits for st in stats) + sweep0.guard_hits)},
    }
    if achieved / peak > 1.0:
        # the records are reused from shared memory (k_big_window stages B once per 32 data): algorithmic bytes over time
        # exceed the HBM peak and say nothing about HBM.  The kernel is an FP64 product: report the compute fraction,
        # explicitly labelled (vector FP64 FMAs, not tensor cores; no measured FP64 peak in MEASURED_PEAKS.json)
        rf = line["roofline"]
        rf["hbm_algorithmic"] = {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                 "note": "above 1: not evidence of HBM utilisation"}
        rf.update({"bound": "tensor", "achieved": rf["fp64_tflops"], "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                   "frac": rf["fp64_tflops"] / FP64_PEAK_TFLOPS, "pipe": "fp64 vector FMA (no tensor cores on this path)",
                   "peak_source": "nominal B200 FP64 peak %.0f TFLOP/s (MEASURED_PEAKS.json holds no FP64 figure)" % FP64_PEAK_TFLOPS})
    if gather_ms is not None:
        line["gather_assignments_ms"] = {"first_call": gather_ms[0], "steady": gather_ms[1]}
        line["per_rank"] = per_rank
    if multi is not None:
        line["multi_chain"] = multi
    if world == 1 and not a.no_cpu:
        line["cpu_baseline"] = cpu_baseline_leg(wl)
    if world > 1:
        dist.destroy_process_group()      # while stdout still points at stderr: NCCL logs its teardown at NCCL_DEBUG=INFO
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if saved_stdout is not None:
        os.dup2(2, 1)      # whatever a library still prints while the process exits (NCCL at NCCL_DEBUG=INFO) is not stdout's


def run_multi_chain(a, _lib, torch, dev, X, z0_first, prior, K_max, cov, power, N, K):
    """`--chains M`: M independent chains (seeds 1..M) on the same data resident on ONE GPU, one CTA each, all advanced
    by one kernel launch per sweep (bgmm_sweep_many).  Same protocol: sweep 0 untimed, sweeps 1..K timed."""
    M = a.chains
    m_0, k_0, v_0, S_0 = prior
    D = X.shape[1]
    K_true = WORKLOADS[a.workload][3]
    first = _lib.Chain(X, m_0, k_0, v_0, S_0, K_max, covariance_type=cov, device=dev.index)
    chains = [first] + [first.fork() for _ in range(M - 1)]
    stream = torch.cuda.current_stream(dev)
    for c in chains:
        c.set_stream(stream.cuda_stream)
    for m, c in enumerate(chains):
        c.set_assignments(gen_data(N, D, K_true, 1, chain_seed=1 + m)[2] if m else z0_first)
    gen = torch.Generator(device=dev)
    gen.manual_seed(4242)

    def inputs():   # per-chain scan orders and uniforms, generated on the device (this line is not a parity run)
        o = torch.rand(M, N, device=dev, generator=gen).argsort(dim=1) if power > 1 else None
        u = torch.rand(M, N, device=dev, dtype=torch.float64, generator=gen)
        return o, u
    group = _lib.ChainGroup(chains)
    o, u = inputs()
    group.sweep_dev(1.0, 1.0, o, u)                  # sweep 0 untimed
    torch.cuda.synchronize(dev)
    ins = [inputs() for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    evals, moves = 0, []
    for s in range(K):
        sts = group.sweep_dev(1.0, power if power > 1 else 1.0, ins[s][0], ins[s][1])
        evals += sum(st.evals for st in sts)
        moves.append(int(sum(st.moves for st in sts)))
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    return {"chains": M, "value": evals / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K,
            "moves_per_sweep_all_chains": moves,
            "note": "%d independent exact chains on one GPU (one CTA per chain, register-resident sequential engine), "
                    "one launch per sweep; aggregate evals/s over sweeps 1..%d from rand initial states" % (M, K)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override N per GPU / rows per CPU chain (debug)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--replicas", action="store_true", help="N>1: every rank runs the same seeded chain")
    ap.add_argument("--chains", type=int, default=0, help="also time this many independent chains on one GPU")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
