#!/usr/bin/env python
"""bench.py -- Gibbs sweeps/s and datum-component log_post_pred evals/s of the B200 engine (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4|c5] [--impl ours|reference]

A *step* is one Gibbs sweep (one pass of crpmm.py:57-88 / pcrpmm.py:93-131 over all N data).  Default workload
= BASELINE.json's target configuration C3: PCRPMM, NIW full covariance, N=1e6, D=16, K_true=100, r=1.5, random scan
(SURVEY.md 8d), synthetic data from the demos' generator, `rand` initial assignments with K=K_true.  The chain
starts from that initial state: W warm-up sweeps, then K timed sweeps; no hidden burn-in.

Prints ONE JSON line (rank 0).  `value` = whole-job evals/s with the per-step inputs (scan order + uniforms)
already resident in HBM; `e2e` = the same K sweeps replayed from the same initial state through the host C-ABI
call (bgmm_sweep) with pinned HOST buffers -> H2D of order+uniforms and D2H of the assignments inside the timed
region.  Under torchrun (N>1) every rank runs an independent chain on its own shard (weak scaling, no data-path
collective) and the assignments are all-gathered with NCCL after the timed region.
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


def gen_data(N, D, K_true, seed):
    """examples/crpmm_2d_demo.py:41-55 scaled."""
    rs = np.random.RandomState(seed)
    z_true = rs.randint(0, K_true, N)
    mu = rs.randn(D, K_true) * 4.0
    X = np.ascontiguousarray((mu[:, z_true] + rs.randn(D, N) * 0.7).T)
    z0 = rs.randint(0, K_true, N).astype(np.int64)
    for k in range(z0.max()):  # consecutive labels (igmm.py:89-94)
        while not (z0 == k).any():
            z0[z0 > k] -= 1
        if z0.max() == k:
            break
    return X, z_true, z0


def prior_for(D, cov):
    v_0 = D + 3
    return np.zeros(D), 0.7 ** 2 / 4.0 ** 2, v_0, 0.7 ** 2 * v_0 * (np.eye(D) if cov == "full" else np.ones(D))


def step_inputs(N, n_steps, power, seed):
    """Per-step scan order (pcrpmm.py:89) and uniforms (utils.py:15), generated on the host like the reference."""
    rs = np.random.RandomState(seed + 7919)
    orders = [rs.permutation(N).astype(np.int64) if power > 1 else None for _ in range(n_steps)]
    unis = [rs.random_sample(N) for _ in range(n_steps)]
    return orders, unis


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


def ncu_traffic(workload):
    """DRAM bytes per launch of the sweep kernel from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            return json.load(fh).get(workload)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------------
# CPU arms: the reference's own implementation (oracle/_ref, mechanically shimmed Python) or the C oracle port
# ------------------------------------------------------------------------------------------------------------
def _cpu_chain(kind, sampler, X, z0, cov, K_max, power, seed):
    """Returns step() -> evals of one sweep on the CPU."""
    import random
    m_0, k_0, v_0, S_0 = prior_for(X.shape[1], cov)
    random.seed(seed)
    np.random.seed(seed)
    if kind == "reference":
        from oracle.make_ref import import_ref
        NIW, CRPMM, PCRPMM, _, _ = import_ref()
        cls = CRPMM if sampler == "CRPMM" else PCRPMM
        model = cls(X, NIW(m_0, k_0, v_0, S_0), 1.0, None, assignments=z0.tolist(), K_max=K_max, covariance_type=cov)
        state = {"i": 0}

        def step():
            # one sweep through the reference's own public method, continuing the chain: the power schedule is
            # `i_iter > power_burnin` (pcrpmm.py:105), so power_burnin=-1 keeps the power on for a 1-sweep call
            K_before = model.components.K
            if sampler == "CRPMM":
                model.collapsed_gibbs_sampler(1, None, num_saved=0)
            else:
                model.collapsed_gibbs_sampler(1, None, n_power=power, power_burnin=(-1 if state["i"] > 0 else 0),
                                              num_saved=0)
            state["i"] += 1
            return X.shape[0] * 0.5 * (K_before + model.components.K)
        # the reference's update_record_dict needs labels for its metrics; give it a no-op to time the sweep only
        model.update_record_dict = lambda rec, i, z, t: rec
        return step
    from oracle import oracle as O
    orc = O.Oracle(X, m_0, k_0, v_0, S_0, K_max=K_max, covariance_type=cov)
    orc.set_assignments(z0)
    tab = O.logcount_table(X.shape[0], power) if power > 1 else None
    state = {"i": 0}
    rs = np.random.RandomState(seed)

    def step():
        order = rs.permutation(X.shape[0]) if power > 1 else None
        st = orc.sweep(rs.random_sample(X.shape[0]), 1.0, order=order, logcount_tab=tab if state["i"] > 0 else None)
        state["i"] += 1
        return st.evals
    return step


def cpu_kind():
    try:
        from oracle.make_ref import import_ref
        import_ref()
        return "reference"
    except Exception:
        return "port"


def _cpu_worker(args):
    kind, wl, n_rows, seed, n_warm, n_steps, conn = args
    sampler, N, D, K_true, power, cov = WORKLOADS[wl]
    X, _, z0 = gen_data(n_rows, D, K_true, seed)
    step = _cpu_chain(kind, sampler, X, z0, cov, 4 * K_true + 64, power, seed)
    conn.send("ready")
    while True:
        msg = conn.recv()
        if msg == "stop":
            break
        t = time.perf_counter()
        ev = step()
        conn.send((ev, time.perf_counter() - t))


def cpu_rows(kind, wl):
    # bounded sample: the reference is ~5e2..3e3 data/s per core, the C port ~2e4 data/s
    D = WORKLOADS[wl][2]
    if kind == "reference":
        return {2: 6000, 8: 4000, 16: 2500, 64: 600}.get(D, 2500)
    return {2: 200000, 8: 60000, 16: 30000, 64: 3000}.get(D, 30000)


def run_reference_arm(a):
    """bench.py --impl reference: the reference's CPU path on all host cores (independent seeded chains, one per
    process -- the reference itself is single threaded), same workload / metric / unit, bounded row sample."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = cpu_kind()
    wl = a.workload
    sampler, N, D, K_true, power, cov = WORKLOADS[wl]
    cores = max(1, min(os.cpu_count() or 1, 64))
    n_rows = cpu_rows(kind, wl)
    ctx = mp.get_context("spawn")
    pipes, procs = [], []
    for w in range(cores):
        pa, pb = ctx.Pipe()
        p = ctx.Process(target=_cpu_worker, args=((kind, wl, n_rows, 1 + w, a.warmup, a.steps, pb),), daemon=True)
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
    evals, t0 = 0.0, time.perf_counter()
    for _ in range(a.steps):
        ev, _ = one_step()
        evals += ev
    wall = time.perf_counter() - t0
    for pa in pipes:
        pa.send("stop")
    value = evals / wall
    sample = "%d independent seeded chains (one per core) x first-%d-row sample of %s, %s sampler, K_true=%d" % (
        cores, n_rows, wl, sampler, K_true)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * wall / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s D=%d %s K_true=%d r=%s (CPU: rows=%d per chain)" % (
            wl, sampler, D, cov, K_true, power, n_rows)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sweeps_per_s_per_chain": a.steps / wall,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(wl, budget_s=25.0):
    """1-core CPU baseline timed inside the default run (rank 0, N=1): the reference if importable, else the port."""
    kind = cpu_kind()
    sampler, N, D, K_true, power, cov = WORKLOADS[wl]
    n_rows = cpu_rows(kind, wl)
    X, _, z0 = gen_data(n_rows, D, K_true, 1)
    step = _cpu_chain(kind, sampler, X, z0, cov, 4 * K_true + 64, power, 1)
    step()  # warm-up sweep
    evals, t0, n = 0.0, time.perf_counter(), 0
    while n < 1 or (time.perf_counter() - t0 < budget_s * 0.5 and n < 5):
        evals += step()
        n += 1
    wall = time.perf_counter() - t0
    return {"value": evals / wall, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "first %d rows of %s, %d timed sweep(s) after 1 warm-up, single process (the reference is "
                      "single threaded)" % (n_rows, wl, n)}


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
        # stdout carries the ONE JSON line: anything the collectives library prints there at communicator set-up (NCCL's
        # version banner) goes to stderr instead -- file descriptor 1 is pointed at stderr until the line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when NCCL_DEBUG is VERSION/INFO; stdout carries the JSON line only
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "WARN"
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

    # Multi-GPU: a single exact chain does not shard: every rank runs a replica of the workload -- the same
    # seeded shard and chain, so the per-GPU work is identical at every N (weak scaling of independent processes, no
    # data-path collective).  --distinct gives every rank its own data and chain instead; the time of a Gibbs sweep
    # follows the number of data that move, which differs from chain to chain by 3x in the timed sweeps, and the
    # max-over-ranks time then measures the unluckiest chain rather than the hardware.
    seed_r = 1 + (rank if a.distinct else 0)
    X, z_true, z0 = gen_data(N, D, K_true, seed_r)
    m_0, k_0, v_0, S_0 = prior_for(D, cov)
    orders, unis = step_inputs(N, W + K, power, seed_r)
    chain = _lib.Chain(X, m_0, k_0, v_0, S_0, K_max, covariance_type=cov, device=local_rank)
    stream = torch.cuda.current_stream(dev)
    chain.set_stream(stream.cuda_stream)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def pw(s):  # pcrpmm.py:105: the power applies once i_iter > power_burnin (= 0)
        return power if (power > 1 and s > 0) else 1.0

    # ---------------- e2e arm: host buffers through the C-ABI call -------------------------------------------
    pin_o = [torch.from_numpy(o).pin_memory() if o is not None else None for o in orders]
    pin_u = [torch.from_numpy(u).pin_memory() for u in unis]
    z_host = torch.empty(N, dtype=torch.int64).pin_memory()
    z_host_np = z_host.numpy()
    chain.set_assignments(z0)
    for s in range(W):
        chain.sweep(1.0, pw(s), None if pin_o[s] is None else pin_o[s].numpy(), pin_u[s].numpy())
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    e0.record(stream)
    e2e_evals = 0
    for s in range(W, W + K):
        st = chain.sweep(1.0, pw(s), None if pin_o[s] is None else pin_o[s].numpy(), pin_u[s].numpy())
        e2e_evals += st.evals
        # device -> host read of the step's result: the relabelled assignments
        _lib._check(_lib.lib().bgmm_get_state(chain._h, _lib._ip(z_host_np), None, None, None, None, None, None))
    e1.record(stream)
    sync_all()
    e2e_ms = reduce_max(max(e0.elapsed_time(e1), 1e3 * 0))
    e2e_wall_ms = reduce_max(1e3 * (time.perf_counter() - t_wall))
    e2e_ms = max(e2e_ms, e2e_wall_ms)  # host work between launches counts end to end
    e2e_total_evals = reduce_sum(float(e2e_evals))
    z_e2e = z_host_np.copy()
    h2d = (8 * N if power > 1 else 0) + 8 * N
    d2h = 8 * N

    # ---------------- device-resident arm: same initial state, same inputs, already in HBM -------------------
    d_o = [torch.from_numpy(o).to(dev) if o is not None else None for o in orders]
    d_u = [torch.from_numpy(u).to(dev) for u in unis]
    chain.set_assignments(z0)
    sampler_thread = ClockSampler(local_rank)
    if rank == 0:
        sampler_thread.start()
    warm = []
    for s in range(W):
        warm.append(chain.sweep_dev(1.0, pw(s), 0 if d_o[s] is None else d_o[s].data_ptr(), d_u[s].data_ptr()))
    sync_all()
    sync_all()
    t_region0 = time.time()
    e0.record(stream)
    stats = []
    for s in range(W, W + K):
        stats.append(chain.sweep_dev(1.0, pw(s), 0 if d_o[s] is None else d_o[s].data_ptr(), d_u[s].data_ptr()))
    e1.record(stream)
    sync_all()
    t_region1 = time.time()
    clocks = sampler_thread.stop(t_region0, t_region1) if rank == 0 else None
    ms_rank = e0.elapsed_time(e1)
    ms = reduce_max(ms_rank)
    evals = float(sum(st.evals for st in stats))
    per_rank = None
    if world > 1:   # every rank's own time and movers in the timed region (what the max is taken over)
        t = torch.zeros(world, 2, dtype=torch.float64, device=dev)
        t[rank, 0] = ms_rank
        t[rank, 1] = float(sum(st.moves for st in stats))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        per_rank = {"ms": [round(v, 3) for v in t[:, 0].tolist()], "moves": [int(v) for v in t[:, 1].tolist()]}
    total_evals = reduce_sum(evals)
    kernel_ms = sum(st.sweep_kernel_ms for st in stats)    # CUDA events around the sweep kernel alone, on its stream
    launches = int(sum(st.launches for st in stats))
    same = bool((chain.assignments() == z_e2e).all())      # both arms walked the same chain

    # gather of per-rank assignments at the end (NCCL over NVLink), outside the timed steps
    gather_ms = None
    if world > 1:
        zt = fanout.chain_assignments_tensor(chain, dev)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        g0.record()
        outs, ks = fanout.gather_assignments(zt, chain.K)
        g1.record()
        sync_all()
        gather_ms = reduce_max(g0.elapsed_time(g1))

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
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "%s: %s D=%d %s N=%d per GPU, K_true=%d, r=%s, rand init K=%d, K_max=%d" % (
            wl, sampler, D, cov, N, K_true, power, K_true, K_max),
            "chains": ("one independent chain on its own shard per GPU (--distinct)" if a.distinct else
                       "replicas: every GPU runs the same seeded shard and chain (a single exact chain does not shard)"),
            "l2": "inputs_larger_than_l2 (X %.0f MB + per-step "
            "order/uniform buffers %.0f MB, never reused; no explicit flush)" % (8e-6 * N * D, 16e-6 * N),
            "K_live_mean": evals / (K * N), "moves_per_sweep": [int(st.moves) for st in stats],
            "K_live": [int(st.K) for st in stats]},
        "sweeps_per_s": world * K / (ms * 1e-3),
        "e2e": {"value": e2e_total_evals / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / K, "same_chain_as_value_arm": same},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(wl), "peak_source": peak_src, "kernel": "k_fast_sweep<16> (one launch = one sweep)" if (cov == "full" and D <= 16) else "k_sweep",
                     "algorithmic_bytes_per_eval": b_eval, "algorithmic_bytes_per_datum": b_datum,
                     "kernel_ms_per_launch": kernel_ms / K,
                     "fp64_tflops": evals * 2 * flops_eval / (kernel_ms * 1e-3) / 1e12,
                     "note": "algorithmic bytes per SURVEY.md 8(d): one sufficient-statistic record per eval; the "
                             "records live in every CTA's shared memory and are reused across data, so DRAM "
                             "traffic is far below this figure and the sweep is bound by the sequential "
                             "dependency (one window round per mover), not by HBM (see DESIGN.md 4)"},
        "cold_chain": {"note": "the %d untimed warm-up sweeps from the rand initial state, device-resident arm" % W,
                       "ms": [round(st.device_ms, 3) for st in warm], "moves": [int(st.moves) for st in warm],
                       "K_live": [int(st.K) for st in warm],
                       "evals_per_s": [st.evals / (st.device_ms * 1e-3) for st in warm]},
        "engine": {"windows": [int(st.windows) for st in stats], "seq_data": [int(st.seq_data) for st in stats],
                   "wasted": [int(st.wasted) for st in stats], "min_margin": min(st.min_margin for st in stats)},
    }
    if gather_ms is not None:
        line["gather_assignments_ms"] = gather_ms
        line["per_rank"] = per_rank
    if world == 1 and not a.no_cpu:
        line["cpu_baseline"] = cpu_baseline_leg(wl)
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override N per GPU (debug)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--distinct", action="store_true", help="N>1: every rank gets its own data and chain (seed 1+rank)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
