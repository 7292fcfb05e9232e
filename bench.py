#!/usr/bin/env python
"""bench.py -- noisy+ideal circuits/sec (exact <O>) of the expectation-value hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic circuits of the named workload:
every circuit is evolved twice -- density matrix under the backend-derived noise table (noisy
values) and statevector (ideal values) -- and all its Pauli observables are evaluated.

Default workload `brick10_guadalupe_twirl` = BASELINE.json configs[1]: 10-qubit random brickwork
circuits (steps 1..5) on a 16-qubit heavy-hex-like chain table, 100 Pauli twirls per base circuit
(20 base circuits -> 2000 circuit instances per rank), 10 single-Z observables each.

JSON keys: see the contract in the task statement; `value` is timed with the lowered programs
resident in HBM (bwq_dm_execute + bwq_sv_execute), `e2e` through the C ABI with HOST buffers
(one bwq_meas_data_run per step = ideal + noisy values of the batch: lowering, H2D of the programs,
kernels, D2H of the values; the density-matrix side is pipelined in segments, the statevector side
runs concurrently on a companion context).
`e2e_estimator` is the same work from the BASE circuit objects through B200Estimator.run(...,
variants=...) (variants generated inside the library); `workloads` carries the other BASELINE configs
(cfg1 tfim4_lima_zne -- runs on dm_onchip_kernel, `value` = that launch alone --, cfg3 tfim14_dm, cfg4
tfim30_sv, amplitude-sharded with exchange figures when --gpus > 1) at reduced steps.
`--impl reference` times the Aer-style CPU restatement (oracle/cpu_ref.cpp; qiskit-aer itself is
not installable here) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import gc


def _park_setup_objects():
    """The synthetic workloads are millions of small Python objects (2000 circuits x 484 gate tuples);
    a generation-2 garbage collection that rescans them costs 50-100 ms and fires at random inside the
    timed host-buffer calls (seen as sporadic 100 ms calls next to 47 ms ones).  gc.freeze() moves
    everything built during setup into the permanent generation: the collector stays on, it just no
    longer rescans the workload."""
    gc.collect()
    gc.freeze()


METRIC = "noisy+ideal circuits/sec (exact <O>)"
UNIT = "circuits/s"


# ----------------------------------------------------------------------------------------------
# workloads (synthetic, seeded)
# ----------------------------------------------------------------------------------------------
def build_workload(name, rank, scale=1.0):
    """-> dict(circuits, observables, backend, desc).  Seeds depend on the rank (weak scaling: every
    rank owns its own batch of the same shape)."""
    from ml_qem_b200 import backends, families as F

    extra = {}  # base circuits + variants descriptor of the workloads that are variants of base circuits
    if name == "brick10_guadalupe_twirl":
        n_base = max(1, int(round(20 * scale)))
        circs, base, obs = F.config_brick10_twirl(n_base=n_base, n_twirls=100, seed=1 + 1000 * rank)
        extra = {"base": base, "variants": {"twirls": 100, "seed": 1 + 1000 * rank}}
        be = backends.synthetic_chain(16, seed=2, name="synthetic_guadalupe_like_16q")
        desc = {"workload": name, "n_qubits": 10, "register": 16, "base_circuits": n_base, "twirls": 100,
                "circuits_per_rank": len(circs), "observables_per_circuit": len(obs), "trotter_steps": "1..5"}
    elif name == "tfim4_lima_zne":
        n_base = max(1, int(round(2000 * scale)))
        circs, base, obs = F.config_tfim4_lima_zne(n_base=n_base, seed=1000 * rank)
        extra = {"base": base, "variants": {"folds": (1, 3, 5)}}
        be = backends.fake_lima()
        desc = {"workload": name, "n_qubits": 4, "register": 5, "base_circuits": n_base, "zne_factors": [1, 3, 5],
                "circuits_per_rank": len(circs), "observables_per_circuit": len(obs)}
    elif name.startswith("tfim") and name.endswith("_dm"):
        n = int(name[4:-3])
        n_c = max(1, int(round(8 * scale)))
        circs, obs = F.config_tfim_dm(n=n, n_circuits=n_c, seed=2 + 1000 * rank)
        be = backends.synthetic_chain(n, seed=n, name=f"synthetic_chain_{n}q")
        desc = {"workload": name, "n_qubits": n, "register": n, "circuits_per_rank": len(circs),
                "observables_per_circuit": len(obs), "trotter_steps": "1..10"}
    elif name == "mixed6_12_dataset":
        n_c = max(3, int(round(5000 * scale)))
        circs, obs_each = F.config_mixed_dataset(n_circuits=n_c, seed=5 + 1000 * rank)
        be = backends.synthetic_chain(12, seed=12, name="synthetic_chain_12q")
        desc = {"workload": name, "n_qubits": 12, "active_qubits": "6..12 uniform", "register": 12, "circuits_per_rank": len(circs),
                "families": "tfim / brickwork / random basis layers, equal parts",
                "observables_per_circuit": "one single-Z per active qubit",
                "note": "generation part of BASELINE configs[4] (50k circuits = 10 steps of this batch)",
                "resident_state_bytes": int(sum(8 * 4 ** len({q for _, qs, _ in c.gate_ops() for q in qs}) for c in circs))}
        return {"circuits": circs, "observables": obs_each, "backend": be, "desc": desc}
    else:
        raise SystemExit(f"unknown workload {name!r}")
    return dict({"circuits": circs, "observables": [obs] * len(circs), "backend": be, "desc": desc}, **extra)


# ----------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi in the background during the timed region)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "100"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "power_w_max": float(max(power)), "samples": len(sm)}
        return out


# ----------------------------------------------------------------------------------------------
# CPU arm: real qiskit-aer when the box has it, else the Aer-style restatement (oracle/cpu_ref.cpp)
# ----------------------------------------------------------------------------------------------
def host_threads():
    """All host cores of the box.  torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm
    (rank 0 only) must not inherit that, so the thread count is passed to the library explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def probe_qiskit_aer():
    """SURVEY 8(c) run-time probe (oracle/aer_probe.py): the reference's own simulator, if this box
    has it (also under a driver-provided baseline/_ref).  Nothing in the image ships it; when it is
    absent the CPU arm is the C++ restatement and says so (kind = "port")."""
    from oracle import aer_probe

    return aer_probe.find()


def aer_reference_rate(wl, n, cores):
    """Real Aer through the reference's own call sequence (blackwater/data/utils.py:422-430) on the
    first n circuits of the workload.  Returns (circuits/s, noisy values, ideal values)."""
    from oracle import aer_probe

    props = wl["backend"].to_dict()
    t = time.perf_counter()
    noisy, ideal = [], []
    for c, obs in zip(wl["circuits"][:n], wl["observables"][:n]):
        noisy.append(aer_probe.estimate(c.num_qubits, c.gate_ops(), obs, props, threads=cores))
        ideal.append(aer_probe.estimate(c.num_qubits, c.gate_ops(), obs, None))
    dt = time.perf_counter() - t
    return n / dt, np.concatenate(noisy), np.concatenate(ideal), dt


def cpu_reference_rate(workload_name, budget_s=15.0, max_circuits=512, wl=None, first=None):
    """Times noisy (density matrix) + ideal (statevector) evaluation of the first circuits of the
    rank-0 workload on all host cores.  Returns (circuits/s, cores, sample description, parity data)."""
    from ml_qem_b200 import engine
    from ml_qem_b200.gateset import OPCODES
    from oracle import cpu_ref, noise_model as onm

    cpu_ref.build()
    if wl is None:
        wl = build_workload(workload_name, 0, scale=0.3)
    if probe_qiskit_aer() is not None:
        try:  # the genuine reference: small bounded sample, the reference's per-circuit calls
            n = int(first or 4)
            rate, v_dm, v_sv, dt = aer_reference_rate(wl, n, host_threads())
            fb = engine.encode_batch(wl["circuits"][:n], wl["observables"][:n])
            return (rate, host_threads(), f"REAL qiskit-aer: first {n} circuits of {workload_name}, AerEstimator density_matrix "
                    f"(approximation=True, shots=None) + qiskit Estimator, {dt:.2f} s", (fb, v_dm, v_sv), "reference")
        except Exception as exc:  # noqa: BLE001
            print(f"bench.py: qiskit-aer present but unusable ({exc!r}); using the C++ restatement", file=sys.stderr)
    onoise = cpu_ref.noise_arrays(onm.from_backend(wl["backend"].to_dict()), OPCODES)
    cores = host_threads()
    # size the sample: time one circuit, then as many as fit the budget
    fb1 = engine.encode_batch(wl["circuits"][:1], wl["observables"][:1])
    cpu_ref.prepare(fb1)  # gate matrices from oracle/gates.py: data preparation, not part of the timed simulation
    t = time.perf_counter()
    cpu_ref.run_dm(fb1, onoise, threads=cores)
    cpu_ref.run_sv(fb1, threads=cores)
    t1 = max(time.perf_counter() - t, 1e-4)
    # small states run circuit-parallel (one circuit per core), large ones amplitude-parallel
    est_rate = (cores if wl["desc"]["n_qubits"] < 7 else 1.0) / t1
    cap = max_circuits if wl["desc"]["n_qubits"] >= 7 else 4096
    n = int(max(1, min(cap, len(wl["circuits"]), budget_s * est_rate)))
    if first is not None:
        n = min(n, first)
    fb = engine.encode_batch(wl["circuits"][:n], wl["observables"][:n])
    cpu_ref.prepare(fb)
    reps, dt = 0, 0.0
    while True:  # small circuits: repeat the sample until it amounts to ~10 s of CPU work
        t = time.perf_counter()
        v_dm, s1 = cpu_ref.run_dm(fb, onoise, threads=cores)
        v_sv, s2 = cpu_ref.run_sv(fb, threads=cores)
        dt += time.perf_counter() - t
        reps += 1
        if dt >= min(10.0, budget_s) or dt * (reps + 1) / reps > budget_s:
            break
    assert not s1.any() and not s2.any()
    return (n * reps / dt, cores, f"first {n} circuits of {workload_name} (rank-0 seed) x {reps} pass(es), noisy DM + ideal SV, "
            f"{dt:.2f} s on {cores} threads", (fb, v_dm, v_sv), "port")


# ----------------------------------------------------------------------------------------------
class Ctx:
    """One rank of the run: engine, process group, helpers for the max/sum over ranks."""

    def __init__(self, args):
        import torch

        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.torch = torch
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist_mod

            self.dist = dist_mod
            self.dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        from ml_qem_b200 import engine

        self.eng = engine.Engine(self.local_rank)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            self.peak, self.peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        else:
            self.peak, self.peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX) if self.dist else x

    def sum(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM) if self.dist else x

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# BASELINE configs[3]: ideal-label statevector of ONE wide TFIM circuit, amplitudes sharded over
# all ranks (the engine's P2P exchange kernel / NCCL all_to_all for the global-qubit swaps).
# Strong scaling: the circuit is fixed.
# ----------------------------------------------------------------------------------------------
def bench_sharded_sv(ctx, name, steps, warmup, first_trotter=1, cpu_check=True, clocks=False):
    import torch

    from ml_qem_b200 import engine, families as F
    from ml_qem_b200.statevector import GpuExecutor, ShardedStatevector

    n = int(name[4:-3])
    rank, world, dist, eng = ctx.rank, ctx.world, ctx.dist, ctx.eng
    sv = ShardedStatevector(GpuExecutor(eng), dist)
    rng = np.random.default_rng(4)
    obs = F.tfim_observables(list(range(n)), n)
    # timed circuit i has first_trotter + (i mod 10) Trotter steps whatever the warm-up count (the
    # warm-up reuses the first timed circuits), so runs with equal --steps are comparable across GPU counts
    Js = [float(rng.uniform(0, 1)) for _ in range(steps)]
    trot = [first_trotter + i % 10 for i in range(steps)]
    timed = [F.tfim_circuit(n, trot[i], Js[i], dt=0.25) for i in range(steps)]
    circs = [timed[i % steps] for i in range(warmup)] + timed

    for c in circs[:warmup]:
        sv.estimate(c, obs)
    _park_setup_objects()
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0 and clocks:
        sampler.start()
    ctx.barrier()
    t0 = time.perf_counter()
    dev_ms = sweep_ms = exch_ms = fused_ms = 0.0
    swept = exch_bytes = launches = n_exch = n_fused = n_fused_sweeps = n_sweeps_all = 0
    all_vals = []
    for c in circs[warmup:]:
        vals = sv.estimate(c, obs, profile=True)
        all_vals.append(vals)
        pl = sv.last_plan
        dev_ms += pl["ms_total"]; sweep_ms += pl["ms"].get("sweeps", 0.0); exch_ms += pl["ms"].get("exchange", 0.0)
        swept += pl["kernel_bytes"] - 16 * (1 << pl["n_local"]) * pl["n_expval_passes"]  # sweeps only, live tiles
        exch_bytes += pl["exchanged_bytes_per_rank"]; n_exch += pl["n_exchanges"]
        fused_ms += pl["ms"].get("sweeps+exchange", 0.0); n_fused += pl.get("fused_exchanges", 0)
        n_fused_sweeps += pl.get("n_sweeps_in_fused_segments", 0); n_sweeps_all += pl["n_sweeps"]
        launches += pl["n_sweeps"] + 2 * pl["n_expval_passes"] + pl["n_exchanges"] - pl.get("fused_exchanges", 0)
    ctx.barrier()
    wall_s = time.perf_counter() - t0
    clk = sampler.stop() if (rank == 0 and clocks) else None
    t_dev, t_wall = ctx.max(dev_ms / 1e3), ctx.max(wall_s)
    # sharded == unsharded: rank 0 re-runs the first timed circuit on ONE GPU (the batched wide
    # path of bwq_sv_run: 16 B x 2^n state on this GPU) and compares the values
    diff_1rank = None
    if world > 1:
        if rank == 0:
            v1, st = eng.run_sv(engine.encode_batch([timed[0]], [obs]))
            diff_1rank = float(np.max(np.abs(v1 - all_vals[0]))) if not st.any() else None
        ctx.barrier()
    if rank != 0:
        return None
    # fused exchanges (the last sweep of a segment stores into the peers' new shards): the plain
    # segments give the time of an ordinary sweep; what a fused segment takes beyond its sweep count
    # x that time is the exchange's cost, and the pushing sweeps' own time carries the NVLink bytes
    n_plain = n_sweeps_all - n_fused_sweeps
    ms_per_plain_sweep = sweep_ms / n_plain if n_plain > 0 else 0.0
    fused_overhead_ms = max(0.0, fused_ms - n_fused_sweeps * ms_per_plain_sweep) if n_fused else 0.0
    pushed_ms = fused_overhead_ms + n_fused * ms_per_plain_sweep
    xfer_ms = exch_ms + pushed_ms   # time during which exchange bytes move
    swept_plain = swept * (n_plain / n_sweeps_all) if n_sweeps_all else swept
    achieved = swept_plain / (sweep_ms / 1e3) / 1e9 if sweep_ms > 0 else 0.0
    out = {
        "value": steps / t_dev, "unit": UNIT, "ms_per_step": 1e3 * t_dev / steps, "scaling": "strong", "steps": steps, "warmup": warmup,
        "config": {"workload": name, "n_qubits": n, "observables_per_circuit": len(obs),
                   "trotter_steps": "timed circuit i has %d + (i mod 10) steps: " % first_trotter + ",".join(map(str, trot)),
                   "parallelism": f"amplitude-sharded x{world} (rank = top {world.bit_length() - 1} index bits)",
                   "l2": "shard of %.2f GiB >> 126 MB L2 (no flush needed)" % (16 * 2 ** n / world / 2 ** 30),
                   "timing": "CUDA events on torch's stream, first segment to all_reduce (max over ranks)"},
        "e2e": {"value": steps / t_wall, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8 * len(obs),
                "ms_per_step": 1e3 * t_wall / steps, "note": "circuit object -> plan -> upload -> run -> values on host"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": ctx.peak, "unit": "GB/s", "frac": achieved / ctx.peak, "traffic": None,
                     "kernel": "sv_sweep_kernel", "bytes_per_launch": 2 * 16 * 2 ** n / world,
                     "sweep_share_of_step": sweep_ms / dev_ms if dev_ms else None},
        "exchange": {"exchanges_per_circuit": n_exch / steps, "fused_into_sweeps_per_circuit": n_fused / steps,
                     "bytes_per_rank_per_step": exch_bytes // steps,
                     "ms_per_step": (exch_ms + fused_overhead_ms) / steps,
                     "share_of_step": (exch_ms + fused_overhead_ms) / dev_ms if dev_ms else None,
                     "GBps_per_rank": (exch_bytes / (xfer_ms / 1e3) / 1e9) if xfer_ms > 0 else None,
                     "nvlink_reference_GBps": 770.0,
                     "impl": (("P2P stores from sv_sweep_kernel (bwq_svx_run_segment_push); " if n_fused else "") + sv.last_plan["exchange_impl"]) if world > 1 else None,
                     "definition": "ms_per_step = separate exchange kernels + (fused segments' time - their sweep count x the mean time of a "
                                   "plain sweep); GBps = bytes / (exchange kernels + pushing sweeps)"},
        "max_abs_diff_vs_1rank": diff_1rank, "last_values_head": [float(x) for x in all_vals[-1][:3]],
    }
    if clk is not None:
        out["clocks"] = clk
    if cpu_check and world == 1:
        # CPU parity + rate at a width the host finishes in seconds: the same circuit family through
        # the same wide-statevector kernels at 22 qubits against the C++ restatement
        from oracle import cpu_ref

        cpu_ref.build()
        m = 22
        c_small = [F.tfim_circuit(m, trot[i], Js[i], dt=0.25) for i in range(min(2, steps))]
        fb = engine.encode_batch(c_small, [F.tfim_observables(list(range(m)), m)] * len(c_small))
        cpu_ref.prepare(fb)
        cores = host_threads()
        t = time.perf_counter()
        ref, st = cpu_ref.run_sv(fb, threads=cores)
        dt = time.perf_counter() - t
        got, st2 = eng.run_sv(fb)
        out["cpu_baseline"] = {"value": len(c_small) / dt, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"{len(c_small)} circuits of the same family at {m} qubits (2^-{n - m} of the amplitudes; "
                                         f"a {n}-qubit state does not fit the CPU budget), ideal SV, {dt:.2f} s on {cores} threads",
                               "max_abs_diff_vs_gpu": float(np.max(np.abs(got - ref))) if not (st.any() or st2.any()) else None}
        out["max_abs_diff_vs_cpu"] = out["cpu_baseline"]["max_abs_diff_vs_gpu"]
    return out


# ----------------------------------------------------------------------------------------------
# density-matrix workloads (cfg1, cfg2, cfg3, cfg5-generation): circuit-sharded, weak scaling
# ----------------------------------------------------------------------------------------------
def bench_dm(ctx, name, steps, warmup, scale=1.0, cpu_budget=15.0, cpu_first=None, clocks=False):
    import torch

    from ml_qem_b200 import engine, noise

    args, rank, world, dist, eng = ctx.args, ctx.rank, ctx.world, ctx.dist, ctx.eng
    wl = build_workload(name, rank, scale)
    t_enc = time.perf_counter()
    batch = engine.encode_batch(wl["circuits"], wl["observables"])
    t_enc = time.perf_counter() - t_enc
    n_circ = batch.n_circuits
    # lowering threads: the ranks of one box share its host cores
    eng.set_options(tile_qubits=args.tile_qubits, low_qubits=args.low_qubits, chunk_circuits=args.chunk_circuits,
                    host_threads=max(1, host_threads() // world) if world > 1 else 0, flags=args.flags)
    eng.set_noise(noise.from_backend(wl["backend"]))

    _park_setup_objects()
    # ---- resident-program throughput (`value`)
    st = eng.prepare_dm(batch)
    assert not st.any(), "lowering failed"
    st = eng.prepare_sv(batch)
    assert not st.any()
    for _ in range(warmup):
        noisy = eng.execute_dm()
        ideal = eng.execute_sv()
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0 and clocks:
        sampler.start()
    ctx.barrier()
    t0 = time.perf_counter()
    dev_ms = sweep_ms = 0.0
    launches = swept = sweeps = 0
    for _ in range(steps):
        noisy = eng.execute_dm()
        s = eng.stats()
        dev_ms += s["kernel_ms"]; sweep_ms += s["sweep_kernel_ms"]
        launches += s["n_sweep_launches"] + s["n_other_launches"]
        swept += s["state_bytes_swept"]; sweeps += s["n_state_sweeps"]
        n_sweep_launches = s["n_sweep_launches"]; n_passes = s["n_passes"]; n_tma_launches = s["n_tma_sweep_launches"]
        ideal = eng.execute_sv()
        s = eng.stats()
        dev_ms += s["kernel_ms"]
        launches += s["n_other_launches"]
    ctx.barrier()
    wall_s = time.perf_counter() - t0
    clk = sampler.stop() if (rank == 0 and clocks) else None
    # device time (CUDA events on the engine's stream, summed over the K steps), max over ranks
    t_dev = ctx.max(dev_ms / 1e3)
    t_wall = ctx.max(wall_s)
    total_circ = ctx.sum(float(n_circ)) * steps
    value = total_circ / t_dev

    # ---- end to end through the C ABI with host buffers (`e2e`)
    # one call per step: bwq_meas_data_run = (ideal, noisy) of every circuit from host buffers --
    # lowering, H2D of the programs, kernels, D2H of the values all inside the timed region
    ideal_e, noisy_e, st2, st1 = eng.run_meas_data(batch)
    sv_io = eng.run_sv(batch) and eng.stats()  # program/value bytes of the statevector side (untimed probe)
    for _ in range(min(warmup, 2)):
        eng.run_meas_data(batch)
    ctx.barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    oc_kernel_ms = pre_ms = call_ms = 0.0
    for _ in range(steps):
        ideal_e, noisy_e, st2, st1 = eng.run_meas_data(batch)
        s = eng.stats()
        pre_ms += s["host_pre_ms"]; call_ms += s["call_wall_ms"]
        onchip = s["n_onchip_circuits"] > 0   # dm_onchip_kernel took the batch: one launch, both sides, no lowering
        oc_kernel_ms += s["kernel_ms"]
        h2d += s["h2d_bytes"] + (0 if onchip else sv_io["h2d_bytes"]); d2h += s["d2h_bytes"] + (0 if onchip else sv_io["d2h_bytes"])
    ctx.barrier()
    e2e_s = ctx.max(time.perf_counter() - t0)
    e2e_value = total_circ / e2e_s
    if onchip:
        # the host-buffer call runs these circuits on dm_onchip_kernel (raw gate stream, one warp per circuit); the
        # prepared path above is the lowering + tile-sweep path: two implementations, equal to rounding
        assert np.max(np.abs(noisy_e - noisy)) <= 1e-12 and np.max(np.abs(ideal_e - ideal)) <= 1e-12, "on-chip and tile-sweep paths differ"
        # `value` for this workload: the on-chip launch alone (CUDA events around it inside the library, the raw
        # batch already in HBM), since that is the kernel the product path runs
        value_tile_sweep = value
        t_dev = ctx.max(oc_kernel_ms / 1e3)
        value = total_circ / t_dev
        launches = steps
    else:
        assert np.array_equal(noisy_e, noisy) and np.array_equal(ideal_e, ideal), "resident and host-buffer paths differ"

    # ---- the same work through the reference-facing API from BASE circuits: B200Estimator.run(circuits,
    # observables, variants=...) -- Python normalisation, encoding of the base circuits, variant
    # generation (folds / twirls) inside the library, kernels, values back; plus the ideal estimator
    est_res = None
    if wl.get("base") is not None:
        from ml_qem_b200.engine import Variants
        from ml_qem_b200.estimator import B200Estimator

        V = Variants(**wl["variants"])
        base_c, obs0 = wl["base"], wl["observables"][0]
        pairs_c = [c for c in base_c for _ in obs0]
        pairs_o = [o for _ in base_c for o in obs0]
        est_n, est_i = B200Estimator(backend=wl["backend"], engine=eng), B200Estimator(engine=eng)
        for _ in range(2):
            rn = est_n.run(pairs_c, pairs_o, variants=V).result()
            ri = est_i.run(pairs_c, pairs_o).result()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            rn = est_n.run(pairs_c, pairs_o, variants=V).result()
            ri = est_i.run(pairs_c, pairs_o).result()
        ctx.barrier()
        est_s = ctx.max(time.perf_counter() - t0)
        n_var_circ = ctx.sum(float(len(base_c) * V.n_variants)) * steps
        est_res = {"value": n_var_circ / est_s, "unit": UNIT, "ms_per_step": 1e3 * est_s / steps,
                   "call": "B200Estimator(backend).run(base circuits x observables, variants=Variants(%s)) + B200Estimator().run(...) "
                           "(ideal); variants generated inside the library (bwq_dm_run_variants)" % wl["variants"],
                   "base_circuits": len(base_c), "variants_per_circuit": V.n_variants, "pairs_per_call": len(pairs_c),
                   "values_head": [float(x) for x in rn.values[:2]], "ideal_head": [float(x) for x in ri.values[:2]]}
        # the C ABI alone from the BASE batch (host buffers): bwq_meas_data_run_variants = variant generation in
        # the library + lowering + H2D + kernels (noisy: every variant, ideal: base circuits) + D2H
        nm_v = noise.from_backend(wl["backend"])
        base_batch = engine.encode_batch(base_c, [obs0] * len(base_c))
        for _ in range(2):
            eng.run_meas_data_variants(base_batch, V, noise=nm_v)
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            eng.run_meas_data_variants(base_batch, V, noise=nm_v)
        ctx.barrier()
        var_s = ctx.max(time.perf_counter() - t0)
        est_res["c_abi_variants"] = {"value": n_var_circ / var_s, "unit": UNIT, "ms_per_step": 1e3 * var_s / steps,
                                     "call": "bwq_meas_data_run_variants(base batch, variants) from host buffers",
                                     "h2d_base_batch_bytes": int(base_batch.nbytes())}
        eng.set_noise(nm_v)

    # ---- final gather of the labels (the only collective of the density-matrix path)
    if dist is not None:
        mine = torch.from_numpy(np.stack([noisy, ideal])).cuda()
        parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, parts, dst=0)
    if rank != 0:
        return None

    # ---- roofline of the dominant kernel (dm_sweep): algorithmic bytes = 2 x 8 B x 4^n per state
    # sweep (one read + one write of every Pauli-basis element), time = CUDA events around the
    # sweep launches (rank 0)
    achieved = swept / (sweep_ms / 1e3) / 1e9 if sweep_ms > 0 else 0.0
    on_chip = wl["desc"]["n_qubits"] <= 6
    roofline = {"bound": "hbm", "achieved": achieved, "peak": ctx.peak, "unit": "GB/s", "frac": achieved / ctx.peak,
                "traffic": None, "kernel": ("dm_sweep_tma_kernel<false> (TMA tile load/store)" if n_tma_launches else
                           "dm_sweep_kernel<%d,false>" % min(6, wl["desc"]["n_qubits"])),
                "peak_source": ctx.peak_src,
                "bytes_definition": "actual layout: 2 x 8 B x 4^n per state sweep (real Pauli-basis elements); "
                                    "SURVEY 8(d) counts the reference's complex128 layout, 2 x 16 B x 4^n",
                "achieved_survey_units": 2.0 * achieved, "frac_survey_units": 2.0 * achieved / ctx.peak,
                "bytes_per_launch": swept / max(1, steps * n_sweep_launches),
                "launches_per_step": n_sweep_launches, "state_sweeps_per_step": sweeps // steps,
                "register_passes_per_step": n_passes,
                "sweep_share_of_step": sweep_ms / dev_ms if dev_ms else None}
    if on_chip:
        roofline["note"] = "state of <= 6 qubits never leaves the SM (one launch per circuit batch): HBM fraction is not the bound here"
    if onchip:
        # algorithmic HBM bytes of the on-chip launch: the raw batch read once + the values written
        oc_bytes = float(batch.nbytes() + 8 * (noisy.size + ideal.size))
        roofline.update({"kernel": "dm_onchip_kernel", "achieved": oc_bytes * steps / (oc_kernel_ms / 1e3) / 1e9,
                         "bytes_per_launch": oc_bytes, "launches_per_step": 1, "sweep_share_of_step": 1.0,
                         "bytes_definition": "raw gate stream + observables read once, values written (the state stays in shared memory)",
                         "tile_sweep_path": {"value": value_tile_sweep, "achieved": achieved, "kernel": roofline["kernel"],
                                             "note": "same batch through lowering + dm_sweep_kernel (bwq_dm_prepare/execute), kept for comparison"},
                         "note": "one warp interprets the gate stream of one circuit, state in shared memory: latency/issue bound, "
                                 "HBM traffic is the 8-byte ops only"})
        roofline["frac"] = roofline["achieved"] / ctx.peak
        roofline.pop("achieved_survey_units", None); roofline.pop("frac_survey_units", None)
        try:  # SURVEY 8(d), n <= 6 regime: report the FP64-pipe fraction (from the committed ncu capture of this kernel)
            oc = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name, {})
            if "fp64_pipe_pct" in oc:
                roofline["on_chip_pipes_ncu"] = {k: oc[k] for k in ("fp64_pipe_pct", "issue_active_pct", "warps_active_pct") if k in oc}
                roofline["on_chip_pipes_source"] = "profiles/r2/ncu_dm_onchip_cfg1.txt (ncu --set full of this kernel on this workload)"
        except Exception:  # noqa: BLE001
            pass
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_path):
        try:
            ratio = json.load(open(traffic_path)).get(name, {}).get("traffic_over_algorithmic")
            if ratio is not None:  # dram bytes per launch, scaled from the ncu capture to this launch size
                roofline["traffic"] = ratio * roofline["bytes_per_launch"]
                roofline["traffic_source"] = "ncu --set full capture (profiles/traffic.json), scaled to this launch size"
        except Exception:  # noqa: BLE001
            pass

    cpu = None
    if cpu_budget > 0 and world == 1:
        rate, cores, sample, (fb_s, v_dm, v_sv), kind = cpu_reference_rate(name, budget_s=cpu_budget, wl=wl if scale == 1.0 else None,
                                                                           first=cpu_first)
        # the CPU sample doubles as a parity check of this very run
        k = fb_s.n_observables
        err = max(float(np.max(np.abs(noisy[:k] - v_dm))), float(np.max(np.abs(ideal[:k] - v_sv))))
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
               "max_abs_diff_vs_gpu": err}

    out = {
        "value": value, "unit": UNIT, "ms_per_step": 1e3 * t_dev / steps, "scaling": "weak", "steps": steps, "warmup": warmup,
        "config": dict(wl["desc"], parallelism=f"circuit-sharded x{world}", state_layout="Pauli-basis density matrix, 8 B/element",
                       l2="per-rank working set %.1f GiB of resident states >> 126 MB L2 (no flush needed)" %
                          (wl["desc"].get("resident_state_bytes", n_circ * 8 * 4 ** wl["desc"]["n_qubits"]) / 2 ** 30),
                       timing="CUDA events on the engine stream summed over steps (max over ranks); wall %.1f ms/step" %
                              (1e3 * t_wall / steps),
                       host_encode_ms=1e3 * t_enc),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // steps, "d2h_bytes_per_step": d2h // steps,
                "ms_per_step": 1e3 * e2e_s / steps,
                # of which the GPU was busy (CUDA events inside the call, first launch to last value copy): the rest is host
                "device_ms_per_step": oc_kernel_ms / steps,
                # density-matrix side inside the library: entry -> first enqueue, and the whole C call (pipelined runs only)
                "host_pre_ms_per_step": pre_ms / steps, "c_call_ms_per_step": call_ms / steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "max_abs_diff_vs_cpu": cpu["max_abs_diff_vs_gpu"] if cpu else None,
    }
    if est_res is not None:
        out["e2e_estimator"] = est_res
    if clk is not None:
        out["clocks"] = clk
    return out


# sub-workloads reported next to the headline line: the other BASELINE configs that fit the run
# (name, steps, warmup, kwargs)
SUB_WORKLOADS = [
    ("tfim4_lima_zne", 3, 1, {"cpu_budget": 6.0}),
    ("tfim14_dm", 2, 1, {"cpu_budget": 12.0, "cpu_first": 1}),
    ("tfim30_sv", 3, 1, {"first_trotter": 3}),
]


def sub_summary(d):
    """Compact form of a workload result for the `workloads` object of the headline line."""
    if d is None:
        return None
    r = d["roofline"]
    out = {"value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"], "steps": d["steps"], "warmup": d["warmup"],
           "scaling": d["scaling"], "e2e": d["e2e"]["value"],
           "roofline": {"achieved": r["achieved"], "frac": r["frac"], "kernel": r["kernel"], "unit": r["unit"],
                        "bytes_per_launch": r["bytes_per_launch"], "sweep_share_of_step": r.get("sweep_share_of_step"),
                        **{k: r[k] for k in ("on_chip_pipes_ncu", "note", "tile_sweep_path") if k in r}},
           "max_abs_diff_vs_cpu": d.get("max_abs_diff_vs_cpu"), "gpu_launches": d["gpu_launches"],
           "config": d["config"]}
    if d.get("cpu_baseline"):
        out["cpu_baseline"] = {k: d["cpu_baseline"][k] for k in ("value", "cores", "kind", "sample")}
    for k in ("exchange", "max_abs_diff_vs_1rank", "e2e_estimator"):
        if k in d:
            out[k] = d[k]
    return out


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="brick10_guadalupe_twirl")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the named batch size (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-workloads", action="store_true", help="only the headline workload (profiling runs)")
    ap.add_argument("--tile-qubits", type=int, default=0)
    ap.add_argument("--low-qubits", type=int, default=0)
    ap.add_argument("--chunk-circuits", type=int, default=0)
    ap.add_argument("--flags", type=int, default=0, help="BWQ_OPT_* bits (kernel experiments)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = args.steps, max(args.warmup, 0)

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        t_all = time.perf_counter()
        aer = probe_qiskit_aer()
        rates = []
        for i in range(warmup + steps):
            # each step = one bounded sample; budget so that the whole run ends within minutes
            rate, cores, sample, _, kind = cpu_reference_rate(args.workload, budget_s=10.0)
            if i >= warmup:
                rates.append(rate)
        value = float(np.mean(rates))
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": 1e3 * (time.perf_counter() - t_all) / max(1, warmup + steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "note": "real qiskit-aer" if kind == "reference" else
                       "Aer-style C++/OpenMP restatement (oracle/cpu_ref.cpp); qiskit-aer is not installable offline",
                       "qiskit_aer_probe": "found" if aer else "absent",
                       "host_threads": cores, "env_OMP_NUM_THREADS_ignored": os.environ.get("OMP_NUM_THREADS")},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (B200)
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback (use --impl reference for the CPU arm)")
    ctx = Ctx(args)
    is_sv = args.workload.startswith("tfim") and args.workload.endswith("_sv")
    cpu_budget = 0.0 if args.no_cpu_baseline else 15.0
    if is_sv:
        main_res = bench_sharded_sv(ctx, args.workload, steps, warmup, cpu_check=not args.no_cpu_baseline, clocks=True)
    else:
        main_res = bench_dm(ctx, args.workload, steps, warmup, args.scale, cpu_budget=cpu_budget, clocks=True)

    subs = {}
    if not args.no_sub_workloads and args.workload == "brick10_guadalupe_twirl" and args.scale == 1.0:
        for name, s_steps, s_warm, kw in SUB_WORKLOADS:
            t_sub = time.perf_counter()
            try:
                if name.endswith("_sv"):
                    r = bench_sharded_sv(ctx, name, s_steps, s_warm, cpu_check=not args.no_cpu_baseline, **kw)
                else:
                    kw = dict(kw)
                    if args.no_cpu_baseline:
                        kw["cpu_budget"] = 0.0
                    r = bench_dm(ctx, name, s_steps, s_warm, **kw)
                if ctx.rank == 0:
                    subs[name] = sub_summary(r)
                    subs[name]["wall_s_incl_setup"] = time.perf_counter() - t_sub
            except Exception as exc:  # noqa: BLE001 - a failing sub-workload must not void the headline line
                if ctx.world > 1:
                    raise
                subs[name] = {"error": repr(exc)[:300]}

    if ctx.rank != 0:
        ctx.close()
        return 0
    line = {
        "metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": steps, "warmup": warmup,
        "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": main_res["scaling"], "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": main_res["config"], "e2e": main_res["e2e"],
        "gpu_launches": main_res["gpu_launches"], "roofline": main_res["roofline"], "cpu_baseline": main_res.get("cpu_baseline"),
        "clocks": main_res.get("clocks"),
    }
    for k in ("exchange", "max_abs_diff_vs_1rank", "last_values_head", "e2e_estimator"):
        if k in main_res:
            line[k] = main_res[k]
    if subs:
        line["workloads"] = subs
    print(json.dumps(line))
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
