"""Ideal labels for wide registers: one circuit whose 2^n amplitudes are sharded over the GPUs of
a box (one process per GPU), the only path of the engine with a data-path collective.

Replaces qiskit.primitives.Estimator / Statevector.evolve (shots=None) where the reference labels
circuits exactly -- docs/tutorials/h13_ising_data_gen_tomo.ipynb:811 (16-qubit layout),
docs/tutorials/vqe_data_gen_parallel.py:31 -- at widths a single numpy statevector cannot reach
(BASELINE configs[3]: 30 qubits, 2.15 GB of amplitudes per GPU on 8 GPUs).

The C library plans the circuit (bwq_svx_lower): local tile sweeps, EXCHANGE segments and Z-type
expectation passes.  Rank r owns the amplitudes whose top log2(world) index bits equal r.
Diagonal gates (every TFIM bond: cx rz cx = exp(-i t ZZ)) and controlled gates with a global
control never communicate; a non-diagonal gate on a global qubit is preceded by an EXCHANGE = the
top log2(world) local index bits swap with the rank bits: one ``all_to_all_single`` of 2^g
contiguous blocks (NCCL over NVLink), 1 - 1/world of the shard leaves the GPU.  Values are reduced
with one ``all_reduce`` of n_observables doubles.
"""
import numpy as np

from .engine import SEG_EXCHANGE, SEG_EXPVAL, SEG_SWEEPS, EngineError, SvxProgram, encode_batch


class GpuExecutor:
    """Runs the local segments through the C ABI on torch-owned device memory (zero copy)."""

    def __init__(self, engine):
        import torch

        self.engine = engine
        self.device = torch.device("cuda", engine.device)

    def prepare(self, program):
        program.upload(self.engine)

    def init_state(self, state, rank):
        pass  # the first sweep synthesises |0...0>

    def run_segment(self, program, seg, state, rank, obs):
        import torch

        # run on torch's current stream so the segments are ordered with the NCCL exchange; the
        # C ABI reads 0 as "the engine's own stream", so torch's default stream (handle 0) is
        # passed as cudaStreamLegacy (0x1)
        stream = torch.cuda.current_stream(self.device).cuda_stream or 1
        program.run_segment(self.engine, seg, state.data_ptr(), rank, obs.data_ptr(), stream)


class ShardedStatevector:
    """estimate(circuit, observables) -> values, amplitudes sharded over the default process group.

    ``executor``: GpuExecutor(engine) on a GPU box; tests pass a CPU emulator to exercise the
    orchestration under gloo.  With no process group (or world size 1) everything stays local."""

    def __init__(self, executor, dist=None):
        self.ex = executor
        self.dist = dist if (dist is not None and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.world = self.dist.get_world_size() if self.dist else 1
        self.rank = self.dist.get_rank() if self.dist else 0
        if self.world & (self.world - 1):
            raise ValueError("world size must be a power of two")
        self.n_global = self.world.bit_length() - 1
        self.last_plan = None

    def estimate(self, circuit, observables, tile_bits=0, profile=False):
        """profile=True additionally records CUDA-event times per segment kind in ``last_plan``
        (GPU executor only; adds a synchronisation at the end)."""
        import torch

        batch = encode_batch([circuit], [observables])
        prog = SvxProgram(batch, 0, tile_bits, self.n_global)
        info = prog.info
        if info["status"] != 0:
            raise EngineError(f"statevector planner rejected the circuit (status {info['status']})")
        self.ex.prepare(prog)
        n_amp = 1 << info["n_local"]
        dev = self.ex.device
        state = torch.empty(n_amp, dtype=torch.complex128, device=dev)
        spare = torch.empty(n_amp, dtype=torch.complex128, device=dev) if info["n_exchanges"] else None
        obs = torch.zeros(max(1, info["n_observables"]), dtype=torch.float64, device=dev)
        self.ex.init_state(state, self.rank)
        exchanged_bytes = 0
        profile = profile and dev != "cpu"
        marks = []

        def mark(kind):
            if profile:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((kind, ev))

        mark("begin")
        for seg, (kind, first, count, _) in enumerate(info["segs"]):
            if kind == SEG_EXCHANGE:
                # block v of rank s <-> block s of rank v: top g local bits swap with the rank bits
                self.dist.all_to_all_single(spare, state)
                state, spare = spare, state
                exchanged_bytes += state.numel() * 16 * (self.world - 1) // self.world
                mark("exchange")
            elif kind in (SEG_SWEEPS, SEG_EXPVAL):
                self.ex.run_segment(prog, seg, state, self.rank, obs)
                mark("sweeps" if kind == SEG_SWEEPS else "expval")
        if self.dist:
            self.dist.all_reduce(obs)
            mark("allreduce")
        self.last_plan = {"n_bits": info["n_bits"], "n_local": info["n_local"], "n_sweeps": len(info["sweeps"]),
                          "n_passes": info["n_passes"], "n_exchanges": info["n_exchanges"],
                          "exchanged_bytes_per_rank": exchanged_bytes, "kernel_bytes": prog.algorithmic_bytes(self.rank),
                          "n_expval_passes": int(sum(-(-int(c) // 32) for k, _, c, _ in info["segs"] if k == SEG_EXPVAL))}
        vals = obs[:info["n_observables"]].cpu().numpy()
        if profile:
            torch.cuda.synchronize()
            ms = {}
            for (_, e0), (kind, e1) in zip(marks[:-1], marks[1:]):
                ms[kind] = ms.get(kind, 0.0) + e0.elapsed_time(e1)
            self.last_plan["ms"] = ms
            self.last_plan["ms_total"] = marks[0][1].elapsed_time(marks[-1][1])
        if dev != "cpu":
            prog.close()
        return vals
