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
contiguous blocks, 1 - 1/world of the shard leaves the GPU.  The exchange is the engine's own
kernel: the shards live in symmetric memory and every rank pulls its blocks straight out of the
peers' shards over NVLink (bwq_svx_exchange_push, P2P stores, 705 GB/s per rank on 2 GPUs;
bwq_svx_exchange_pull, P2P loads, 670 GB/s; NCCL all_to_all: 424 GB/s); NCCL ``all_to_all_single`` is the fallback
(BWQ_SVX_EXCHANGE=nccl, or no symmetric memory).  Values are reduced with one ``all_reduce`` of
n_observables doubles.
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

    def _stream(self):
        import torch

        # run on torch's current stream so the segments are ordered with the exchange; the
        # C ABI reads 0 as "the engine's own stream", so torch's default stream (handle 0) is
        # passed as cudaStreamLegacy (0x1)
        return torch.cuda.current_stream(self.device).cuda_stream or 1

    def run_segment(self, program, seg, state, rank, obs):
        program.run_segment(self.engine, seg, state.data_ptr(), rank, obs.data_ptr(), self._stream())

    # ---- EXCHANGE through NVLink peer memory (the engine's own kernel instead of NCCL) -------
    def exchange_buffers(self, n_amp, dist):
        """Two shard buffers in symmetric memory (every rank maps the peers' copies), cached per
        shard size.  Returns None when symmetric memory is unavailable or BWQ_SVX_EXCHANGE=nccl:
        the caller then falls back to ``all_to_all_single``."""
        import os

        import torch

        if os.environ.get("BWQ_SVX_EXCHANGE", "push").lower() == "nccl":  # push (default) | pull | nccl
            return None
        cache = self.__dict__.setdefault("_xbuf", {})
        if n_amp in cache:
            return cache[n_amp]
        try:
            import torch.distributed._symmetric_memory as symm

            group = dist.group.WORLD
            bufs = []
            for _ in range(2):
                raw = symm.empty(2 * n_amp, dtype=torch.float64, device=self.device)
                hdl = symm.rendezvous(raw, group)
                bufs.append((torch.view_as_complex(raw.view(n_amp, 2)), hdl, [int(p) for p in hdl.buffer_ptrs]))
            cache[n_amp] = bufs
        except Exception as exc:  # noqa: BLE001
            import warnings

            warnings.warn(f"symmetric memory unavailable ({exc!r}): EXCHANGE falls back to NCCL all_to_all")
            cache[n_amp] = None
        return cache[n_amp]

    def run_segment_push(self, program, seg, src, dst, rank):
        """SWEEPS segment + the EXCHANGE after it in one go: the last sweep of the segment writes its
        tiles into the peers' new shards (P2P stores from sv_sweep_kernel), barrier after."""
        program.run_segment_push(self.engine, seg, src[0].data_ptr(), rank, dst[2], self._stream())
        dst[1].barrier(channel=0)

    def exchange(self, src, dst, rank, world):
        """src/dst: entries of exchange_buffers().  Cross-rank barrier (every peer has finished
        writing its old shard; it also fences the previous pull out of ``dst``), then the pull."""
        import os

        if os.environ.get("BWQ_SVX_EXCHANGE", "push").lower() != "pull":
            # P2P stores into the peers' new shards, barrier after: all blocks have landed
            self.engine.svx_exchange(src[0].data_ptr(), dst[2], rank, src[0].numel(), self._stream(), push=True)
            dst[1].barrier(channel=0)
            return
        src[1].barrier(channel=0)
        self.engine.svx_exchange(dst[0].data_ptr(), src[2], rank, dst[0].numel(), self._stream())


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
        xbuf = None
        if self.dist and info["n_exchanges"] and hasattr(self.ex, "exchange_buffers"):
            xbuf = self.ex.exchange_buffers(n_amp, self.dist)
        if xbuf is not None:
            cur, other = xbuf
            state, spare = cur[0], other[0]
        else:
            # shard buffers are kept between calls (a 30-qubit shard is a 17 GB allocation)
            cache = self.__dict__.setdefault("_shards", {})
            key = (n_amp, str(dev))
            if key not in cache:
                cache.clear()
                cache[key] = [torch.empty(n_amp, dtype=torch.complex128, device=dev), None]
            if info["n_exchanges"] and cache[key][1] is None:
                cache[key][1] = torch.empty(n_amp, dtype=torch.complex128, device=dev)
            state, spare = cache[key]
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
        import os
        # the EXCHANGE fused into the store of the sweep before it (bwq_svx_run_segment_push).  Measured on
        # tfim30 (profiles/r2/bench_n{2,8}_tfim30_sv_*): 2 GPUs 74.0 -> 68.1 ms per circuit (half of every tile
        # stays local, the rest leaves in runs the sweep produces anyway); 8 GPUs 20.8 -> 22.3 ms (7/8 of the
        # stores cross NVLink as 128-byte runs: 413 GB/s against 696 GB/s of the streaming exchange kernel) --
        # so it is the default on 2 ranks only; BWQ_SVX_FUSED_EXCHANGE=1 / 0 forces it on / off
        want = os.environ.get("BWQ_SVX_FUSED_EXCHANGE", "auto")
        fuse = xbuf is not None and hasattr(self.ex, "run_segment_push") and (want == "1" or (want == "auto" and self.world == 2))
        segs = info["segs"]
        fused = fused_sweeps = 0
        skip = False
        for seg, (kind, first, count, _) in enumerate(segs):
            if skip:  # the EXCHANGE that the previous segment's last sweep already performed
                skip = False
                continue
            if fuse and kind == SEG_SWEEPS and seg + 1 < len(segs) and segs[seg + 1][0] == SEG_EXCHANGE:
                self.ex.run_segment_push(prog, seg, cur, other, self.rank)
                cur, other = other, cur
                state, spare = spare, state
                exchanged_bytes += state.numel() * 16 * (self.world - 1) // self.world
                fused += 1
                fused_sweeps += int(count)
                skip = True
                mark("sweeps+exchange")
                continue
            if kind == SEG_EXCHANGE:
                # block v of rank s <-> block s of rank v: top g local bits swap with the rank bits
                if xbuf is not None:
                    self.ex.exchange(cur, other, self.rank, self.world)
                    cur, other = other, cur
                else:
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
                          "exchanged_bytes_per_rank": exchanged_bytes, "fused_exchanges": fused, "n_sweeps_in_fused_segments": fused_sweeps,
                          "exchange_impl": "own P2P kernel over symmetric memory (bwq_svx_exchange_push/pull)" if xbuf is not None else "nccl all_to_all",
                          "kernel_bytes": prog.algorithmic_bytes(self.rank),
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
