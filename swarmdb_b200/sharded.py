"""Multi-GPU plumbing: one process per GPU, agents hash-partitioned across the shards.

Replaces the reference's partitioned Kafka topic (`_get_partition`, M:309-312: Python's salted
`hash(agent_id) % num_partitions`, different on every process start) with a deterministic
FNV-1a-64 hash, and the broker round trip with one NCCL all-gather of per-SEND wire batches over
NVLink (torch.distributed is only the transport; expansion into per-recipient records happens
in the sm_100a kernels of each shard - see csrc/sdb_xshard.cu).
"""
from __future__ import annotations

from typing import Iterable, Optional

import numpy as np

FNV_OFFSET = 0xCBF29CE484222325
FNV_PRIME = 0x100000001B3
_MASK = (1 << 64) - 1


def fnv1a64(data: bytes) -> int:
    h = FNV_OFFSET
    for b in data:
        h = ((h ^ b) * FNV_PRIME) & _MASK
    return h


def shard_of_agent(agent_id: str, num_shards: int) -> int:
    """Owner shard of an agent id (deterministic replacement of M:312)."""
    return fnv1a64(agent_id.encode("utf-8")) % num_shards


def shard_map_for_names(names: Iterable[str], num_shards: int) -> np.ndarray:
    return np.fromiter((shard_of_agent(n, num_shards) for n in names), dtype=np.uint8)


def shard_map_numbered(prefix: str, width: int, n: int, num_shards: int) -> np.ndarray:
    """Vectorised fnv1a64(f"{prefix}{i:0{width}d}") % num_shards for i in range(n) (bench workloads)."""
    h = np.full(n, FNV_OFFSET, dtype=np.uint64)
    prime = np.uint64(FNV_PRIME)
    with np.errstate(over="ignore"):
        for b in prefix.encode("utf-8"):
            h = (h ^ np.uint64(b)) * prime
        idx = np.arange(n, dtype=np.uint64)
        for k in range(width - 1, -1, -1):
            digit = (idx // np.uint64(10 ** k)) % np.uint64(10)
            h = (h ^ (digit + np.uint64(48))) * prime
    return (h % np.uint64(num_shards)).astype(np.uint8)


class ShardExchange:
    """Owns the wire buffers of one rank and moves them between ranks.

    `backend` abstracts where the bytes live: the CUDA implementation keeps them in torch CUDA
    tensors and all-gathers with NCCL; tests substitute a CPU backend (gloo) with the same calls.
    """

    def __init__(self, shard, rank: int, world: int, max_sends: int, max_payload: int, backend):
        self.shard, self.rank, self.world = shard, rank, world
        self.wire_bytes = shard.wire_bytes(max_sends, max_payload)
        self.backend = backend
        self.send_buf = backend.alloc(self.wire_bytes)
        self.recv_buf = backend.alloc(self.wire_bytes * world) if world > 1 else self.send_buf

    def export(self, sender, group, prio, typ, lens, payload_off, payload, ts=None) -> None:
        self.shard.export_group_batch(sender, group, prio, typ, lens, payload_off, payload,
                                      self.backend.ptr(self.send_buf), self.wire_bytes, ts)

    def export_mixed(self, sender, kind, target, list_off, list_idx, prio, typ, lens, payload_off, payload, ts=None) -> None:
        """Mixed batch in call order: kind 0 = p2p (target receiver), 1 = group, 2 = broadcast list."""
        self.shard.export_mixed_batch(sender, kind, target, list_off, list_idx, prio, typ, lens, payload_off, payload,
                                      self.backend.ptr(self.send_buf), self.wire_bytes, ts)

    def exchange(self) -> None:
        if self.world > 1:
            self.backend.all_gather(self.recv_buf, self.send_buf)

    def import_all(self) -> int:
        return self.shard.import_wire_batches(self.world, self.backend.ptr(self.recv_buf), self.wire_bytes)

    def step(self, sender, group, prio, typ, lens, payload_off, payload, ts=None) -> int:
        self.export(sender, group, prio, typ, lens, payload_off, payload, ts)
        self.exchange()
        return self.import_all()


class PeerExchange:
    """Peer-memory transport: no collective moves message bytes.

    Every rank exports into its own CUDA-IPC-shared buffer (two of them, alternating per step);
    `exchange()` is a one-element all-reduce used purely as a stream-ordered cross-rank barrier;
    `import_all()` hands the shard the table of peer pointers, and the import + fan-out kernels
    pull descriptors and payloads straight out of the exporting GPUs' memory over NVLink.
    """

    def __init__(self, shard, rank: int, world: int, max_sends: int, max_payload: int, device):
        import torch
        import torch.distributed as dist
        self.shard, self.rank, self.world = shard, rank, world
        self.wire_bytes = shard.wire_bytes(max_sends, max_payload)
        self.mine = [shard.wire_alloc(self.wire_bytes) for _ in range(2)]          # (ptr, ipc handle)
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, [h for _, h in self.mine])
        self.ptrs = [[0] * world for _ in range(2)]
        self.opened = []
        for b in range(2):
            for r in range(world):
                if r == rank:
                    self.ptrs[b][r] = self.mine[b][0]
                else:
                    p = shard.wire_open(handles[r][b])
                    self.ptrs[b][r] = p
                    self.opened.append(p)
        self.flag = torch.zeros(1, device=device)
        self.step_no = 0

    @property
    def cur(self) -> int:
        return self.step_no & 1

    def export(self, sender, group, prio, typ, lens, payload_off, payload, ts=None) -> None:
        self.shard.export_group_batch(sender, group, prio, typ, lens, payload_off, payload,
                                      self.mine[self.cur][0], self.wire_bytes, ts)

    def export_mixed(self, sender, kind, target, list_off, list_idx, prio, typ, lens, payload_off, payload, ts=None) -> None:
        self.shard.export_mixed_batch(sender, kind, target, list_off, list_idx, prio, typ, lens, payload_off, payload,
                                      self.mine[self.cur][0], self.wire_bytes, ts)

    def exchange(self) -> None:
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.flag)                     # stream-ordered barrier: every rank's export is complete

    def import_all(self) -> int:
        base = self.shard.import_wire_ptrs(self.ptrs[self.cur])
        self.step_no += 1                                   # the other buffer is free: its readers passed the barrier above
        return base

    def step(self, *batch, ts=None) -> int:
        self.export(*batch, ts=ts)
        self.exchange()
        return self.import_all()

    def close(self) -> None:
        for p in self.opened:
            self.shard.wire_close(p, True)
        for p, _ in self.mine:
            self.shard.wire_close(p, False)
        self.opened, self.mine = [], []


class TorchCudaBackend:
    """Wire buffers in CUDA memory, exchanged with NCCL (torch.distributed), on torch's current stream."""

    def __init__(self, device):
        import torch
        self.torch, self.device = torch, device

    def alloc(self, nbytes: int):
        return self.torch.zeros(nbytes, dtype=self.torch.uint8, device=self.device)

    def ptr(self, t) -> int:
        return t.data_ptr()

    def all_gather(self, out, inp) -> None:
        import torch.distributed as dist
        dist.all_gather_into_tensor(out, inp)


def run_sharded_bench(args, rank: int, world: int, local_rank: int, wl):
    """N-GPU arm of bench.py (torchrun, one rank per GPU): c3 = c2's workload with the 1M agents
    hash-sharded over the ranks; every rank ingests 65,536 group sends per step (weak scaling),
    all-gathers the wire batches over NVLink and drains the agents it owns."""
    import json
    import os
    import time

    import torch
    import torch.distributed as dist

    from ._native import HDR_DTYPE, Shard

    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream()
    K, W = args.steps, args.warmup
    per_rank_msgs = wl.S * wl.F                     # routed messages each rank's sends produce
    recv_cap = per_rank_msgs * 2 + (1 << 16)        # a shard receives ~1/world of world x that, hash imbalance aside
    # every agent receives from the sends of ALL ranks: mean pending per step = world x 4.19, sized with headroom
    mean_pending = world * wl.S * wl.F / wl.A
    ring_slots = 64
    while ring_slots < 3 * mean_pending + 32:
        ring_slots *= 2
    ring_slots = int(os.environ.get("SDB_RING_SLOTS", ring_slots))
    shard = Shard(max_agents=wl.A, ring_slots=ring_slots, arena_bytes=1 << 33,
                  max_payload_bytes=wl.L, max_groups=1 << 14, member_pool_entries=wl.A + 1024, max_batch_sends=wl.S,
                  max_batch_payload=wl.S * wl.L, max_recv_records=recv_cap, max_recv_payload=recv_cap * wl.L,
                  device=local_rank, shard_id=rank, num_shards=world, fanout_variant=2)
    shard.set_stream(stream.cuda_stream)
    smap = shard_map_numbered("agent_", 7, wl.A, world)
    shard.set_agent_shards(smap)
    local_agents = np.nonzero(smap == rank)[0].astype(np.uint32)
    shard.register(local_agents)
    for g in range(wl.G):
        shard.create_group(g, wl.members(g))
    shard.sync()
    transport = os.environ.get("SDB_XSHARD", "peer")
    with torch.cuda.stream(stream):
        if transport == "peer":
            ex = PeerExchange(shard, rank, world, wl.S, wl.S * wl.L, dev)
        else:
            ex = ShardExchange(shard, rank, world, wl.S, wl.S * wl.L, TorchCudaBackend(dev))
        # each rank draws its own slice of the global batch: advance the generators by rank
        for _ in range(rank):
            wl.batch()
        n_distinct = 2
        batches = []
        for _ in range(n_distinct):
            batches.append(wl.batch())
            for _ in range(world - 1):
                wl.batch()
        # device-resident inputs: wire batches exported once, before the timed region
        wires = []
        if transport == "peer":
            for k, b in enumerate(batches):                 # batch k lives in this rank's export buffer k
                ex.step_no = k
                ex.export(*b)
            ex.step_no = 0
            torch.cuda.synchronize(); dist.barrier()
        else:
            for b in batches:
                ex.export(*b)
                torch.cuda.synchronize()
                wires.append(ex.send_buf.clone())

        phase_ev = []

        def device_step(i, timed=False):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if timed else None
            if timed:
                evs[0].record(stream)
            if transport != "peer":
                ex.send_buf.copy_(wires[i % n_distinct], non_blocking=True)   # HBM -> HBM staging of this step's input
            if timed:
                evs[1].record(stream)
            ex.exchange()
            if timed:
                evs[2].record(stream)
            ex.import_all()
            if timed:
                evs[3].record(stream)
            # non-local agents have empty rings on this shard: draining "all" needs no index upload
            _, total, _ = shard.receive_batch(None, 100, 0, copy_out=False)
            if timed:
                evs[4].record(stream)
                phase_ev.append(evs)
            return total

        for i in range(W):
            device_step(i)
        launches0 = shard.stats()["kernel_launches"]
        shard.profile(True)
        from bench import ALG_BYTES_FANOUT, ClockSampler, hbm_peak, traffic_note
        clocks = ClockSampler(local_rank); clocks.start()
        dist.barrier(); torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        delivered = 0
        for i in range(K):
            delivered += device_step(W + i, timed=True)
        ev1.record(stream)
        torch.cuda.synchronize(); dist.barrier()
        clk = clocks.stop()
        ms = ev0.elapsed_time(ev1)
        prof = shard.profile_read(); shard.profile(False)
        names = ["stage_input", "exchange", "import", "receive"]
        phases = {n: float(np.mean([e[k].elapsed_time(e[k + 1]) for e in phase_ev])) for k, n in enumerate(names)}
        st_end = shard.stats()
        launches = st_end["kernel_launches"] - launches0
        assert st_end["ring_overflow"] == 0, f"ring overflow on rank {rank}: {st_end['ring_overflow']} records (ring_slots={ring_slots})"

        # ---- e2e: host buffers in (export H2D), results out (D2H into pinned buffers)
        pin_hdr = torch.empty(recv_cap * 32, dtype=torch.uint8, pin_memory=True).numpy().view(HDR_DTYPE)
        pin_pay = torch.empty(recv_cap * wl.L, dtype=torch.uint8, pin_memory=True).numpy()
        pinned = []
        for b in batches:
            t = torch.empty(b[6].nbytes, dtype=torch.uint8, pin_memory=True); t.numpy()[:] = b[6]
            pinned.append(b[:6] + (t.numpy(),))
        Ke = max(1, min(K, 8))

        def e2e_step(i):
            ex.step(*pinned[i % n_distinct])
            _, hdr, _ = shard.receive_batch(None, 100, 0, copy_out=True, out_hdr=pin_hdr, out_payload=pin_pay)
            return len(hdr)

        e2e_step(0); e2e_step(1)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = 0
        for i in range(Ke):
            got += e2e_step(i)
        torch.cuda.synchronize(); dist.barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3

    t = torch.tensor([ms, float(delivered), e2e_ms, float(got), float(launches)], dtype=torch.float64, device=dev)
    mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms_max, e2e_max = float(mx[0]), float(mx[2])
        total_delivered, total_got = float(sm[1]), float(sm[3])
        assert int(total_delivered) == K * per_rank_msgs * world, (total_delivered, K * per_rank_msgs * world)
        peak, peak_src = hbm_peak()
        fan_ms, fan_n = prof["fanout"]
        fan_avg = fan_ms / max(fan_n, 1)
        local_msgs = delivered / K
        achieved = ALG_BYTES_FANOUT * local_msgs / (fan_avg * 1e-3) / 1e9 if fan_n else 0.0
        line = {
            "metric": "messages/sec routed (send->receive) at 1M agents, 64-way fanout",
            "value": total_delivered / (ms_max * 1e-3), "unit": "messages/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"c3: c2's workload (1M agents, 15625 groups x 64, 256-byte payloads) with agents "
                                   f"hash-sharded (fnv1a64 % {world}) over {world} GPUs; every rank ingests 65536 group "
                                   f"sends/step; " + ("peer-memory transport: import/fan-out kernels pull descriptors and "
                                   "payloads from the exporting GPUs over NVLink (one-element all-reduce as barrier)"
                                   if transport == "peer" else "wire batches all-gathered over NCCL/NVLink") +
                                   "; each shard drains its agents",
                       "transport": transport,
                       "l2": "inputs larger than L2 (each shard writes and reads back ~1.2 GB of records per step)",
                       "parallelism": f"shard{world}", "ring_slots": ring_slots},
            "clocks": clk,
            "e2e": {"value": total_got / (e2e_max * 1e-3), "unit": "messages/s",
                    "h2d_bytes_per_step": (wl.S * wl.L + wl.S * 64) * world,
                    "d2h_bytes_per_step": int(total_got / Ke * (32 + wl.L)) + wl.A * 4, "steps": Ke,
                    "ms_per_step": e2e_max / Ke},
            "gpu_launches": int(float(sm[4])),
            "roofline": {"kernel": "k_group_fanout", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic_note(), "peak_source": peak_src,
                         "algorithmic_bytes_per_msg": ALG_BYTES_FANOUT, "msgs_per_launch": local_msgs,
                         "ms_per_launch": fan_avg, "note": "rank 0's shard; per-send local fan-out is world-times narrower"},
            "kernels": {k: {"ms_per_launch": v[0] / v[1], "launches": v[1]} for k, v in prof.items() if v[1]},
            "phases_ms_rank0": phases,
            "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    if transport == "peer":
        torch.cuda.synchronize(); dist.barrier()
        ex.close()
    shard.close()
    dist.destroy_process_group()
