"""world_size-2 gloo test of the multi-GPU host path (swarmdb_b200/sharded.py): shard map,
rank-ordered exchange of wire batches through torch.distributed, global sequence bases.  The
device is stood in for by an oracle-backed fake shard (tests only - the product has no CPU
path); the CUDA import itself is covered by tests/test_gpu_xshard.py."""
import os
import socket

import numpy as np

from swarmdb_b200 import sharded


def test_fnv1a64_known_answers():
    # published FNV-1a 64-bit test vectors
    assert sharded.fnv1a64(b"") == 0xCBF29CE484222325
    assert sharded.fnv1a64(b"a") == 0xAF63DC4C8601EC8C
    assert sharded.fnv1a64(b"foobar") == 0x85944171F73967E8
    m = sharded.shard_map_numbered("agent_", 7, 5000, 8)
    assert all(int(m[i]) == sharded.shard_of_agent(f"agent_{i:07d}", 8) for i in range(0, 5000, 37))
    assert np.bincount(m, minlength=8).min() > 500          # roughly uniform
    assert sharded.shard_of_agent("agent_0000001", 1) == 0


class FakeShard:
    """Oracle-backed stand-in with the Shard methods ShardExchange uses (numpy wire format)."""
    MAGIC = 0x57424453

    def __init__(self, max_agents, max_groups, rank, world, smap):
        from oracle.cpu_ref import CpuOracle
        self.o, self.rank, self.world, self.smap = CpuOracle(max_agents, max_groups), rank, world, smap
        self.full = {}

    def create_group(self, g, members):
        members = np.asarray(members, np.uint32)
        keep = np.nonzero(self.smap[members] == self.rank)[0]
        self.o.create_group_pos(g, members[keep], keep.astype(np.uint32))
        self.full[g] = len(members)

    def wire_bytes(self, max_sends, max_payload):
        return 64 + max_sends * 64 + max_payload + 64

    def export_group_batch(self, sender, group, prio, typ, lens, payload_off, payload, wire, cap, ts=None):
        n = len(sender)
        buf = wire.numpy()
        rec0 = np.zeros(n, np.uint64)
        sizes = np.array([self.full[int(g)] for g in group], np.uint64)
        rec0[1:] = np.cumsum(sizes)[:-1]
        hdr = np.array([self.MAGIC, n, int(sizes.sum()), len(payload)], np.int64)
        parts = [hdr.view(np.uint8), np.asarray(sender, np.uint32).view(np.uint8), np.asarray(group, np.uint32).view(np.uint8),
                 np.asarray(prio, np.uint8), np.asarray(typ, np.uint8), np.asarray(lens, np.uint16).view(np.uint8),
                 np.asarray(payload_off, np.uint64).view(np.uint8), rec0.view(np.uint8), np.asarray(payload, np.uint8)]
        blob = np.concatenate(parts)
        assert len(blob) <= cap
        buf[:len(blob)] = blob

    def import_wire_batches(self, n_src, recv, stride):
        base = self.o.next_seq
        rec_base = 0
        raw = recv.numpy()
        for s in range(n_src):
            b = raw[s * stride:(s + 1) * stride]
            magic, n, total, pbytes = b[:32].view(np.int64)
            assert magic == self.MAGIC
            o = 32
            def take(dt, count):
                nonlocal o
                nb = np.dtype(dt).itemsize * count
                out = b[o:o + nb].view(dt).copy(); o += nb
                return out
            sender, group = take(np.uint32, n), take(np.uint32, n)
            prio, typ, lens = take(np.uint8, n), take(np.uint8, n), take(np.uint16, n)
            poff, rec0, payload = take(np.uint64, n), take(np.uint64, n), take(np.uint8, pbytes)
            self.o.send_group_seq(sender, group, prio, typ, lens, poff, payload, base + rec_base + rec0)
            rec_base += int(total)
        self.o.next_seq = base + rec_base
        return base


class CpuBackend:
    def alloc(self, n):
        import torch
        return torch.zeros(n, dtype=torch.uint8)

    def ptr(self, t):
        return t

    def all_gather(self, out, inp):
        import torch.distributed as dist
        parts = list(out.chunk(dist.get_world_size()))
        dist.all_gather(parts, inp)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A, G, F, S = 512, 12, 24, 60
        rng = np.random.default_rng(7)                       # same stream on every rank
        smap = sharded.shard_map_numbered("agent_", 7, A, world)
        groups = [rng.choice(A, size=F, replace=False) for _ in range(G)]
        shard = FakeShard(A, G, rank, world, smap)
        for g, m in enumerate(groups):
            shard.create_group(g, m)
        ex = sharded.ShardExchange(shard, rank, world, S, S * 64 + 64, CpuBackend())
        from oracle.cpu_ref import CpuOracle
        single = CpuOracle(A, G) if rank == 0 else None
        if single:
            for g, m in enumerate(groups):
                single.create_group(g, m)
        local = np.nonzero(smap == rank)[0].astype(np.uint32)
        streams = {}
        for step in range(3):
            slices = []
            for r in range(world):                            # the global batch, rank-major
                sender, grp = rng.integers(0, A, S), rng.integers(0, G, S)
                prio, typ = rng.integers(0, 4, S), rng.integers(0, 7, S)
                lens = rng.integers(0, 65, S).astype(np.uint16)
                off = np.arange(S, dtype=np.uint64) * 64
                buf = rng.integers(48, 123, S * 64 + 64).astype(np.uint8)
                slices.append((sender, grp, prio, typ, lens, off, buf))
                if single:
                    single.send_group_batch(*slices[-1])
            base = ex.step(*slices[rank])
            bases = [None] * world
            dist.all_gather_object(bases, base)
            assert len(set(bases)) == 1
            cnt, hdr, pay = shard.o.receive_batch(local, 1000)
            pos = 0
            for a, c in zip(local, cnt):
                streams.setdefault(int(a), []).extend(hdr[pos:pos + c].tobytes() for _ in [0]); pos += int(c)
        gathered = [None] * world
        dist.all_gather_object(gathered, streams)
        if rank == 0:
            merged = {}
            for d in gathered:
                merged.update(d)
            cnt, hdr, pay = single.receive_batch(np.arange(A, dtype=np.uint32), 100000)
            pos, want = 0, {}
            for a, c in enumerate(cnt):
                want[a] = hdr[pos:pos + c].tobytes(); pos += int(c)
            ok = all(b"".join(merged.get(a, [])) == want[a] for a in range(A))
            q.put(("ok" if ok else "mismatch", shard.o.next_seq, single.next_seq))
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_matches_single_shard_oracle():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    status, seq_sharded, seq_single = q.get(timeout=5)
    assert status == "ok" and seq_sharded == seq_single


def test_peer_exchange_step_protocol_on_a_recording_shard():
    """Host side of the flag-synchronised transport (sharded.PeerExchange): which buffer / step number every call
    names, the done-wait before a buffer is reused, prefetch bookkeeping and the misuse errors.  No GPU: the shard is a
    recorder (world = 1, so nothing is exchanged)."""
    import pytest
    from swarmdb_b200.sharded import PeerExchange

    class Rec:
        def __init__(self): self.calls, self.n = [], 0
        def set_stream(self, s): self.calls.append(("set_stream", s))
        def wire_bytes(self, a, b): return 4096
        def wire_alloc(self, nbytes): self.n += 1; return (1000 * self.n, b"h%d" % self.n)
        def wire_wait_done(self, ptrs, nbytes, step): self.calls.append(("wait_done", tuple(ptrs), step))
        def export_group_batch(self, *a): self.calls.append(("export", a[7]))            # a[7] = destination buffer
        def wire_publish(self, ptr, nbytes, step): self.calls.append(("publish", ptr, step))
        def import_prefetch(self, ptrs, nbytes, step): self.calls.append(("prefetch", tuple(ptrs), step))
        def import_wire_ptrs_async(self, ptrs, nbytes, step): self.calls.append(("import", tuple(ptrs), step))
        def wire_close(self, p, opened): self.calls.append(("close", p, opened))

    class Stream:
        cuda_stream = 77
        def synchronize(self): pass

    sh = Rec()
    ex = PeerExchange(sh, 0, 1, 16, 1024, device=None, stream=Stream())
    assert sh.calls[0] == ("set_stream", 77) and ex.mine[0][0] == 1000 and ex.mine[1][0] == 2000
    batch = (None,) * 7
    with pytest.raises(RuntimeError):
        ex.prefetch()                                   # nothing exported yet
    ex.export(*batch)                                   # step 1 uses buffer 1 (step & 1)
    with pytest.raises(RuntimeError):
        ex.export(*batch)                               # step 1 is already out
    ex.prefetch(); ex.prefetch()                        # idempotent
    ex.import_all()
    ex.step(*batch)                                     # step 2: buffer 0
    ex.export(*batch)                                   # step 3: buffer 1 again -> waits for done >= 1 first
    ex.republish()                                      # no-op: already published
    ex.import_all()
    got = [c for c in sh.calls[1:]]
    assert got == [("export", 2000), ("publish", 2000, 1), ("prefetch", (2000,), 1), ("import", (2000,), 1),
                   ("export", 1000), ("publish", 1000, 2), ("import", (1000,), 2),
                   ("wait_done", (2000,), 1), ("export", 2000), ("publish", 2000, 3), ("import", (2000,), 3)]
    ex.close()
    assert ("close", 1000, False) in sh.calls and ("close", 2000, False) in sh.calls
