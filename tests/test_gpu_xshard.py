"""Cross-shard delivery parity: N shards must yield, agent by agent, exactly
the streams of a 1-shard run over the concatenated batch.  All shards live on ONE GPU here (one
handle each, wire batches concatenated in device memory instead of NCCL-all-gathered) so the
kernels and the import protocol are checked without a multi-GPU box; tests/test_sharded_cpu.py
covers the process-group plumbing and the N-GPU bench exercises real NCCL."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _payloads(rng, n, max_len):
    lens = rng.integers(0, max_len + 1, n).astype(np.uint16)
    stride = (max_len + 31) & ~31
    buf = rng.integers(48, 123, n * stride + 64).astype(np.uint8)
    return lens, np.arange(n, dtype=np.uint64) * stride, buf


def _per_agent(counts, hdr, pay, agents):
    from swarmdb_b200._native import payload_offsets
    off = payload_offsets(hdr)
    out, pos = {}, 0
    raw = pay.tobytes()
    for a, c in zip(agents, counts):
        recs = []
        for r in range(pos, pos + int(c)):
            h = hdr[r]
            recs.append((h.tobytes(), raw[int(off[r]): int(off[r]) + ((int(h["len"]) + 31) // 32) * 32]))
        out[int(a)] = recs
        pos += int(c)
    return out


@pytest.mark.parametrize("world", [1, 2, 4])
def test_sharded_streams_equal_single_shard(world):
    import torch
    from oracle.cpu_ref import CpuOracle
    from swarmdb_b200._native import Shard
    from swarmdb_b200.sharded import shard_map_numbered

    rng = np.random.default_rng(100 + world)
    A, G, F, S = 2048, 40, 48, 300                       # 300 sends per rank
    smap = shard_map_numbered("agent_", 7, A, world)
    groups = [rng.choice(A, size=F, replace=False) for _ in range(G - 1)] + [np.array([7, 7, 9, 11], np.uint32)]
    shards = []
    for r in range(world):
        s = Shard(max_agents=A, max_groups=G, ring_slots=2048, arena_bytes=1 << 27, max_batch_sends=S,
                  max_batch_payload=S * 256 + 64, shard_id=r, num_shards=world, max_recv_records=1 << 18)
        s.set_agent_shards(smap)
        for g, m in enumerate(groups):
            s.create_group(g, m)
        shards.append(s)
    oracle = CpuOracle(A, G)
    for g, m in enumerate(groups):
        oracle.create_group(g, m)
    wire_bytes = shards[0].wire_bytes(S, S * 256 + 64)
    wire = torch.zeros(world * wire_bytes, dtype=torch.uint8, device="cuda")
    all_agents = np.arange(A, dtype=np.uint32)
    for step in range(3):
        for r in range(world):                            # every rank ingests its own slice, in rank order
            sender = rng.integers(0, A, S); grp = rng.integers(0, G, S)
            sender[:5] = 7; grp[:5] = G - 1                # sender inside the duplicate group
            prio = rng.integers(0, 4, S); typ = rng.integers(0, 7, S)
            lens, off, buf = _payloads(rng, S, 256)
            ts = rng.random(S)
            shards[r].export_group_batch(sender, grp, prio, typ, lens, off, buf,
                                         wire.data_ptr() + r * wire_bytes, wire_bytes, ts)
            shards[r].sync()
            oracle.send_group_batch(sender, grp, prio, typ, lens, off, buf, ts)
        if step == 1:                                      # pointer-table import (what the peer-memory transport uses)
            ptrs = [wire.data_ptr() + r * wire_bytes for r in range(world)]
            bases = [s.import_wire_ptrs(ptrs) for s in shards]
        else:
            bases = [s.import_wire_batches(world, wire.data_ptr(), wire_bytes) for s in shards]
        assert len(set(bases)) == 1                        # every shard agrees on the global sequence base
        k = [3, 1000, 1000][step]
        merged = {}
        for r, s in enumerate(shards):
            local = np.nonzero(smap == r)[0].astype(np.uint32)
            merged.update(_per_agent(*s.receive_batch(local, k), local))
        want = _per_agent(*oracle.receive_batch(all_agents, k, rec_cap=1 << 18), all_agents)
        assert merged.keys() == want.keys()
        for a in range(A):
            assert merged[a] == want[a], (step, a)
    assert shards[0].stats()["next_seq"] == oracle.next_seq
    for s in shards:
        st = s.stats()
        assert st["ring_overflow"] == 0 and st["enqueued"] == st["delivered"]
        s.close()


def test_wire_path_equals_direct_path_on_one_shard():
    import torch
    from swarmdb_b200._native import Shard
    rng = np.random.default_rng(5)
    A, G, F, S = 4096, 64, 64, 400
    a = Shard(max_agents=A, max_groups=G, ring_slots=256, arena_bytes=1 << 27, max_batch_sends=S, max_batch_payload=S * 256 + 64)
    b = Shard(max_agents=A, max_groups=G, ring_slots=256, arena_bytes=1 << 27, max_batch_sends=S, max_batch_payload=S * 256 + 64)
    perm = rng.permutation(A)
    for g in range(G):
        a.create_group(g, perm[g * F:(g + 1) * F]); b.create_group(g, perm[g * F:(g + 1) * F])
    wire = torch.zeros(a.wire_bytes(S, S * 256 + 64), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        sender = rng.integers(0, A, S); grp = rng.integers(0, G, S)
        prio = rng.integers(0, 4, S); lens, off, buf = _payloads(rng, S, 256)
        base_a = a.send_group_batch(sender, grp, prio, None, lens, off, buf)
        b.export_group_batch(sender, grp, prio, None, lens, off, buf, wire.data_ptr(), wire.numel())
        assert b.import_wire_batches(1, wire.data_ptr(), wire.numel()) == base_a
    ra, rb = a.receive_batch(None, 1000), b.receive_batch(None, 1000)
    assert np.array_equal(ra[0], rb[0]) and ra[1].tobytes() == rb[1].tobytes() and ra[2].tobytes() == rb[2].tobytes()
    a.close(); b.close()


@pytest.mark.parametrize("world", [1, 3])
def test_sharded_mixed_traffic_equals_single_shard(world):
    """p2p + group + broadcast-list sends interleaved in one batch per rank: global order across kinds."""
    import torch
    from oracle.cpu_ref import CpuOracle
    from swarmdb_b200._native import Shard
    from swarmdb_b200.sharded import shard_map_numbered

    rng = np.random.default_rng(300 + world)
    A, G, S = 600, 10, 150
    smap = shard_map_numbered("agent_", 7, A, world)
    groups = [rng.choice(A, size=int(rng.integers(1, 40)), replace=False) for _ in range(G)]
    cap_pay = S * 128 + 64 + 4 * 2000
    shards = []
    for r in range(world):
        s = Shard(max_agents=A, max_groups=G, ring_slots=4096, arena_bytes=1 << 26, max_batch_sends=S,
                  max_batch_payload=cap_pay, shard_id=r, num_shards=world, max_recv_records=1 << 17,
                  max_payload_bytes=128, list_pool_entries=1 << 16)
        s.set_agent_shards(smap)
        for g, m in enumerate(groups):
            s.create_group(g, m)
        shards.append(s)
    oracle = CpuOracle(A, G)
    for g, m in enumerate(groups):
        oracle.create_group(g, m)
    wire_bytes = shards[0].wire_bytes(S, cap_pay)
    wire = torch.zeros(world * wire_bytes, dtype=torch.uint8, device="cuda")
    all_agents = np.arange(A, dtype=np.uint32)
    for step in range(3):
        for r in range(world):
            kind = rng.integers(0, 3, S).astype(np.uint8)
            sender = rng.integers(0, A, S)
            n_lists = 6
            lists = [rng.choice(A, size=int(rng.integers(0, 120)), replace=False) for _ in range(n_lists)]
            lo = np.zeros(n_lists + 1, np.uint64); lo[1:] = np.cumsum([len(x) for x in lists])
            li = np.concatenate(lists).astype(np.uint32)
            target = np.where(kind == 0, rng.integers(0, A, S), np.where(kind == 1, rng.integers(0, G, S), rng.integers(0, n_lists, S)))
            prio = rng.integers(0, 4, S); typ = rng.integers(0, 7, S)
            lens = rng.integers(0, 129, S).astype(np.uint16)
            off = np.arange(S, dtype=np.uint64) * 128
            buf = rng.integers(48, 123, S * 128 + 64).astype(np.uint8)
            shards[r].export_mixed_batch(sender, kind, target, lo, li, prio, typ, lens, off, buf,
                                         wire.data_ptr() + r * wire_bytes, wire_bytes)
            shards[r].sync()
            for i in range(S):                               # the oracle sees the same sends one by one, in order
                one = slice(i, i + 1)
                if kind[i] == 0:
                    oracle.send_batch(sender[one], target[one], prio[one], typ[one], lens[one], off[one], buf)
                elif kind[i] == 1:
                    oracle.send_group_batch(sender[one], target[one], prio[one], typ[one], lens[one], off[one], buf)
                else:
                    t = int(target[i])
                    oracle.send_list_batch(sender[one], [0, len(lists[t])], lists[t], prio[one], typ[one], lens[one], off[one], buf)
        bases = [s.import_wire_batches(world, wire.data_ptr(), wire_bytes) for s in shards]
        assert len(set(bases)) == 1
        k = [2, 5000, 5000][step]
        merged = {}
        for r, s in enumerate(shards):
            local = np.nonzero(smap == r)[0].astype(np.uint32)
            merged.update(_per_agent(*s.receive_batch(local, k), local))
        want = _per_agent(*oracle.receive_batch(all_agents, k, rec_cap=1 << 18), all_agents)
        for a in range(A):
            assert merged[a] == want[a], (step, a)
    assert shards[0].stats()["next_seq"] == oracle.next_seq
    for s in shards:
        assert s.stats()["ring_overflow"] == 0
        s.close()


@pytest.mark.parametrize("world", [2, 3])
def test_explicit_sequence_bases_equal_replayed_oracle(world):
    """sdb_export_mixed_batch_seq: each rank stamps its own (round, rank, local) sequence numbers at
    export; after import every shard's streams equal an oracle replay that assigns the same numbers."""
    import torch
    from swarmdb_b200._native import Shard
    from swarmdb_b200.sharded import composite_seq, shard_map_numbered
    from tests.fake_shard import OracleShard

    rng = np.random.default_rng(910 + world)
    A, G, S = 500, 8, 120
    smap = shard_map_numbered("agent_", 7, A, world)
    groups = [rng.choice(A, size=int(rng.integers(1, 48)), replace=False) for _ in range(G)]
    cap_pay = S * 128 + 64 + 4 * 512                       # payload + room for the recipient lists
    shards = []
    for r in range(world):
        s = Shard(max_agents=A, max_groups=G, ring_slots=4096, arena_bytes=1 << 26, max_batch_sends=S,
                  max_batch_payload=cap_pay, shard_id=r, num_shards=world, max_recv_records=1 << 17,
                  max_payload_bytes=128, list_pool_entries=1 << 16)
        s.set_agent_shards(smap)
        for g, m in enumerate(groups):
            s.create_group(g, m)
        shards.append(s)
    ref = OracleShard(A, G, 0, 1)
    for g, m in enumerate(groups):
        ref.create_group(g, m)
    wire_bytes = shards[0].wire_bytes(S, cap_pay)
    wire = torch.zeros(world * wire_bytes, dtype=torch.uint8, device="cuda")
    ref_bytes = ref.wire_bytes(S, cap_pay + 8 * 4 * 200)
    ref_wire = np.zeros(world * ref_bytes, np.uint8)
    all_agents = np.arange(A, dtype=np.uint32)
    for step in range(3):
        for r in range(world):
            n = S if (step + r) % 3 else 0                   # an empty batch from some rank in some round
            kind = rng.integers(0, 3, n).astype(np.uint8)
            sender = rng.integers(0, A, n)
            lists = [rng.choice(A, size=int(rng.integers(0, 100)), replace=False) for _ in range(4)]
            lo = np.zeros(5, np.uint64); lo[1:] = np.cumsum([len(x) for x in lists])
            li = np.concatenate(lists).astype(np.uint32)
            target = np.where(kind == 0, rng.integers(0, A, n), np.where(kind == 1, rng.integers(0, G, n), rng.integers(0, 4, n)))
            prio = rng.integers(0, 4, n); typ = rng.integers(0, 7, n)
            lens = rng.integers(0, 129, n).astype(np.uint16)
            off = np.arange(n, dtype=np.uint64) * 128
            buf = rng.integers(48, 123, n * 128 + 64).astype(np.uint8)
            ts = rng.random(n)
            base = composite_seq(step + 1, r)
            shards[r].export_mixed_batch(sender, kind, target, lo, li, prio, typ, lens, off, buf,
                                         wire.data_ptr() + r * wire_bytes, wire_bytes, ts, seq_base=base)
            shards[r].sync()
            ref.export_mixed_batch(sender, kind, target, lo, li, prio, typ, lens, off, buf,
                                   ref_wire[r * ref_bytes:(r + 1) * ref_bytes], ref_bytes, ts, seq_base=base)
        for s in shards:
            s.import_wire_batches(world, wire.data_ptr(), wire_bytes)
        ref.import_wire_batches(world, ref_wire, ref_bytes)
        k = [3, 5000, 5000][step]
        for flags in ((0,) if step != 1 else (1,)):
            merged = {}
            for r, s in enumerate(shards):
                local = np.nonzero(smap == r)[0].astype(np.uint32)
                merged.update(_per_agent(*s.receive_batch(local, k, flags), local))
            want = _per_agent(*ref.o.receive_batch(all_agents, k, flags, rec_cap=1 << 18), all_agents)
            for a in range(A):
                assert merged[a] == want[a], (step, a)
    for s in shards:
        st = s.stats()
        assert st["ring_overflow"] == 0 and st["next_seq"] == ref.o.next_seq
        s.close()


def test_sharded_frontend_two_ranks_on_one_gpu(tmp_path):
    """ShardedSwarmsDB x2 in one process (one handle per emulated rank, wire slots in one CUDA buffer):
    rank-local sends with immediate ids, collective flush phases, owner-local receives - against a
    plain single-handle SwarmsDB fed the same calls in (round, rank, call) order."""
    import torch
    import swarmdb_b200 as sdb
    from swarmdb_b200 import sharded
    from tests.test_sharded_frontend_cpu import AGENTS, _play, _script, _setup, _view

    world = 2
    cfg = sdb.GpuConfig(max_agents=64, max_groups=8, flush_threshold=64, max_payload_bytes=1024, ring_slots=1024,
                        arena_bytes=1 << 24, deterministic_ids=True)
    slots = {}

    class Loop:                                             # the exchange of a 1-process, 1-GPU "cluster"
        def __init__(self, shard, rank, world_, max_sends, max_payload):
            self.shard, self.rank = shard, rank
            self.bytes = shard.wire_bytes(max_sends, max_payload)
            slots.setdefault("wire", torch.zeros(world_ * self.bytes, dtype=torch.uint8, device="cuda"))
            # add_agent_group flushes (collectively, in a real deployment); here the emulated ranks run their
            # setup one after the other, so every slot starts out holding a valid empty batch
            e = np.zeros(0, np.uint32)
            shard.export_mixed_batch(e, e, e, None, None, e, e, e, e, np.zeros(32, np.uint8),
                                     slots["wire"].data_ptr() + rank * self.bytes, self.bytes)
            shard.sync()

        def export_mixed(self, *a, seq_base=0):
            *cols, ts = a
            self.shard.export_mixed_batch(*cols, slots["wire"].data_ptr() + self.rank * self.bytes, self.bytes, ts, seq_base)
            self.shard.sync()

        def exchange(self): pass
        def import_all(self): return self.shard.import_wire_batches(world, slots["wire"].data_ptr(), self.bytes)

    dbs = [sharded.make_sharded_swarmsdb(r, world, exchange_factory=Loop, save_dir=str(tmp_path / f"r{r}"),
                                         auto_save=False, gpu_config=cfg) for r in range(world)]
    single = sdb.SwarmsDB(save_dir=str(tmp_path / "single"), auto_save=False, gpu_config=cfg)
    for d in dbs + [single]:
        _setup(d)
    script = _script(np.random.default_rng(23), world, 6)
    got, sent_ids = {}, []
    for per_rank in script:
        for r, d in enumerate(dbs):
            sent_ids += _play(d, per_rank[r]); _play(single, per_rank[r])
        for d in dbs:
            d._flush_export()
        for d in dbs:
            d._flush_import()
        for a in AGENTS:
            d = dbs[dbs[0].owner(a)]
            got.setdefault(a, []).extend(d.receive_messages(a, 4))
    seen = set()
    for a in AGENTS:
        got[a].extend(dbs[dbs[0].owner(a)].receive_messages(a, 100000))
        want = [_view(m) for m in single.receive_messages(a, 100000)]
        assert [_view(m) for m in got[a]] == want, a
        seen |= {m.id for m in got[a]}
    assert len(sent_ids) == len(set(sent_ids)) and seen <= set(sent_ids) and len(seen) > 50
    with pytest.raises(KeyError):
        dbs[0].send_message(AGENTS[0], "x", "stranger")
    for d in dbs + [single]:
        d.close()


@pytest.mark.parametrize("variant", [2, 3])
def test_many_narrow_sends_with_empty_shares_and_empty_payloads(variant):
    """Every fan-out warp walks more sends than its prefetch rings are deep, and many of them have no
    local recipient (tiny groups, 4 shards) or no payload bytes: slot/phase accounting of the pipelines."""
    import torch
    from oracle.cpu_ref import CpuOracle
    from swarmdb_b200._native import Shard
    from swarmdb_b200.sharded import shard_map_numbered

    world, A, G, S = 4, 4096, 2000, 30000
    rng = np.random.default_rng(77)
    smap = shard_map_numbered("agent_", 7, A, world)
    groups = [rng.choice(A, size=int(rng.integers(1, 4)), replace=False) for _ in range(G)]
    cap = S * 64 + 64
    shards = []
    for r in range(world):
        s = Shard(max_agents=A, max_groups=G, ring_slots=512, arena_bytes=1 << 27, max_batch_sends=S, max_payload_bytes=64,
                  max_batch_payload=cap, shard_id=r, num_shards=world, max_recv_records=1 << 18, fanout_variant=variant)
        s.set_agent_shards(smap)
        for g, m in enumerate(groups):
            s.create_group(g, m)
        shards.append(s)
    oracle = CpuOracle(A, G)
    for g, m in enumerate(groups):
        oracle.create_group(g, m)
    wire_bytes = shards[0].wire_bytes(S, cap)
    wire = torch.zeros(world * wire_bytes, dtype=torch.uint8, device="cuda")
    for step in range(2):
        for r in range(world):
            sender = rng.integers(0, A, S); grp = rng.integers(0, G, S)
            prio = rng.integers(0, 4, S); typ = rng.integers(0, 7, S)
            lens = rng.integers(0, 65, S).astype(np.uint16)
            lens[rng.random(S) < 0.3] = 0                       # a third of the sends carry no payload at all
            off = np.arange(S, dtype=np.uint64) * 64
            buf = rng.integers(48, 123, S * 64 + 64).astype(np.uint8)
            shards[r].export_group_batch(sender, grp, prio, typ, lens, off, buf, wire.data_ptr() + r * wire_bytes, wire_bytes)
            shards[r].sync()
            oracle.send_group_batch(sender, grp, prio, typ, lens, off, buf)
        for s in shards:
            s.import_wire_batches(world, wire.data_ptr(), wire_bytes)
        for r, s in enumerate(shards):
            local = np.nonzero(smap == r)[0].astype(np.uint32)
            cg, hg, pg = s.receive_batch(local, 1000)
            cc, hc, pc = oracle.receive_batch(local, 1000, rec_cap=1 << 18)
            assert np.array_equal(cg, cc) and hg.tobytes() == hc.tobytes() and pg.tobytes() == pc.tobytes(), (step, r)
    for s in shards:
        assert s.stats()["ring_overflow"] == 0
        s.close()


# ------------------------------------------------------------------------------------------------
# flag-synchronised, host-asynchronous import (sdb_import_wire_ptrs_async): shards emulated on ONE GPU, every shard
# with its own stream, synchronised only through the ready / done counters at the end of the export buffers
# ------------------------------------------------------------------------------------------------
def _async_cluster(world, A, G, S, groups, max_payload=256, **kw):
    from swarmdb_b200._native import Shard
    from swarmdb_b200.sharded import shard_map_numbered
    smap = shard_map_numbered("agent_", 7, A, world)
    shards = []
    for r in range(world):
        s = Shard(max_agents=A, max_groups=G, ring_slots=1024, arena_bytes=kw.get("arena_bytes", 1 << 27), max_batch_sends=S,
                  max_batch_payload=S * max_payload + 64, shard_id=r, num_shards=world, max_recv_records=1 << 18,
                  max_payload_bytes=max_payload, list_pool_entries=1 << 16)
        s.set_agent_shards(smap)
        for g, m in enumerate(groups):
            s.create_group(g, m)
        shards.append(s)
    wb = shards[0].wire_bytes(S, S * max_payload + 64 + 4 * 4096)
    bufs = [[shards[r].wire_alloc(wb)[0] for r in range(world)] for _ in range(2)]      # bufs[parity][rank]
    return smap, shards, wb, bufs


@pytest.mark.parametrize("world", [1, 2, 4])
def test_async_import_group_traffic_equals_single_queue(world):
    """The fast shape (exclusive groups, payloads <= 512 B): the import is placed entirely on the device - fused
    localize + span fan-out + group-parallel index build - and the host never learns the arena position in between."""
    from oracle.cpu_ref import CpuOracle
    rng = np.random.default_rng(300 + world)
    A, G, F, S = 4096, 64, 64, 500
    perm = rng.permutation(A)
    groups = [perm[g * F:(g + 1) * F] for g in range(G)]
    smap, shards, wb, bufs = _async_cluster(world, A, G, S, groups)
    oracle = CpuOracle(A, G)
    for g, m in enumerate(groups):
        oracle.create_group(g, m)
    all_agents = np.arange(A, dtype=np.uint32)
    for step in range(1, 7):
        par = step & 1
        for r in range(world):                            # every rank ingests its own slice, in rank order
            n = S if (step, r) != (3, world - 1) else 0     # an empty export once
            sender = rng.integers(0, A, n); grp = rng.integers(0, G, n)
            prio = rng.integers(0, 4, n); typ = rng.integers(0, 7, n)
            lens, off, buf = _payloads(rng, max(n, 1), 256)
            lens, off = lens[:n], off[:n]
            ts = rng.random(n)
            if step > 2:
                shards[r].wire_wait_done(bufs[par], wb, step - 2)
            shards[r].export_group_batch(sender, grp, prio, typ, lens, off, buf, bufs[par][r], wb, ts)
            shards[r].wire_publish(bufs[par][r], wb, step)
            if n:
                oracle.send_group_batch(sender, grp, prio, typ, lens, off, buf, ts)
        for s in shards:
            s.import_wire_ptrs_async(bufs[par], wb, step)
        k = [3, 1000, 2, 1000, 1000, 1000][step - 1]
        merged = {}
        for r, s in enumerate(shards):
            local = np.nonzero(smap == r)[0].astype(np.uint32)
            if step % 2 == 0 and world > 1:               # SDB_RECV_OWNED: the shard's own device-resident agent list
                from swarmdb_b200._native import RECV_OWNED
                res = s.receive_batch(None, k, RECV_OWNED)
                assert len(res[0]) == len(local)
            else:
                res = s.receive_batch(local, k)
            merged.update(_per_agent(*res, local))
        want = _per_agent(*oracle.receive_batch(all_agents, k, rec_cap=1 << 18), all_agents)
        for a in range(A):
            assert merged[a] == want[a], (step, a)
    assert len({s.stats()["next_seq"] for s in shards}) == 1 and shards[0].stats()["next_seq"] == oracle.next_seq
    for s in shards:
        st = s.stats()
        assert st["ring_overflow"] == 0 and st["enqueued"] == st["delivered"]
        s.close()


def test_async_import_mixed_traffic_and_direct_sends_in_between():
    """Point-to-point and broadcast sends inside asynchronously imported batches (their ring entries go through the
    commit sort, driven by the device-side batch record), interleaved with ordinary direct sends on a 1-shard handle:
    host and device counters hand over to each other in both directions."""
    from oracle.cpu_ref import CpuOracle
    from tests.fake_shard import OracleShard
    rng = np.random.default_rng(41)
    world, A, G, S = 2, 600, 6, 150
    perm = rng.permutation(A)
    groups = [perm[g * 64:(g + 1) * 64] for g in range(G)]
    smap, shards, wb, bufs = _async_cluster(world, A, G, S, groups, max_payload=128)
    ref = OracleShard(A, G, 0, 1)
    for g, m in enumerate(groups):
        ref.create_group(g, m)
    ref_bytes = ref.wire_bytes(S, S * 128 + 8 * 4 * 200 + 4096)
    ref_wire = np.zeros(world * ref_bytes, np.uint8)
    all_agents = np.arange(A, dtype=np.uint32)
    for step in range(1, 5):
        par = step & 1
        for r in range(world):
            n = S
            kind = rng.integers(0, 3, n).astype(np.uint8)
            sender = rng.integers(0, A, n)
            lists = [rng.choice(A, size=int(rng.integers(0, 100)), replace=False) for _ in range(4)]
            lo = np.zeros(5, np.uint64); lo[1:] = np.cumsum([len(x) for x in lists])
            li = np.concatenate(lists).astype(np.uint32)
            target = np.where(kind == 0, rng.integers(0, A, n), np.where(kind == 1, rng.integers(0, G, n), rng.integers(0, 4, n)))
            prio = rng.integers(0, 4, n); typ = rng.integers(0, 7, n)
            lens = rng.integers(0, 129, n).astype(np.uint16)
            off = np.arange(n, dtype=np.uint64) * 128
            buf = rng.integers(48, 123, n * 128 + 64).astype(np.uint8)
            ts = rng.random(n)
            if step > 2:
                shards[r].wire_wait_done(bufs[par], wb, step - 2)
            shards[r].export_mixed_batch(sender, kind, target, lo, li, prio, typ, lens, off, buf, bufs[par][r], wb, ts)
            shards[r].wire_publish(bufs[par][r], wb, step)
            ref.export_mixed_batch(sender, kind, target, lo, li, prio, typ, lens, off, buf,
                                   ref_wire[r * ref_bytes:(r + 1) * ref_bytes], ref_bytes, ts)
        for s in shards:
            s.import_wire_ptrs_async(bufs[par], wb, step)
        ref.import_wire_batches(world, ref_wire, ref_bytes)
        k = [2, 1000, 5, 1000][step - 1]
        for flags in ((1,) if step == 2 else (0,)):
            merged = {}
            for r, s in enumerate(shards):
                local = np.nonzero(smap == r)[0].astype(np.uint32)
                merged.update(_per_agent(*s.receive_batch(local, k, flags), local))
            want = _per_agent(*ref.o.receive_batch(all_agents, k, flags, rec_cap=1 << 18), all_agents)
            for a in range(A):
                assert merged[a] == want[a], (step, a)
    for s in shards:
        st = s.stats()
        assert st["ring_overflow"] == 0 and st["next_seq"] == ref.o.next_seq
        s.close()


@pytest.mark.parametrize("world", [1, 2])
def test_prefetched_import_equals_the_unpipelined_sequence(world):
    """sdb_import_prefetch: the flag wait + localize pass of step k+1 run on the handle's second stream while step k is
    still being received; the import then only places, fans out and indexes.  Mixed traffic (group, point-to-point,
    broadcast: both buffer sets carry temporary recipient lists), one step left unprefetched, partial drains."""
    from tests.fake_shard import OracleShard
    rng = np.random.default_rng(77 + world)
    A, G, S, T = 600, 6, 150, 7
    perm = rng.permutation(A)
    groups = [perm[g * 64:(g + 1) * 64] for g in range(G)]
    smap, shards, wb, bufs = _async_cluster(world, A, G, S, groups, max_payload=128)
    ref = OracleShard(A, G, 0, 1)
    for g, m in enumerate(groups):
        ref.create_group(g, m)
    ref_bytes = ref.wire_bytes(S, S * 128 + 8 * 4 * 200 + 4096)
    all_agents = np.arange(A, dtype=np.uint32)

    def make(step, r):
        g = np.random.default_rng(1000 * step + r)
        n = S if (step, r) != (4, world - 1) else 0
        kind = g.integers(0, 3, n).astype(np.uint8)
        sender = g.integers(0, A, n)
        lists = [g.choice(A, size=int(g.integers(0, 100)), replace=False) for _ in range(4)]
        lo = np.zeros(5, np.uint64); lo[1:] = np.cumsum([len(x) for x in lists])
        li = np.concatenate(lists).astype(np.uint32)
        target = np.where(kind == 0, g.integers(0, A, n), np.where(kind == 1, g.integers(0, G, n), g.integers(0, 4, n)))
        return (sender, kind, target, lo, li, g.integers(0, 4, n), g.integers(0, 7, n), g.integers(0, 129, n).astype(np.uint16),
                np.arange(n, dtype=np.uint64) * 128, g.integers(48, 123, max(n, 1) * 128 + 64).astype(np.uint8)), g.random(n)

    def export(step):
        par = step & 1
        for r in range(world):
            b, ts = make(step, r)
            if step > 2:
                shards[r].wire_wait_done(bufs[par], wb, step - 2)
            shards[r].export_mixed_batch(*b, bufs[par][r], wb, ts)
            shards[r].wire_publish(bufs[par][r], wb, step)

    export(1)
    for step in range(1, T + 1):
        for s in shards:
            s.import_wire_ptrs_async(bufs[step & 1], wb, step)
        if step < T:
            export(step + 1)
            if step != 3:                                   # step 4 is imported without a prefetch
                for s in shards:
                    s.import_prefetch(bufs[(step + 1) & 1], wb, step + 1)
                    s.import_prefetch(bufs[(step + 1) & 1], wb, step + 1)     # idempotent
        ref_wire = np.zeros(world * ref_bytes, np.uint8)
        for r in range(world):
            b, ts = make(step, r)
            ref.export_mixed_batch(*b, ref_wire[r * ref_bytes:(r + 1) * ref_bytes], ref_bytes, ts)
        ref.import_wire_batches(world, ref_wire, ref_bytes)
        k = [2, 1000, 5, 1000, 1, 3, 1000][step - 1]
        flags = 1 if step % 3 == 0 else 0
        merged = {}
        for r, s in enumerate(shards):
            local = np.nonzero(smap == r)[0].astype(np.uint32)
            merged.update(_per_agent(*s.receive_batch(local, k, flags), local))
        want = _per_agent(*ref.o.receive_batch(all_agents, k, flags, rec_cap=1 << 18), all_agents)
        for a in range(A):
            assert merged[a] == want[a], (step, a)
    for s in shards:
        st = s.stats()
        assert st["ring_overflow"] == 0 and st["next_seq"] == ref.o.next_seq
        s.close()


def test_async_import_that_does_not_fit_is_dropped_whole_and_reported():
    """Arena too small for an import: nothing of it is delivered, earlier traffic is intact, and the next
    host-synchronising call reports SDB_EARENA_FULL (never a silent drop)."""
    from swarmdb_b200._native import SdbError
    rng = np.random.default_rng(9)
    A, G, F, S = 1024, 16, 64, 200
    perm = rng.permutation(A)
    groups = [perm[g * F:(g + 1) * F] for g in range(G)]
    from swarmdb_b200._native import shared_payload_enabled
    # ~1.1 imports of this size: 200 sends x 64 recipients x 288 B = 3.7 MB, or x (64 headers + one payload) = 0.46 MB
    smap, shards, wb, bufs = _async_cluster(1, A, G, S, groups, arena_bytes=1 << (19 if shared_payload_enabled() else 22))
    s = shards[0]

    def export(step):
        sender = rng.integers(0, A, S); grp = rng.integers(0, G, S)
        lens = np.full(S, 256, np.uint16); off = np.arange(S, dtype=np.uint64) * 256
        buf = rng.integers(48, 123, S * 256 + 64).astype(np.uint8)
        s.export_group_batch(sender, grp, None, None, lens, off, buf, bufs[step & 1][0], wb)
        s.wire_publish(bufs[step & 1][0], wb, step)
    export(1)
    s.import_wire_ptrs_async(bufs[1], wb, 1)
    first = s.stats()["enqueued"]
    assert first > 0
    export(2)
    s.import_wire_ptrs_async(bufs[0], wb, 2)              # the first import is still unconsumed: this one cannot fit
    with pytest.raises(SdbError) as ei:
        s.stats()
    assert ei.value.code == -5
    st = s.stats()                                         # reported once; the queue keeps working
    assert st["enqueued"] == first
    cnt, hdr, _ = s.receive_batch(np.arange(A, dtype=np.uint32), 1000)
    assert len(hdr) == first
    export(3)
    s.import_wire_ptrs_async(bufs[1], wb, 3)              # room again after the drain
    assert s.stats()["enqueued"] > first
    s.close()
