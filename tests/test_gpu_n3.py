"""N3: inbox / load queries answered by the device queue (get_agent_load M:1049-1094,
get_unread_message_count M:1026-1047, get_stats M:973-1024) and the balancer fed by the queue backlog (M:1281-1325,
'get_agent_load is the only load signal').  Device results against oracle/cpu_ref.c and against the host-side
bookkeeping that tests/test_surface_cpu.py pins to the reference's goldens."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_agent_loads_and_queue_stats_match_oracle():
    from oracle.cpu_ref import CpuOracle
    from swarmdb_b200._native import RECV_PRIORITY, Shard
    rng = np.random.default_rng(12)
    A, G, F = 3000, 20, 48
    gpu, cpu = Shard(max_agents=A, max_groups=G, ring_slots=1024, arena_bytes=1 << 27), CpuOracle(A, G)
    for g in range(G):
        m = rng.choice(A, F, replace=False)
        gpu.create_group(g, m); cpu.create_group(g, m)
    idx = np.arange(A, dtype=np.uint32)
    gpu.register(idx); cpu.register(idx)
    for rnd in range(4):
        n = 600
        lens = rng.integers(0, 257, n).astype(np.uint16)
        off = np.arange(n, dtype=np.uint64) * 256
        buf = rng.integers(48, 123, n * 256 + 64).astype(np.uint8)
        s, g, r = rng.integers(0, A, n), rng.integers(0, G, n), rng.integers(0, A, n)
        prio, typ = rng.integers(0, 4, n), rng.integers(0, 7, n)
        gpu.send_group_batch(s, g, prio, typ, lens, off, buf); cpu.send_group_batch(s, g, prio, typ, lens, off, buf)
        gpu.send_batch(s, r, prio, typ, lens, off, buf); cpu.send_batch(s, r, prio, typ, lens, off, buf)
        some = rng.permutation(A)[: A // 3].astype(np.uint32)
        flags = RECV_PRIORITY if rnd % 2 else 0                   # priority receives leave holes in the windows
        gpu.receive_batch(some, 3, flags); cpu.receive_batch(some, 3, flags)
        lg, lc = gpu.agent_loads(), cpu.agent_loads(None, A)
        for f in ("received", "pending", "pending_by_prio", "pending_granules"):
            assert np.array_equal(lg[f], lc[f][: len(lg)]), (rnd, f)
        pick = rng.integers(0, A, 40).astype(np.uint32)
        assert gpu.agent_loads(pick).tobytes() == cpu.agent_loads(pick).tobytes()
        qg, qc = gpu.queue_stats(), cpu.queue_stats(gpu.stats_n_agents())
        assert qg == qc, (rnd, qg, qc)
    gpu.close(); cpu.close()


def test_balancer_fed_by_queue_backlog():
    from oracle.cpu_ref import CpuOracle
    from swarmdb_b200._native import Shard
    rng = np.random.default_rng(3)
    A, B = 2000, 16
    gpu, cpu = Shard(max_agents=A, ring_slots=256, arena_bytes=1 << 26, max_backends=B), CpuOracle(A, 1)
    idx = np.arange(A, dtype=np.uint32)
    gpu.register(idx); cpu.register(idx)
    w = rng.integers(1, 9, B).astype(np.uint32)
    gpu.set_backends(w); cpu.set_backends(w)
    agent_backend = rng.integers(0, B + 3, A)                      # some agents have no (valid) backend
    assigned = np.nonzero(agent_backend < B)[0].astype(np.uint32)
    gpu.assign_agent_backends(assigned, agent_backend[assigned])
    n = 5000
    lens = np.full(n, 64, np.uint16); off = np.arange(n, dtype=np.uint64) * 64
    buf = rng.integers(48, 123, n * 64 + 64).astype(np.uint8)
    s, r = rng.integers(0, A, n), rng.integers(0, A, n)
    gpu.send_batch(s, r, None, None, lens, off, buf); cpu.send_batch(s, r, None, None, lens, off, buf)
    gpu.receive_batch(idx[:500], 2); cpu.receive_batch(idx[:500], 2)
    gpu.backend_loads_from_queues()
    pend = cpu.agent_loads(None, A)["pending"].astype(np.uint64)
    want = np.zeros(B, np.uint64)
    np.add.at(want, agent_backend[assigned], pend[assigned])
    assert np.array_equal(gpu.backend_loads(), want)
    cpu.set_backends(w, want)
    assert np.array_equal(gpu.select_backends(3000, None, 0, 1), cpu.select_backends(3000, None, 0, 1))   # picks see the backlog
    gpu.close(); cpu.close()


def test_surface_queue_load_agrees_with_host_bookkeeping(tmp_path):
    """For point-to-point and group traffic the reference's inbox_size / unread_count (host dictionaries, pinned to
    the goldens by the bookkeeping scenario) and the device's answer coincide."""
    import swarmdb_b200 as sdb
    rng = np.random.default_rng(8)
    db = sdb.SwarmsDB(save_dir=str(tmp_path), auto_save=False,
                      gpu_config=sdb.GpuConfig(max_agents=256, max_groups=8, ring_slots=512, arena_bytes=1 << 24, deterministic_ids=True))
    names = [f"n{i:02d}" for i in range(30)]
    db.register_agents(names)
    db.add_agent_group("team", names[5:20])
    db.register_llm_backends(["b0", "b1", "b2"], [1, 2, 1])
    for a in names[:12]:
        db.assign_llm_backend(a, ["b0", "b1", "b2"][int(rng.integers(0, 3))])
    for k in range(200):
        s = names[int(rng.integers(0, 30))]
        if rng.random() < 0.3:
            db.send_to_group(s, "team", f"g{k}", priority=sdb.MessagePriority(int(rng.integers(0, 4))))
        else:
            db.send_message(s, f"m{k}", names[int(rng.integers(0, 30))], priority=sdb.MessagePriority(int(rng.integers(0, 4))))
        if k % 17 == 0:
            db.receive_messages(names[int(rng.integers(0, 30))], int(rng.integers(1, 6)))
    for a in names:
        host, dev = db.get_agent_load(a), db.get_agent_queue_load(a)
        assert dev["inbox_size"] == host["inbox_size"] and dev["unread_count"] == host["unread_count"], a
        assert sum(dev["unread_by_priority"]) == dev["unread_count"]
    st = db.get_stats()
    assert st["queue"]["pending"] == sum(db.get_unread_message_count(a) for a in names)
    loads = db.refresh_llm_backend_loads()
    want = {"b0": 0, "b1": 0, "b2": 0}
    for a in names[:12]:
        want[db.get_llm_backend(a)] += db.get_unread_message_count(a)
    assert loads == want
    db.close()
