#!/usr/bin/env python
"""Secondary measurements for the other BASELINE configs (one JSON line each):

  c1  2 agents, 1k point-to-point 128-byte messages through the SwarmsDB Python surface
  c4  priority dequeue: 10k agents x 10k pending 256-byte messages, 4 levels, receive_batch(all, 100)
  c5  backend balancer: 1M requests over 256 backends (weighted least-load, weighted random)
  k1  point-to-point enqueue kernel roofline (1M records x 256 B)

bench.py remains the contract benchmark (c2/c3); this file only feeds DESIGN.md / profiles/.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from bench import ALNUM, hbm_peak  # noqa: E402


def c1(tmp="/tmp/sdb_c1"):
    import swarmdb_b200 as sdb
    rng = np.random.default_rng(1)
    contents = [ALNUM[rng.integers(0, 62, 128)].tobytes().decode() for _ in range(1000)]
    db = sdb.SwarmsDB(save_dir=tmp, auto_save=False, gpu_config=sdb.GpuConfig(max_agents=1024, ring_slots=2048, arena_bytes=1 << 26))
    db.register_agent("agent_a"); db.register_agent("agent_b")
    for c in contents[:50]:
        db.send_message("agent_a", c, "agent_b")
    db.receive_messages("agent_b", 100)
    t0 = time.perf_counter()
    for c in contents:
        db.send_message("agent_a", c, "agent_b")
    t1 = time.perf_counter()
    got = db.receive_messages("agent_b", 2000)
    t2 = time.perf_counter()
    assert [m.content for m in got] == contents
    db.close()
    return {"config": "c1: 2 agents, 1k p2p 128-byte messages, SwarmsDB Python surface (buffered sends, one flush at receive)",
            "send_msgs_per_s": 1000 / (t1 - t0), "receive_msgs_per_s": 1000 / (t2 - t1),
            "send_to_receive_msgs_per_s": 1000 / (t2 - t0),
            "reference_python_probe": "send 9.2k/s, receive 46k/s, end-to-end 7.6k/s (BASELINE.md, in-process Kafka stub)"}


def c4(n_agents=10_000, depth=10_000, reps=4):
    from swarmdb_b200._native import RECV_PRIORITY, Shard
    peak, src = hbm_peak()
    per_batch = 100                                   # messages per agent per preload batch
    n = n_agents * per_batch
    shard = Shard(max_agents=n_agents, ring_slots=16384, arena_bytes=1 << 35, max_payload_bytes=256, max_batch_sends=n,
                  max_batch_payload=n * 256, max_recv_records=n_agents * 100, max_recv_payload=n_agents * 100 * 256)
    rng = np.random.default_rng(4)
    recv = np.tile(np.arange(n_agents, dtype=np.uint32), per_batch)
    send = rng.integers(0, n_agents, n).astype(np.uint32)
    lens = np.full(n, 256, np.uint16)
    off = np.arange(n, dtype=np.uint64) * 256
    payload = ALNUM[rng.integers(0, 62, n * 256)]
    staged = [shard.stage(0, send, recv, rng.integers(0, 4, n).astype(np.uint8), None, lens, off, payload) for _ in range(2)]
    t0 = time.perf_counter()
    for b in range(depth // per_batch):
        shard.submit(staged[b % 2])
    shard.sync()
    load_s = time.perf_counter() - t0
    st = shard.stats()
    assert st["enqueued"] == n_agents * depth and st["ring_overflow"] == 0, st
    shard.profile(True)
    sweeps = []
    for r in range(reps):
        t1 = time.perf_counter()
        _, total, _ = shard.receive_batch(None, 100, RECV_PRIORITY, copy_out=False)
        sweeps.append((time.perf_counter() - t1) * 1e3)
        assert total == n_agents * 100
    prof = shard.profile_read()
    shard.close()
    sel = prof["recv_select"][0] / prof["recv_select"][1]
    gat = prof["recv_gather"][0] / prof["recv_gather"][1]
    k5_ms = sum(v[0] for k, v in prof.items() if k.startswith("recv")) / reps
    alg = n_agents * (depth * 1 + 2 * 100 * (256 + 16))          # 64,400 B per agent-call
    return {"config": f"c4: {n_agents} agents x {depth} pending 256-B messages, 4 priorities, receive_batch(all agents, 100, PRIORITY)",
            "preload_s": load_s, "preload_msgs_per_s": n_agents * depth / load_s,
            "sweep_wall_ms": sweeps, "p50_sweep_wall_ms": float(np.median(sweeps)),
            "k5_kernels_ms_per_sweep": k5_ms, "select_ms": sel, "gather_ms": gat,
            "roofline": {"bound": "hbm", "algorithmic_bytes_per_sweep": alg, "achieved": alg / (k5_ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": alg / (k5_ms * 1e-3) / 1e9 / peak, "peak_source": src},
            "per_agent_call_us": k5_ms * 1e3 / n_agents}


def c5(n_req=1_000_000, n_backends=256):
    from oracle.cpu_ref import CpuOracle
    from swarmdb_b200._native import Shard
    rng = np.random.default_rng(5)
    w = rng.integers(1, 9, n_backends)
    out = {"config": f"c5: {n_req} unit-cost requests over {n_backends} backends, weights U{{1..8}}, load0 = 0"}
    for mode, name in ((0, "weighted_least_load"), (1, "weighted_random")):
        g = Shard(max_agents=16)
        g.set_backends(w)
        g.select_backends(1000, None, mode, 6)                   # warm-up
        g.set_backends(w)
        g.profile(True)
        t0 = time.perf_counter()
        picks = g.select_backends(n_req, None, mode, 6)
        wall = time.perf_counter() - t0
        prof = g.profile_read()
        o = CpuOracle(16)
        o.set_backends(w)
        t1 = time.perf_counter()
        ref = o.select_backends(n_req, None, mode, 6)
        cpu = time.perf_counter() - t1
        assert np.array_equal(picks, ref) and np.array_equal(g.backend_loads(), o.backend_loads())
        out[name] = {"picks_per_s_end_to_end": n_req / wall, "wall_ms": wall * 1e3,
                     "cpu_oracle_picks_per_s_1_core": n_req / cpu, "bit_exact_vs_oracle": True}
        g.close(); o.close()
    out["note"] = "not HBM-bound: the 256-entry table lives in shared memory / L1 (SURVEY 8d says report picks/s)"
    return out


def k1(n=1_000_000):
    from swarmdb_b200._native import Shard
    peak, src = hbm_peak()
    A = 1_000_000
    shard = Shard(max_agents=A, ring_slots=16, arena_bytes=1 << 32, max_payload_bytes=256, max_batch_sends=n,
                  max_batch_payload=n * 256, max_recv_records=n + 4096, max_recv_payload=(n + 4096) * 256)
    rng = np.random.default_rng(8)
    lens = np.full(n, 256, np.uint16)
    off = np.arange(n, dtype=np.uint64) * 256
    payload = ALNUM[rng.integers(0, 62, n * 256)]
    st = shard.stage(0, rng.integers(0, A, n).astype(np.uint32), rng.permutation(A)[:n].astype(np.uint32), None, None, lens, off, payload)
    shard.register(np.arange(A, dtype=np.uint32))
    for _ in range(3):
        shard.submit(st); shard.receive_batch(None, 100, 0, copy_out=False)
    shard.profile(True)
    for _ in range(8):
        shard.submit(st); shard.receive_batch(None, 100, 0, copy_out=False)
    prof = shard.profile_read()
    shard.close()
    ms = prof["p2p"][0] / prof["p2p"][1]
    alg = 2 * (256 + 16) * n
    return {"config": f"k1: {n} point-to-point 256-B records to distinct agents per batch",
            "p2p_ms": ms, "commit_ms": prof["commit"][0] / prof["commit"][1],
            "roofline": {"bound": "hbm", "algorithmic_bytes": alg, "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak, "peak_source": src}}


def bcast(n_agents=1_000_000, n_sends=16):
    """R11 at scale: broadcasts to every registered agent (one record image per recipient)."""
    from swarmdb_b200._native import Shard
    peak, src = hbm_peak()
    shard = Shard(max_agents=n_agents, ring_slots=64, arena_bytes=1 << 33, max_payload_bytes=256, max_batch_sends=n_sends,
                  max_batch_payload=n_sends * 256 + 64, list_pool_entries=n_sends * n_agents + 1024,
                  max_recv_records=n_sends * n_agents + 4096, max_recv_payload=(n_sends * n_agents + 4096) * 256)
    rng = np.random.default_rng(11)
    shard.register(np.arange(n_agents, dtype=np.uint32))
    lo = np.arange(n_sends + 1, dtype=np.uint64) * n_agents
    li = np.tile(np.arange(n_agents, dtype=np.uint32), n_sends)
    lens = np.full(n_sends, 256, np.uint16)
    off = np.arange(n_sends, dtype=np.uint64) * 256
    payload = ALNUM[rng.integers(0, 62, n_sends * 256 + 64)]
    sender = rng.integers(0, n_agents, n_sends).astype(np.uint32)
    out = []
    shard.profile(True)
    for rep in range(3):
        shard.sync()
        t0 = time.perf_counter()
        shard.send_list_batch(sender, lo, li, None, None, lens, off, payload)
        shard.sync()
        t1 = time.perf_counter()
        _, total, _ = shard.receive_batch(None, 100, 0, copy_out=False)
        shard.sync()
        t2 = time.perf_counter()
        assert total == n_sends * n_agents
        out.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3))
    prof = shard.profile_read()
    shard.close()
    kern = {k: v[0] / v[1] for k, v in prof.items() if v[1]}
    send_ms = kern.get("fanout", min(o[0] for o in out))
    alg = n_sends * n_agents * (256 + 16 + 4)
    return {"config": f"bcast: {n_sends} broadcasts x {n_agents} recipients, 256-B payloads (send_list_batch incl. the H2D of the recipient lists), then full drain",
            "send_wall_ms": [o[0] for o in out], "drain_wall_ms": [o[1] for o in out], "kernel_ms_per_launch": kern,
            "roofline": {"bound": "hbm", "algorithmic_bytes": alg, "achieved": alg / (send_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (send_ms * 1e-3) / 1e9 / peak, "peak_source": src,
                         "note": "fan-out kernel only; the call's wall time adds host validation and 64 MB of recipient indices over PCIe"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["c1", "c4", "c5", "k1", "bcast"])
    args = ap.parse_args()
    for w in args.which:
        print(json.dumps({"bench": w, **globals()[w]()}), flush=True)


if __name__ == "__main__":
    main()
