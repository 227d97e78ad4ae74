"""Single-GPU proxy of one shard's work at N=8: import 8 local wire batches (65,536 group sends each)
into a shard that owns 1/8 of the agents, then drain.  Prints per-kernel-class ms (CUDA events)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import Workload
from swarmdb_b200._native import Shard
from swarmdb_b200.sharded import shard_map_numbered

W = int(os.environ.get("W", "8"))
wl = Workload()
shard = Shard(max_agents=wl.A, ring_slots=256, arena_bytes=1 << 33, max_payload_bytes=256, max_groups=1 << 14,
              member_pool_entries=wl.A + 1024, max_batch_sends=wl.S, max_batch_payload=wl.S * 256,
              max_recv_records=wl.S * 64 * 2, max_recv_payload=wl.S * 64 * 2 * 256, shard_id=0, num_shards=W)
smap = shard_map_numbered("agent_", 7, wl.A, W)
shard.set_agent_shards(smap)
for g in range(wl.G):
    shard.create_group(g, wl.members(g))
wb = shard.wire_bytes(wl.S, wl.S * 256)
bufs = []
for r in range(W):
    p, _ = shard.wire_alloc(wb)
    shard.export_group_batch(*wl.batch(), p, wb)
    bufs.append(p)
shard.sync()
shard.profile(True)
tot = 0
for i in range(10):
    shard.import_wire_ptrs(bufs)
    _, t, _ = shard.receive_batch(None, 100, 0, copy_out=False)
    tot += t
p = shard.profile_read()
print("W=%d delivered/step=%d" % (W, tot // 10), {k: round(v[0] / v[1], 4) for k, v in p.items() if v[1]})
