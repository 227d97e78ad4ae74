import sys, os, json, numpy as np
sys.path.insert(0, '/root/repo')
import torch
from bench import Workload
from swarmdb_b200._native import Shard
wl = Workload()
shard = Shard(max_agents=wl.A, ring_slots=64, arena_bytes=1<<33, max_payload_bytes=256, max_groups=1<<14, member_pool_entries=wl.A+1024, max_batch_sends=wl.S, max_batch_payload=wl.S*256, max_recv_records=wl.S*64+65536, max_recv_payload=(wl.S*64+65536)*256, fanout_variant=2)
shard.register(np.arange(wl.A, dtype=np.uint32))
for g in range(wl.G): shard.create_group(g, wl.members(g))
st = [shard.stage(1, *wl.batch()) for _ in range(4)]
shard.profile(True)
for i in range(12):
    shard.submit(st[i % 4])
    shard.receive_batch(None, 100, 0, copy_out=False)
p = shard.profile_read()
print(os.environ.get("SDB_DEBUG_FANOUT"), {k: round(v[0]/v[1], 4) for k, v in p.items() if v[1]})
