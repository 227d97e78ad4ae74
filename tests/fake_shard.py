"""Oracle-backed stand-in for `swarmdb_b200._native.Shard` (TESTS ONLY - the product has no CPU path).
It lets the host-side multi-GPU logic (swarmdb_b200/sharded.py) run under gloo on CPU: wire batches are
numpy blobs, `import` replays them into oracle/cpu_ref.c with the shard's ownership filter."""
import numpy as np

from oracle.cpu_ref import CpuOracle

MAGIC = 0x57424453


class OracleShard:
    def __init__(self, max_agents, max_groups, rank, world):
        self.o, self.rank, self.world = CpuOracle(max_agents, max_groups), rank, world
        self.max_agents = max_agents
        self.smap = np.full(max_agents, rank, np.uint8)
        self.full, self.groups = {}, {}

    # ---- Shard surface used by SwarmsDB / ShardedSwarmsDB / ShardExchange
    def register(self, idx): self.o.register(idx)
    def deregister(self, idx): pass
    def sync(self): pass
    def close(self): self.o.close()
    def agent_loads(self, agents=None, n=None): return self.o.agent_loads(agents, n)
    def queue_stats(self): return self.o.queue_stats()

    def assign_agent_backends(self, agents, backends):
        self._ab = getattr(self, "_ab", {})
        self._ab.update({int(a): int(b) for a, b in zip(agents, backends)})

    def backend_loads_from_queues(self):
        ld = self.o.agent_loads(None, self.max_agents)["pending"]
        loads = np.zeros(self.o._nb, np.uint64)
        for a, b in getattr(self, "_ab", {}).items():
            if b < self.o._nb:
                loads[b] += int(ld[a])
        self.o.set_backends(self._w, loads)

    def set_backends(self, weight, load0=None):
        self._w = np.asarray(weight, np.uint32)
        self.o.set_backends(self._w, load0)

    def backend_loads(self): return self.o.backend_loads()
    def select_backends(self, n_req, cost=None, mode=0, seed=0): return self.o.select_backends(n_req, cost, mode, seed)
    def release_backends(self, backend, cost=None):
        l = self.o.backend_loads()
        for i, b in enumerate(backend):
            c = 1 if cost is None else int(cost[i])
            l[b] = l[b] - c if l[b] >= c else 0
        self.o.set_backends(self._w, l)

    def advance_seq(self, next_seq): self.o.next_seq = max(self.o.next_seq, next_seq)

    def stats(self): return {"next_seq": self.o.next_seq, "ring_overflow": 0, "n_agents": self.max_agents}
    def overflow_log(self, cap=4096): return np.zeros(0, np.uint32), np.zeros(0, np.uint64), 0

    def set_agent_shards(self, shard_of):
        s = np.asarray(shard_of, np.uint8)
        self.smap[:len(s)] = s
        for g, m in self.groups.items():
            self._install(g, m)

    def _install(self, g, members):
        keep = np.nonzero(self.smap[members] == self.rank)[0]
        self.o.create_group_pos(g, members[keep], keep.astype(np.uint32))

    def create_group(self, g, members):
        members = np.asarray(members, np.uint32)
        self.groups[g] = members; self.full[g] = len(members)
        self._install(g, members)

    def wire_bytes(self, max_sends, max_payload):
        return 4096 + max_sends * 128 + max_payload

    def export_mixed_batch(self, sender, kind, target, list_off, list_idx, prio, typ, lens, payload_off, payload, wire, cap,
                           ts=None, seq_base=0):
        n = len(sender)
        kind = np.asarray(kind, np.uint8); target = np.asarray(target, np.uint32)
        sizes = np.array([self.full[int(t)] if k == 1 else 1 for k, t in zip(kind, target)], np.uint64) if n else np.zeros(0, np.uint64)
        rec0 = np.zeros(n, np.uint64)
        if n > 1:
            rec0[1:] = np.cumsum(sizes)[:-1]
        lo = np.asarray(list_off if list_off is not None else [0], np.uint64)
        li = np.asarray(list_idx if list_idx is not None else [], np.uint32)
        payload = np.asarray(payload, np.uint8)
        hdr = np.array([MAGIC, n, int(sizes.sum()) if n else 0, len(payload), int(seq_base), len(lo), len(li)], np.int64)
        ts = np.zeros(n, np.float64) if ts is None else np.asarray(ts, np.float64)
        parts = [hdr.view(np.uint8), ts.view(np.uint8), np.asarray(sender, np.uint32).view(np.uint8), kind, target.view(np.uint8),
                 np.asarray(prio, np.uint8), np.asarray(typ, np.uint8), np.asarray(lens, np.uint16).view(np.uint8),
                 np.asarray(payload_off, np.uint64).view(np.uint8), rec0.view(np.uint8), lo.view(np.uint8), li.view(np.uint8), payload]
        blob = np.concatenate(parts)
        assert len(blob) <= cap, (len(blob), cap)
        (wire if isinstance(wire, np.ndarray) else wire.numpy())[:len(blob)] = blob

    def send_mixed_batch(self, sender, kind, target, list_off, list_idx, prio, typ, lens, payload_off, payload, ts=None):
        cap = self.wire_bytes(len(sender), len(payload)) + 4 * len(list_idx if list_idx is not None else []) + \
            8 * len(list_off if list_off is not None else [0])
        buf = np.zeros(cap, np.uint8)
        self.export_mixed_batch(sender, kind, target, list_off, list_idx, prio, typ, lens, payload_off, payload, buf, cap, ts)
        return self.import_wire_batches(1, buf, cap)

    def export_group_batch(self, sender, group, prio, typ, lens, payload_off, payload, wire, cap, ts=None):
        n = len(sender)
        self.export_mixed_batch(sender, np.ones(n, np.uint8), group, None, None, prio, typ, lens, payload_off, payload, wire, cap, ts)

    def import_wire_batches(self, n_src, recv, stride):
        base = self.o.next_seq
        run = 0
        end = 0
        raw = recv if isinstance(recv, np.ndarray) else recv.numpy()
        for s in range(n_src):
            b = raw[s * stride:(s + 1) * stride]
            magic, n, total, pbytes, seq_base, nlo, nli = (int(x) for x in b[:56].view(np.int64))
            assert magic == MAGIC
            o = [56]
            def take(dt, count):
                nb = np.dtype(dt).itemsize * count
                out = b[o[0]:o[0] + nb].view(dt).copy(); o[0] += nb
                return out
            ts = take(np.float64, n)
            sender, kind, target = take(np.uint32, n), take(np.uint8, n), take(np.uint32, n)
            prio, typ, lens = take(np.uint8, n), take(np.uint8, n), take(np.uint16, n)
            poff, rec0, lo, li, payload = take(np.uint64, n), take(np.uint64, n), take(np.uint64, nlo), take(np.uint32, nli), take(np.uint8, pbytes)
            src_base = seq_base if seq_base else base + run
            for i in range(n):
                one = slice(i, i + 1)
                self.o.next_seq = src_base + int(rec0[i])
                if kind[i] == 1:
                    self.o.send_group_seq(sender[one], target[one], prio[one], typ[one], lens[one], poff[one], payload,
                                          np.array([src_base + int(rec0[i])], np.uint64), ts[one])
                elif kind[i] == 0:
                    if self.smap[target[i]] == self.rank:
                        self.o.send_batch(sender[one], target[one], prio[one], typ[one], lens[one], poff[one], payload, ts[one])
                else:
                    t = int(target[i])
                    rec = li[int(lo[t]):int(lo[t + 1])]
                    rec = rec[self.smap[rec] == self.rank]
                    self.o.send_list_batch(sender[one], [0, len(rec)], rec, prio[one], typ[one], lens[one], poff[one], payload, ts[one])
            run += total
            end = max(end, src_base + total)
        self.o.next_seq = max(base + run, end)
        return base

    def receive_one(self, agent, max_messages=100, flags=0):
        cnt, hdr, pay = self.o.receive_batch([agent], max_messages, flags, rec_cap=max(max_messages, 1), pay_cap=1 << 22)
        return hdr, pay, max_messages

    def receive_batch(self, agents, max_messages, flags=0, **kw):
        return self.o.receive_batch(agents, max_messages, flags)
