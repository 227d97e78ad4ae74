"""ctypes binding of the C ABI in include/swarmdb_b200.h (the drop-in boundary).

`Shard` is one GPU-resident queue shard (one handle, one GPU).  There is no CPU
implementation behind it: if the CUDA library is missing or no device is present, loading /
construction fails loudly - it never falls back.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional, Tuple

import numpy as np

_CSRC = Path(__file__).resolve().parent / "csrc"
LIB_PATH = _CSRC / "libswarmdb_b200.so"

LOAD_DTYPE = np.dtype([("received", "<u4"), ("pending", "<u4"), ("pending_by_prio", "<u4", (4,)), ("pending_granules", "<u4"),
                       ("reserved", "<u4")])
QSTATS_FIELDS = ["agents_with_pending", "pending", "p0", "p1", "p2", "p3", "pending_granules", "received", "max_pending",
                 "max_pending_agent"]
HDR_DTYPE = np.dtype([("seq", "<u8"), ("timestamp", "<f8"), ("sender", "<u4"), ("receiver", "<u4"),
                      ("group", "<u4"), ("len", "<u2"), ("prio", "u1"), ("type", "u1")])
assert HDR_DTYPE.itemsize == 32

NO_GROUP = 0xFFFFFFFF
NO_RECEIVER = 0xFFFFFFFF
RECV_PRIORITY = 1
RECV_PEEK = 2
RECV_ASYNC = 4
RECV_OWNED = 8
TYPEF_JSON = 0x08
TYPEF_EXTRAS = 0x10
TYPE_MASK = 0x07

STATUS_NAMES = {0: "SDB_OK", -1: "SDB_EINVAL", -2: "SDB_ECUDA", -3: "SDB_ENOMEM", -4: "SDB_ERING_OVERFLOW",
                -5: "SDB_EARENA_FULL", -6: "SDB_ECAPACITY", -7: "SDB_ENOTFOUND", -8: "SDB_EOUTPUT"}


def shared_payload_enabled() -> bool:
    """Mirror of the library's SDB_SHARED_PAYLOAD switch (csrc/sdb_api.cu, sdb_create): group sends above the pull
    threshold keep ONE payload per send in the arena log instead of one per recipient."""
    return os.environ.get("SDB_SHARED_PAYLOAD", "1") != "0"


class SdbError(RuntimeError):
    def __init__(self, code: int, msg: str) -> None:
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {msg}")
        self.code = code


class SdbConfig(C.Structure):
    _fields_ = [("struct_bytes", C.c_uint32), ("device", C.c_int32), ("shard_id", C.c_uint32),
                ("num_shards", C.c_uint32), ("max_agents", C.c_uint32), ("ring_slots", C.c_uint32),
                ("arena_bytes", C.c_uint64), ("max_payload_bytes", C.c_uint32), ("max_groups", C.c_uint32),
                ("member_pool_entries", C.c_uint64), ("max_backends", C.c_uint32), ("max_batch_sends", C.c_uint32),
                ("max_batch_payload", C.c_uint64), ("max_recv_records", C.c_uint64), ("max_recv_payload", C.c_uint64),
                ("list_pool_entries", C.c_uint64), ("fanout_variant", C.c_uint32), ("flags", C.c_uint32)]


class SdbStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("next_seq", "enqueued", "delivered", "ring_overflow", "skipped_sender",
                                          "arena_tail_bytes", "arena_floor_bytes", "n_agents", "kernel_launches",
                                          "backend_picks")]


EXPORTS = ["sdb_abi_version", "sdb_create", "sdb_destroy", "sdb_set_stream", "sdb_sync", "sdb_last_error",
           "sdb_get_stats", "sdb_debug_set_arena_pos", "sdb_advance_seq", "sdb_profile", "sdb_profile_read", "sdb_register_agents", "sdb_deregister_agents", "sdb_create_group", "sdb_send_batch",
           "sdb_send_group_batch", "sdb_send_list_batch", "sdb_send_mixed_batch", "sdb_stage_batch", "sdb_submit_staged", "sdb_free_staged",
           "sdb_receive_batch", "sdb_latency_server", "sdb_last_receive_dev", "sdb_last_receive_totals",
           "sdb_digest_reset", "sdb_digest_fold", "sdb_digest_read", "sdb_wire_bytes", "sdb_set_agent_shards",
           "sdb_export_group_batch", "sdb_export_mixed_batch", "sdb_export_mixed_batch_seq", "sdb_import_wire_batches", "sdb_wire_alloc", "sdb_wire_open",
           "sdb_wire_close", "sdb_import_wire_ptrs", "sdb_wire_wait_done", "sdb_wire_publish", "sdb_import_wire_ptrs_async", "sdb_import_prefetch", "sdb_overflow_log", "sdb_set_backends", "sdb_get_backend_loads",
           "sdb_release_backends", "sdb_select_backend_batch", "sdb_agent_loads", "sdb_queue_stats",
           "sdb_assign_agent_backends", "sdb_backend_loads_from_queues"]

_lib = None


def load_library() -> C.CDLL:
    """dlopen the in-tree CUDA library.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("SWARMDB_B200_LIB", str(LIB_PATH)))
    if not path.exists():
        raise ImportError(f"{path} not built - run `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"(or make -C swarmdb_b200/csrc). swarmdb_b200 has no CPU fallback.")
    L = C.CDLL(str(path))
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    L.sdb_abi_version.restype = i32
    L.sdb_create.restype = i32; L.sdb_create.argtypes = [C.POINTER(SdbConfig), C.POINTER(vp)]
    L.sdb_destroy.restype = i32; L.sdb_destroy.argtypes = [vp]
    L.sdb_set_stream.restype = i32; L.sdb_set_stream.argtypes = [vp, vp]
    L.sdb_sync.restype = i32; L.sdb_sync.argtypes = [vp]
    L.sdb_last_error.restype = C.c_char_p; L.sdb_last_error.argtypes = [vp]
    L.sdb_get_stats.restype = i32; L.sdb_get_stats.argtypes = [vp, C.POINTER(SdbStats)]
    L.sdb_debug_set_arena_pos.restype = i32; L.sdb_debug_set_arena_pos.argtypes = [vp, u64]
    L.sdb_advance_seq.restype = i32; L.sdb_advance_seq.argtypes = [vp, u64]
    L.sdb_profile.restype = i32; L.sdb_profile.argtypes = [vp, i32]
    L.sdb_profile_read.restype = i32; L.sdb_profile_read.argtypes = [vp, vp, vp]
    L.sdb_register_agents.restype = i32; L.sdb_register_agents.argtypes = [vp, u32, vp]
    L.sdb_deregister_agents.restype = i32; L.sdb_deregister_agents.argtypes = [vp, u32, vp]
    L.sdb_create_group.restype = i32; L.sdb_create_group.argtypes = [vp, u32, u32, vp]
    L.sdb_send_batch.restype = i32; L.sdb_send_batch.argtypes = [vp, u32] + [vp] * 7 + [u64, vp, vp]
    L.sdb_send_group_batch.restype = i32; L.sdb_send_group_batch.argtypes = [vp, u32] + [vp] * 7 + [u64, vp, vp]
    L.sdb_send_list_batch.restype = i32; L.sdb_send_list_batch.argtypes = [vp, u32] + [vp] * 8 + [u64, vp, vp]
    L.sdb_send_mixed_batch.restype = i32
    L.sdb_send_mixed_batch.argtypes = [vp, u32, vp, vp, vp, u32] + [vp] * 7 + [u64, vp, vp]
    L.sdb_stage_batch.restype = i32; L.sdb_stage_batch.argtypes = [vp, u32, u32] + [vp] * 7 + [u64, vp, C.POINTER(vp)]
    L.sdb_submit_staged.restype = i32; L.sdb_submit_staged.argtypes = [vp, vp, vp]
    L.sdb_free_staged.restype = i32; L.sdb_free_staged.argtypes = [vp, vp]
    L.sdb_receive_batch.restype = i32
    L.sdb_receive_batch.argtypes = [vp, u32, vp, u32, u32, vp, vp, u64, vp, u64, vp, vp]
    L.sdb_latency_server.restype = i32; L.sdb_latency_server.argtypes = [vp, i32]
    L.sdb_last_receive_dev.restype = i32; L.sdb_last_receive_dev.argtypes = [vp, vp, vp, vp]
    L.sdb_last_receive_totals.restype = i32; L.sdb_last_receive_totals.argtypes = [vp, vp, vp]
    L.sdb_digest_reset.restype = i32; L.sdb_digest_reset.argtypes = [vp]
    L.sdb_digest_fold.restype = i32; L.sdb_digest_fold.argtypes = [vp]
    L.sdb_digest_read.restype = i32; L.sdb_digest_read.argtypes = [vp, u32, vp, vp]
    L.sdb_wire_bytes.restype = u64; L.sdb_wire_bytes.argtypes = [vp, u32, u64]
    L.sdb_set_agent_shards.restype = i32; L.sdb_set_agent_shards.argtypes = [vp, u32, vp]
    L.sdb_export_group_batch.restype = i32; L.sdb_export_group_batch.argtypes = [vp, u32] + [vp] * 7 + [u64, vp, vp, u64]
    L.sdb_export_mixed_batch.restype = i32
    L.sdb_export_mixed_batch.argtypes = [vp, u32, vp, vp, vp, u32] + [vp] * 7 + [u64, vp, vp, u64]
    L.sdb_export_mixed_batch_seq.restype = i32
    L.sdb_export_mixed_batch_seq.argtypes = [vp, u64, u32, vp, vp, vp, u32] + [vp] * 7 + [u64, vp, vp, u64]
    L.sdb_wire_alloc.restype = i32; L.sdb_wire_alloc.argtypes = [vp, u64, C.POINTER(vp), vp]
    L.sdb_wire_open.restype = i32; L.sdb_wire_open.argtypes = [vp, vp, C.POINTER(vp)]
    L.sdb_wire_close.restype = i32; L.sdb_wire_close.argtypes = [vp, vp, i32]
    L.sdb_import_wire_ptrs.restype = i32; L.sdb_import_wire_ptrs.argtypes = [vp, u32, vp, vp]
    L.sdb_wire_wait_done.restype = i32; L.sdb_wire_wait_done.argtypes = [vp, u32, vp, u64, u32]
    L.sdb_wire_publish.restype = i32; L.sdb_wire_publish.argtypes = [vp, vp, u64, u32]
    L.sdb_import_wire_ptrs_async.restype = i32; L.sdb_import_wire_ptrs_async.argtypes = [vp, u32, vp, u64, u32]
    L.sdb_import_prefetch.restype = i32; L.sdb_import_prefetch.argtypes = [vp, u32, vp, u64, u32]
    L.sdb_overflow_log.restype = i32; L.sdb_overflow_log.argtypes = [vp, u32, vp, vp, vp, vp]
    L.sdb_import_wire_batches.restype = i32; L.sdb_import_wire_batches.argtypes = [vp, u32, vp, u64, vp]
    L.sdb_set_backends.restype = i32; L.sdb_set_backends.argtypes = [vp, u32, vp, vp]
    L.sdb_get_backend_loads.restype = i32; L.sdb_get_backend_loads.argtypes = [vp, u32, vp]
    L.sdb_release_backends.restype = i32; L.sdb_release_backends.argtypes = [vp, u32, vp, vp]
    L.sdb_select_backend_batch.restype = i32; L.sdb_select_backend_batch.argtypes = [vp, u32, vp, u32, u64, vp]
    L.sdb_agent_loads.restype = i32; L.sdb_agent_loads.argtypes = [vp, u32, vp, vp]
    L.sdb_queue_stats.restype = i32; L.sdb_queue_stats.argtypes = [vp, vp]
    L.sdb_assign_agent_backends.restype = i32; L.sdb_assign_agent_backends.argtypes = [vp, u32, vp, vp]
    L.sdb_backend_loads_from_queues.restype = i32; L.sdb_backend_loads_from_queues.argtypes = [vp]
    if L.sdb_abi_version() != 1:
        raise ImportError("swarmdb_b200 ABI version mismatch")
    _lib = L
    return L


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    raise TypeError(type(a))


def _arr(x, dt) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=dt)


def pad32(n):
    return (n + 31) & ~31


# kernel variant used when the caller does not choose one (SDB_FANOUT_VARIANT overrides, for A/B runs)
DEFAULT_FANOUT_VARIANT = int(os.environ.get("SDB_FANOUT_VARIANT", "2"))


class Shard:
    """One GPU-resident shard: per-agent rings + message arena + backend table on one device."""

    def __init__(self, max_agents: int, ring_slots: int = 64, arena_bytes: int = 1 << 28,
                 max_payload_bytes: int = 256, max_groups: int = 1, member_pool_entries: int = 0,
                 max_backends: int = 256, max_batch_sends: int = 65536, max_batch_payload: int = 0,
                 max_recv_records: int = 1 << 20, max_recv_payload: int = 0, list_pool_entries: int = 0,
                 device: int = 0, shard_id: int = 0, num_shards: int = 1, fanout_variant: Optional[int] = None) -> None:
        self._L = load_library()
        cfg = SdbConfig()
        cfg.struct_bytes = C.sizeof(SdbConfig)
        cfg.device, cfg.shard_id, cfg.num_shards = device, shard_id, num_shards
        cfg.max_agents, cfg.ring_slots, cfg.arena_bytes = max_agents, ring_slots, arena_bytes
        cfg.max_payload_bytes, cfg.max_groups = max_payload_bytes, max_groups
        cfg.member_pool_entries = member_pool_entries or max(1024, 2 * max_agents)
        cfg.max_backends, cfg.max_batch_sends, cfg.max_batch_payload = max_backends, max_batch_sends, max_batch_payload
        cfg.max_recv_records, cfg.max_recv_payload, cfg.list_pool_entries = max_recv_records, max_recv_payload, list_pool_entries
        cfg.fanout_variant = DEFAULT_FANOUT_VARIANT if fanout_variant is None else fanout_variant
        self.cfg = cfg
        self.max_agents = max_agents
        self._h = C.c_void_p()
        rc = self._L.sdb_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            msg = self._L.sdb_last_error(self._h).decode() if self._h else "sdb_create failed"
            if self._h:
                self._L.sdb_destroy(self._h)
                self._h = C.c_void_p()
            raise SdbError(rc, msg)
        self.max_recv_records = int(cfg.max_recv_records) or (1 << 20)
        self._keep = None          # keeps the last payload buffer alive while its H2D copy is in flight
        # pinned-or-pageable host output buffers, grown on demand
        self._out_hdr = np.zeros(0, HDR_DTYPE)
        self._out_pay = np.zeros(0, np.uint8)
        self._nb = 0

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise SdbError(rc, self._L.sdb_last_error(self._h).decode())

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.sdb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int) -> None:
        self._check(self._L.sdb_set_stream(self._h, C.c_void_p(cuda_stream)))

    def sync(self) -> None:
        self._check(self._L.sdb_sync(self._h))

    def stats(self) -> dict:
        s = SdbStats()
        self._check(self._L.sdb_get_stats(self._h, C.byref(s)))
        return {n: int(getattr(s, n)) for n, _ in SdbStats._fields_}

    def debug_set_arena_pos(self, granules: int) -> None:
        self._check(self._L.sdb_debug_set_arena_pos(self._h, granules))

    def advance_seq(self, next_seq: int) -> None:
        """Move the handle's sequence counter forward (ids of a refused batch are never reissued)."""
        self._check(self._L.sdb_advance_seq(self._h, next_seq))

    PROFILE_KINDS = ["p2p", "fanout", "commit", "recv_count", "recv_scan", "recv_select", "recv_gather",
                     "arena_floor", "pick", "xshard", "index", "xwait"]

    def profile(self, enable: bool) -> None:
        self._check(self._L.sdb_profile(self._h, 1 if enable else 0))

    def profile_read(self) -> dict:
        """{kernel class: (total ms, launches)} measured with CUDA events on the launching stream."""
        ms = np.zeros(16, np.float64)
        cnt = np.zeros(16, np.uint64)
        self._check(self._L.sdb_profile_read(self._h, _p(ms), _p(cnt)))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.PROFILE_KINDS)}

    # ------------------------------------------------------------------ registry / groups
    def register(self, idx) -> None:
        a = _arr(np.atleast_1d(idx), np.uint32)
        self._check(self._L.sdb_register_agents(self._h, len(a), _p(a)))

    def deregister(self, idx) -> None:
        a = _arr(np.atleast_1d(idx), np.uint32)
        self._check(self._L.sdb_deregister_agents(self._h, len(a), _p(a)))

    def create_group(self, g: int, members) -> None:
        m = _arr(members, np.uint32)
        self._check(self._L.sdb_create_group(self._h, g, len(m), _p(m)))

    # ------------------------------------------------------------------ enqueue
    @staticmethod
    def _common(n, prio, typ, lens, payload_off, payload, ts):
        prio = None if prio is None else _arr(prio, np.uint8)
        typ = None if typ is None else _arr(typ, np.uint8)
        lens = _arr(lens, np.uint16)
        payload_off = _arr(payload_off, np.uint64)
        payload = _arr(payload, np.uint8)
        ts = None if ts is None else _arr(ts, np.float64)
        assert len(lens) == n and len(payload_off) == n
        return prio, typ, lens, payload_off, payload, ts

    def send_batch(self, sender, receiver, prio, typ, lens, payload_off, payload, ts=None) -> int:
        s, r = _arr(sender, np.uint32), _arr(receiver, np.uint32)
        prio, typ, lens, po, pl, ts = self._common(len(s), prio, typ, lens, payload_off, payload, ts)
        base = C.c_uint64(0)
        self._keep = pl
        self._check(self._L.sdb_send_batch(self._h, len(s), _p(s), _p(r), _p(prio), _p(typ), _p(lens), _p(po), _p(pl),
                                           pl.nbytes, _p(ts), C.cast(C.byref(base), C.c_void_p)))
        return base.value

    def send_group_batch(self, sender, group, prio, typ, lens, payload_off, payload, ts=None) -> int:
        s, g = _arr(sender, np.uint32), _arr(group, np.uint32)
        prio, typ, lens, po, pl, ts = self._common(len(s), prio, typ, lens, payload_off, payload, ts)
        base = C.c_uint64(0)
        self._keep = pl
        self._check(self._L.sdb_send_group_batch(self._h, len(s), _p(s), _p(g), _p(prio), _p(typ), _p(lens), _p(po),
                                                 _p(pl), pl.nbytes, _p(ts), C.cast(C.byref(base), C.c_void_p)))
        return base.value

    def send_list_batch(self, sender, list_off, list_idx, prio, typ, lens, payload_off, payload, ts=None) -> int:
        s = _arr(sender, np.uint32)
        lo, li = _arr(list_off, np.uint64), _arr(list_idx, np.uint32)
        prio, typ, lens, po, pl, ts = self._common(len(s), prio, typ, lens, payload_off, payload, ts)
        base = C.c_uint64(0)
        self._keep = pl
        self._check(self._L.sdb_send_list_batch(self._h, len(s), _p(s), _p(lo), _p(li), _p(prio), _p(typ), _p(lens),
                                                _p(po), _p(pl), pl.nbytes, _p(ts), C.cast(C.byref(base), C.c_void_p)))
        return base.value

    def send_mixed_batch(self, sender, kind, target, list_off, list_idx, prio, typ, lens, payload_off, payload,
                         ts=None) -> int:
        """kind[i]: 0 p2p (target = receiver), 1 group (target = group), 2 list (target = list number)."""
        s, k, t = _arr(sender, np.uint32), _arr(kind, np.uint8), _arr(target, np.uint32)
        lo = _arr(list_off if list_off is not None else [0], np.uint64)
        li = _arr(list_idx if list_idx is not None else [], np.uint32)
        prio, typ, lens, po, pl, ts = self._common(len(s), prio, typ, lens, payload_off, payload, ts)
        base = C.c_uint64(0)
        self._keep = pl
        self._check(self._L.sdb_send_mixed_batch(self._h, len(s), _p(s), _p(k), _p(t), len(lo) - 1, _p(lo), _p(li),
                                                 _p(prio), _p(typ), _p(lens), _p(po), _p(pl), pl.nbytes, _p(ts),
                                                 C.cast(C.byref(base), C.c_void_p)))
        return base.value

    def stage(self, kind: int, sender, second, prio, typ, lens, payload_off, payload, ts=None) -> int:
        """Copy a p2p (kind 0) or group (kind 1) batch to device memory; returns a staged handle."""
        s, g = _arr(sender, np.uint32), _arr(second, np.uint32)
        prio, typ, lens, po, pl, ts = self._common(len(s), prio, typ, lens, payload_off, payload, ts)
        out = C.c_void_p()
        self._check(self._L.sdb_stage_batch(self._h, kind, len(s), _p(s), _p(g), _p(prio), _p(typ), _p(lens), _p(po),
                                            _p(pl), pl.nbytes, _p(ts), C.byref(out)))
        return out.value

    def submit(self, staged: int) -> int:
        base = C.c_uint64(0)
        self._check(self._L.sdb_submit_staged(self._h, C.c_void_p(staged), C.cast(C.byref(base), C.c_void_p)))
        return base.value

    def free_staged(self, staged: int) -> None:
        self._check(self._L.sdb_free_staged(self._h, C.c_void_p(staged)))

    # ------------------------------------------------------------------ cross-shard
    def wire_bytes(self, max_sends: int, max_payload: int) -> int:
        return int(self._L.sdb_wire_bytes(self._h, max_sends, max_payload))

    def set_agent_shards(self, shard_of) -> None:
        s = _arr(shard_of, np.uint8)
        self._check(self._L.sdb_set_agent_shards(self._h, len(s), _p(s)))
        full = np.full(self.max_agents, int(self.cfg.shard_id), np.uint8)
        full[: len(s)] = s
        self.n_owned = int((full == int(self.cfg.shard_id)).sum()) if int(self.cfg.num_shards) > 1 else None

    def export_group_batch(self, sender, group, prio, typ, lens, payload_off, payload, wire_dev: int, wire_cap: int,
                           ts=None) -> None:
        """Write this rank's batch of group sends as one wire batch into device memory at `wire_dev`."""
        s, g = _arr(sender, np.uint32), _arr(group, np.uint32)
        prio, typ, lens, po, pl, ts = self._common(len(s), prio, typ, lens, payload_off, payload, ts)
        self._keep = pl
        self._check(self._L.sdb_export_group_batch(self._h, len(s), _p(s), _p(g), _p(prio), _p(typ), _p(lens), _p(po),
                                                   _p(pl), pl.nbytes, _p(ts), C.c_void_p(wire_dev), wire_cap))

    def export_mixed_batch(self, sender, kind, target, list_off, list_idx, prio, typ, lens, payload_off, payload,
                           wire_dev: int, wire_cap: int, ts=None, seq_base: int = 0) -> None:
        """seq_base != 0: the exporter numbers its own sends from that (composite) sequence number."""
        s, k, t = _arr(sender, np.uint32), _arr(kind, np.uint8), _arr(target, np.uint32)
        lo = _arr(list_off if list_off is not None else [0], np.uint64)
        li = _arr(list_idx if list_idx is not None else [], np.uint32)
        prio, typ, lens, po, pl, ts = self._common(len(s), prio, typ, lens, payload_off, payload, ts)
        self._keep = pl
        if seq_base:
            self._check(self._L.sdb_export_mixed_batch_seq(self._h, seq_base, len(s), _p(s), _p(k), _p(t), len(lo) - 1, _p(lo),
                                                           _p(li), _p(prio), _p(typ), _p(lens), _p(po), _p(pl), pl.nbytes,
                                                           _p(ts), C.c_void_p(wire_dev), wire_cap))
            return
        self._check(self._L.sdb_export_mixed_batch(self._h, len(s), _p(s), _p(k), _p(t), len(lo) - 1, _p(lo), _p(li),
                                                   _p(prio), _p(typ), _p(lens), _p(po), _p(pl), pl.nbytes, _p(ts),
                                                   C.c_void_p(wire_dev), wire_cap))

    def wire_alloc(self, nbytes: int):
        """Device export buffer + its 64-byte CUDA-IPC handle (bytes) for the peer-memory transport."""
        dev = C.c_void_p()
        handle = C.create_string_buffer(64)
        self._check(self._L.sdb_wire_alloc(self._h, nbytes, C.byref(dev), C.cast(handle, C.c_void_p)))
        return dev.value, handle.raw

    def wire_open(self, ipc_handle: bytes) -> int:
        dev = C.c_void_p()
        buf = C.create_string_buffer(ipc_handle, 64)
        self._check(self._L.sdb_wire_open(self._h, C.cast(buf, C.c_void_p), C.byref(dev)))
        return dev.value

    def wire_close(self, dev: int, opened: bool) -> None:
        self._check(self._L.sdb_wire_close(self._h, C.c_void_p(dev), 1 if opened else 0))

    def import_wire_ptrs(self, ptrs) -> int:
        """Import one wire batch per source rank from a table of device pointers (rank order)."""
        arr = (C.c_void_p * len(ptrs))(*[C.c_void_p(int(x)) for x in ptrs])
        base = C.c_uint64(0)
        self._check(self._L.sdb_import_wire_ptrs(self._h, len(ptrs), C.cast(arr, C.c_void_p), C.cast(C.byref(base), C.c_void_p)))
        return base.value

    # flag-synchronised, host-asynchronous peer transport (see include/swarmdb_b200.h)
    @staticmethod
    def _ptr_table(ptrs):
        return (C.c_void_p * len(ptrs))(*[C.c_void_p(int(x)) for x in ptrs])

    def wire_wait_done(self, ptrs, wire_bytes: int, step: int) -> None:
        """Stream-ordered wait until the owner of every listed buffer has finished importing `step`."""
        arr = self._ptr_table(ptrs)
        self._check(self._L.sdb_wire_wait_done(self._h, len(ptrs), C.cast(arr, C.c_void_p), wire_bytes, step))

    def wire_publish(self, wire_dev: int, wire_bytes: int, step: int) -> None:
        """Stream-ordered: this buffer now holds the complete export of `step`."""
        self._check(self._L.sdb_wire_publish(self._h, C.c_void_p(wire_dev), wire_bytes, step))

    def overflow_log(self, cap: int = 4096):
        """(agents, seqs, dropped): the records that found their receiver's ring full since the last call
        (include/swarmdb_b200.h: sdb_overflow_log); reading clears the log."""
        ag = np.zeros(cap, np.uint32); sq = np.zeros(cap, np.uint64)
        n = C.c_uint32(0); dropped = C.c_uint64(0)
        self._check(self._L.sdb_overflow_log(self._h, cap, _p(ag), _p(sq), C.cast(C.byref(n), C.c_void_p),
                                             C.cast(C.byref(dropped), C.c_void_p)))
        return ag[:n.value].copy(), sq[:n.value].copy(), int(dropped.value)

    def import_prefetch(self, ptrs, wire_bytes: int, step: int) -> None:
        """Run the flag wait + localize pass of `step` on the handle's prefetch stream, beside what is enqueued next
        (include/swarmdb_b200.h: sdb_import_prefetch); the following import_wire_ptrs_async(step) only places it."""
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        self._check(self._L.sdb_import_prefetch(self._h, len(ptrs), C.cast(arr, C.c_void_p), wire_bytes, step))

    def import_wire_ptrs_async(self, ptrs, wire_bytes: int, step: int) -> None:
        """Wait on the device for every source's `step`, import without a host round trip, report done."""
        arr = self._ptr_table(ptrs)
        self._check(self._L.sdb_import_wire_ptrs_async(self._h, len(ptrs), C.cast(arr, C.c_void_p), wire_bytes, step))

    def import_wire_batches(self, n_src: int, wire_dev_all: int, stride: int) -> int:
        """Expand the wire batches of `n_src` ranks (rank order, `stride` bytes apart) for the agents this shard owns."""
        base = C.c_uint64(0)
        self._check(self._L.sdb_import_wire_batches(self._h, n_src, C.c_void_p(wire_dev_all), stride,
                                                    C.cast(C.byref(base), C.c_void_p)))
        return base.value

    # ------------------------------------------------------------------ dequeue
    def receive_batch(self, agents, max_messages: int, flags: int = 0, copy_out: bool = True,
                      out_hdr: Optional[np.ndarray] = None, out_payload: Optional[np.ndarray] = None, wait: bool = True):
        """Returns (counts[n_agents], headers[total], payload bytes) - views into reusable buffers
        unless explicit output arrays are given.  copy_out=False leaves results on the device and returns
        (None, total, payload_bytes); with wait=False the call only enqueues the receive (SDB_RECV_ASYNC,
        totals later through last_receive_totals()) and returns (None, None, None)."""
        if not wait:
            if copy_out:
                raise ValueError("wait=False needs copy_out=False")
            flags |= RECV_ASYNC
        if agents is None:
            a, n = None, 0
            counts = np.zeros(self.max_agents, np.uint32) if copy_out else None
        else:
            a = _arr(agents, np.uint32)
            n = len(a)
            counts = np.zeros(n, np.uint32) if copy_out else None
        total, pbytes = C.c_uint64(0), C.c_uint64(0)
        hdr = pay = None
        hdr_cap = pay_cap = 0
        if copy_out:
            hdr = out_hdr if out_hdr is not None else self._grow_hdr()
            pay = out_payload if out_payload is not None else self._grow_pay()
            hdr_cap, pay_cap = len(hdr), pay.nbytes
        self._check(self._L.sdb_receive_batch(self._h, n, _p(a), max_messages, flags, _p(counts), _p(hdr), hdr_cap,
                                              _p(pay), pay_cap, C.cast(C.byref(total), C.c_void_p),
                                              C.cast(C.byref(pbytes), C.c_void_p)))
        if not wait:
            return None, None, None
        if not copy_out:
            return None, total.value, pbytes.value
        if agents is None:
            owned = (flags & RECV_OWNED) and getattr(self, "n_owned", None) is not None
            counts = counts[: self.n_owned if owned else self.stats_n_agents()]
        return counts, hdr[: total.value], pay[: pbytes.value]

    def latency_server(self, enable: bool) -> None:
        """Start / stop the persistent low-latency dequeue server (single-agent receives then skip launch, sync and D2H)."""
        self._check(self._L.sdb_latency_server(self._h, 1 if enable else 0))

    def last_receive_totals(self):
        """(records, payload bytes) of the last receive call; waits for it."""
        total, pbytes = C.c_uint64(0), C.c_uint64(0)
        self._check(self._L.sdb_last_receive_totals(self._h, C.cast(C.byref(total), C.c_void_p), C.cast(C.byref(pbytes), C.c_void_p)))
        return total.value, pbytes.value

    # stream digests (definition in include/swarmdb_b200.h): order + content of every agent's stream, on the device
    def digest_reset(self) -> None:
        self._check(self._L.sdb_digest_reset(self._h))

    def digest_fold(self) -> None:
        """Fold the records of the last bulk receive (still on the device) into the per-agent digests."""
        self._check(self._L.sdb_digest_fold(self._h))

    def digest_read(self, agents=None) -> np.ndarray:
        if agents is None:
            a, n = None, self.max_agents
        else:
            a = _arr(agents, np.uint32)
            n = len(a)
        out = np.zeros(n, np.uint64)
        self._check(self._L.sdb_digest_read(self._h, n, _p(a), _p(out)))
        return out

    def receive_one(self, agent: int, max_messages: int = 100, flags: int = 0):
        """Latency path for a single agent (one kernel launch, one D2H): returns (headers, payload)
        views into reusable buffers.  Pre-built ctypes arguments keep the Python overhead minimal."""
        st = getattr(self, "_one", None)
        if st is None:
            mp = pad32(max(int(self.cfg.max_payload_bytes), 1))
            cap = max(1, min(1024, (1 << 22) // mp))          # records per call: a 4 MiB payload window
            hdr = np.zeros(cap, HDR_DTYPE)
            pay = np.zeros(cap * mp, np.uint8)
            a = np.zeros(1, np.uint32); cnt = np.zeros(1, np.uint32)
            total, pbytes = C.c_uint64(0), C.c_uint64(0)
            st = self._one = (a, cnt, hdr, pay, total, pbytes, _p(a), _p(cnt), _p(hdr), _p(pay),
                              C.cast(C.byref(total), C.c_void_p), C.cast(C.byref(pbytes), C.c_void_p), cap, pay.nbytes)
        a, cnt, hdr, pay, total, pbytes, pa, pc, ph, pp, pt, pb, cap, pcap = st
        a[0] = agent
        k = max_messages if max_messages <= cap else cap
        rc = self._L.sdb_receive_batch(self._h, 1, pa, k, flags, pc, ph, cap, pp, pcap, pt, pb)
        if rc != 0:
            self._check(rc)
        return hdr[: total.value], pay[: pbytes.value], k

    def stats_n_agents(self) -> int:
        return self.stats()["n_agents"]

    def _grow_hdr(self) -> np.ndarray:
        if len(self._out_hdr) < self.max_recv_records:
            self._out_hdr = np.zeros(self.max_recv_records, HDR_DTYPE)
        return self._out_hdr

    def _grow_pay(self) -> np.ndarray:
        need = int(self.cfg.max_recv_payload) or self.max_recv_records * 256
        if self._out_pay.nbytes < need:
            self._out_pay = np.zeros(need, np.uint8)
        return self._out_pay

    def last_receive_dev(self) -> Tuple[int, int, int]:
        c, h, p = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self._L.sdb_last_receive_dev(self._h, C.byref(c), C.byref(h), C.byref(p)))
        return c.value, h.value, p.value

    # ------------------------------------------------------------------ inbox / load queries from the rings (N3)
    def agent_loads(self, agents=None, n: Optional[int] = None) -> np.ndarray:
        """Per-agent `received` (inbox size), `pending` (unread) and its priority histogram, computed on the device."""
        if agents is None:
            a, cnt = None, int(n if n is not None else self.stats_n_agents())
        else:
            a = _arr(agents, np.uint32)
            cnt = len(a)
        out = np.zeros(cnt, LOAD_DTYPE)
        self._check(self._L.sdb_agent_loads(self._h, cnt, _p(a), _p(out)))
        return out

    def queue_stats(self) -> dict:
        raw = np.zeros(10, np.uint64)
        self._check(self._L.sdb_queue_stats(self._h, _p(raw)))
        d = {k: int(v) for k, v in zip(QSTATS_FIELDS, raw)}
        d["pending_by_prio"] = [d.pop("p0"), d.pop("p1"), d.pop("p2"), d.pop("p3")]
        return d

    def assign_agent_backends(self, agents, backends) -> None:
        a, b = _arr(agents, np.uint32), _arr(backends, np.uint32)
        self._check(self._L.sdb_assign_agent_backends(self._h, len(a), _p(a), _p(b)))

    def backend_loads_from_queues(self) -> None:
        """load[b] = pending records of the agents assigned to backend b (stream-ordered, no host round trip)."""
        self._check(self._L.sdb_backend_loads_from_queues(self._h))

    # ------------------------------------------------------------------ backends
    def set_backends(self, weight, load0=None) -> None:
        w = _arr(weight, np.uint32)
        l0 = None if load0 is None else _arr(load0, np.uint64)
        self._check(self._L.sdb_set_backends(self._h, len(w), _p(w), _p(l0)))
        self._nb = len(w)

    def backend_loads(self) -> np.ndarray:
        out = np.zeros(self._nb, np.uint64)
        self._check(self._L.sdb_get_backend_loads(self._h, self._nb, _p(out)))
        return out

    def release_backends(self, backend, cost=None) -> None:
        b = _arr(backend, np.uint32)
        c = None if cost is None else _arr(cost, np.uint32)
        self._check(self._L.sdb_release_backends(self._h, len(b), _p(b), _p(c)))

    def select_backends(self, n_req: int, cost=None, mode: int = 0, seed: int = 0) -> np.ndarray:
        c = None if cost is None else _arr(cost, np.uint32)
        out = np.zeros(n_req, np.uint32)
        self._check(self._L.sdb_select_backend_batch(self._h, n_req, _p(c), mode, seed, _p(out)))
        return out


def payload_offsets(hdr: np.ndarray) -> np.ndarray:
    """Byte offset of each record's payload inside the packed payload stream of a receive."""
    pl = pad32(hdr["len"].astype(np.int64))
    off = np.zeros(len(hdr), np.int64)
    if len(hdr) > 1:
        np.cumsum(pl[:-1], out=off[1:])
    return off
