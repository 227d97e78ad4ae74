"""`SwarmsDB` - the reference's Python surface over GPU-resident per-agent rings.

Mirrors the public interface of the reference class (/root/reference/swarmdb/" main.py", tag
`M:`; signatures in SURVEY.md section 8b(i)): same method names, argument meaning, defaults, return
values and error behaviour for the hot path

    register_agent / deregister_agent   M:314-372
    send_message                        M:393-519
    receive_messages                    M:521-601
    broadcast_message                   M:810-850
    add_agent_group (= create_group)    M:1208-1227
    send_to_group                       M:1229-1279
    set_llm_load_balancing / assign_llm_backend / get_llm_backend   M:1281-1325

What changed underneath: the Kafka producer/consumer pair (M:192-204, M:334-345) and the
JSON envelope (M:466, M:575-576) are gone.  Sends are staged in a host buffer (the role of
the producer's `linger.ms` batching, M:197) and flushed as one mixed batch through the C ABI
(`swarmdb_b200._native.Shard` -> libswarmdb_b200.so -> sm_100a kernels); receives pop the
agent's ring on the device and decode the returned bytes.  There is no CPU fallback.

Bulk entry points for index-level traffic (`send_to_group_batch`, `send_batch`,
`receive_batch`, `select_llm_backends`) bypass per-message Python objects entirely.
"""
from __future__ import annotations

import datetime
import json
import os
import secrets
import time
import uuid
from dataclasses import dataclass
from enum import Enum
from pathlib import Path
from typing import Any, Callable, Dict, Iterable, List, Optional, Set, Union

import numpy as np
from pydantic import BaseModel, Field

from . import _native
from ._native import NO_GROUP, RECV_PEEK, RECV_PRIORITY, TYPE_MASK, TYPEF_EXTRAS, TYPEF_JSON, SdbError, Shard, pad32


class RingOverflow(SdbError):
    """Some records of a flushed batch found their receiver's ring full.  `exact` tells whether the lost messages were
    identified (only they are FAILED) or the whole batch had to be marked."""
    exact = False


try:  # the reference logs through loguru; keep the same logger object when it is installed
    from loguru import logger
except Exception:  # pragma: no cover
    import logging

    logger = logging.getLogger("swarmdb_b200")


# ----------------------------------------------------------------------------- vocabularies (M:23-51)
class MessageType(str, Enum):
    CHAT = "chat"
    COMMAND = "command"
    FUNCTION_CALL = "function_call"
    FUNCTION_RESULT = "function_result"
    SYSTEM = "system"
    ERROR = "error"
    STATUS = "status"


class MessagePriority(int, Enum):
    LOW = 0
    NORMAL = 1
    HIGH = 2
    CRITICAL = 3


class MessageStatus(str, Enum):
    PENDING = "pending"
    DELIVERED = "delivered"
    READ = "read"
    PROCESSED = "processed"
    FAILED = "failed"


_TYPE_CODE = {t: i for i, t in enumerate(MessageType)}
_TYPE_BY_CODE = list(MessageType)


class Message(BaseModel):
    """Same eleven fields and defaults as the reference model (M:54-82)."""

    id: str = Field(default_factory=lambda: str(uuid.uuid4()))
    sender_id: str
    receiver_id: Optional[str] = None
    content: Union[str, Dict[str, Any], List[Any]]
    type: MessageType = MessageType.CHAT
    priority: MessagePriority = MessagePriority.NORMAL
    timestamp: float = Field(default_factory=time.time)
    status: MessageStatus = MessageStatus.PENDING
    metadata: Dict[str, Any] = Field(default_factory=dict)
    token_count: Optional[int] = None
    visible_to: List[str] = Field(default_factory=list)

    def to_dict(self) -> Dict[str, Any]:
        """Plain dict with enum fields flattened to their values (the intent of M:91-98)."""
        d = self.model_dump()
        d["type"], d["priority"], d["status"] = self.type.value, self.priority.value, self.status.value
        return d

    @classmethod
    def from_dict(cls, data: Dict[str, Any]) -> "Message":
        d = dict(data)
        if isinstance(d.get("type"), str):
            d["type"] = MessageType(d["type"])
        if isinstance(d.get("priority"), int):
            d["priority"] = MessagePriority(d["priority"])
        if isinstance(d.get("status"), str):
            d["status"] = MessageStatus(d["status"])
        return cls(**d)


@dataclass
class KafkaConfig:
    """Accepted for drop-in compatibility (M:114-127).  Only `num_partitions` has a meaning
    here (informational: the shard count comes from the process group); the other knobs
    configured librdkafka and are ignored."""

    bootstrap_servers: str = "localhost:9092"
    group_id: str = "agent_messaging_system"
    auto_offset_reset: str = "earliest"
    num_partitions: int = 3
    replication_factor: int = 1
    retention_ms: int = 604800000
    max_poll_interval_ms: int = 300000
    session_timeout_ms: int = 30000
    heartbeat_interval_ms: int = 10000
    consumer_timeout_ms: int = 1000


@dataclass
class GpuConfig:
    """Sizing of the GPU-resident queue (replaces broker-side capacity planning)."""

    device: int = 0
    max_agents: int = 1 << 16
    ring_slots: int = 256                 # pending messages per agent (power of two)
    arena_bytes: int = 1 << 28            # message log (power of two)
    max_payload_bytes: int = 32768
    max_groups: int = 4096
    member_pool_entries: int = 0          # 0: 4 x max_agents
    max_backends: int = 256
    flush_threshold: int = 4096           # buffered sends before an automatic flush
    max_recv_records: int = 1 << 16
    max_recv_payload: int = 1 << 26
    priority_dequeue: bool = False        # True: receive in (priority desc, arrival) order - extension
    deterministic_ids: bool = False       # True: ids are uuid.UUID(int=seq) (tests / reproducibility)
    id_nonce: Optional[int] = None        # high 64 bits of every id (must match across the ranks of a sharded deployment)
    fanout_variant: Optional[int] = None      # None: the library default (_native.DEFAULT_FANOUT_VARIANT)
    shard_id: int = 0                     # multi-GPU: this process's shard / total shards (see sharded.ShardedSwarmsDB)
    num_shards: int = 1


def _encode_content(content: Any) -> (bytes, int):
    if isinstance(content, str):
        return content.encode("utf-8"), 0
    return json.dumps(content).encode("utf-8"), TYPEF_JSON


class _TransportAdmin:
    """What callers probe through `db.admin_client` (the REST health check lists topics as a liveness test,
    api.py:798): here the "topic" is the device queue and its "partitions" are the shards."""

    def __init__(self, db: "SwarmsDB") -> None:
        self._db = db

    def list_topics(self, timeout: Optional[float] = None):
        from types import SimpleNamespace
        self._db.shard.stats()                     # raises if the device transport is gone
        parts = {i: SimpleNamespace(id=i) for i in range(max(1, self._db.gpu_config.num_shards))}
        return SimpleNamespace(topics={self._db.base_topic: SimpleNamespace(partitions=parts)})


class SwarmsDB:
    """Agent message queue + LLM-backend balancer; reference-compatible surface (see module doc)."""

    def __init__(
        self,
        base_topic: str = "agent_messaging",
        config: Optional[KafkaConfig] = None,
        save_dir: Optional[Union[str, Path]] = None,
        auto_save: bool = True,
        save_interval: int = 300,
        max_messages_per_file: int = 10000,
        token_counter: Optional[Callable[[str], int]] = None,
        gpu_config: Optional[GpuConfig] = None,
        _shard: Any = None,
    ):
        self.base_topic = base_topic
        self.config = config or KafkaConfig()
        self.gpu_config = gpu_config or GpuConfig()
        g = self.gpu_config
        # the transport: constructing it fails loudly when the CUDA library or a device is missing
        self.shard = _shard if _shard is not None else Shard(
            max_agents=g.max_agents, ring_slots=g.ring_slots, arena_bytes=g.arena_bytes,
            max_payload_bytes=g.max_payload_bytes, max_groups=g.max_groups,
            member_pool_entries=g.member_pool_entries or 4 * g.max_agents, max_backends=g.max_backends,
            max_batch_sends=max(g.flush_threshold, 1), max_batch_payload=max(1 << 22, 2 * pad32(g.max_payload_bytes)),
            max_recv_records=g.max_recv_records, max_recv_payload=g.max_recv_payload, device=g.device,
            fanout_variant=g.fanout_variant, shard_id=g.shard_id, num_shards=g.num_shards)

        # local state, same names as the reference (M:210-233)
        self.messages: Dict[str, Message] = {}
        self.agent_inbox: Dict[str, List[str]] = {}
        self.message_count = 0
        self.registered_agents: Set[str] = set()
        self.metadata: Dict[str, Any] = {}
        self.token_counter = token_counter
        self.llm_load_balancing = False

        self.save_dir = Path(save_dir) if save_dir else Path(os.getcwd()) / "message_history"
        self.save_dir.mkdir(parents=True, exist_ok=True)
        self.auto_save = auto_save
        self.save_interval = save_interval
        self.max_messages_per_file = max_messages_per_file
        self.last_save_time = time.time()

        # index maps (the device only knows dense indices)
        self._agent_idx: Dict[str, int] = {}
        self._agent_name: List[str] = []
        self._group_idx: Dict[str, int] = {}
        self._group_name: List[str] = []
        self._group_snapshot: List[List[str]] = []
        self._backend_idx: Dict[str, int] = {}
        self._backend_name: List[str] = []
        self._id_hi = 0 if g.deterministic_ids else (((g.id_nonce if g.id_nonce is not None else secrets.randbits(63)) & ((1 << 63) - 1)) << 64)
        self._next_seq = 1                      # mirror of the handle's sequence counter
        self._list_cap = 2 * g.max_agents + 1024 - 8      # the handle's default list_pool_entries, minus slack
        self._seq_to_id: Dict[int, str] = {}
        self._reset_buffer()
        self._closed = False
        self.admin_client = _TransportAdmin(self)          # M:202-204: callers only list topics through it
        logger.info(f"SwarmsDB (B200) initialized with base topic: {base_topic}")

    # ------------------------------------------------------------------ helpers
    def _reset_buffer(self) -> None:
        self._b_sender: List[int] = []
        self._b_kind: List[int] = []
        self._b_target: List[int] = []
        self._b_prio: List[int] = []
        self._b_type: List[int] = []
        self._b_len: List[int] = []
        self._b_off: List[int] = []
        self._b_ts: List[float] = []
        self._b_payload = bytearray()
        self._b_list_off: List[int] = [0]
        self._b_list_idx: List[int] = []
        self._b_msgs: List[Message] = []
        self._b_first_seq = self._next_seq

    def _index(self, agent_id: str) -> int:
        i = self._agent_idx.get(agent_id)
        if i is None:
            i = len(self._agent_name)
            if i >= self.gpu_config.max_agents:
                raise RuntimeError(f"agent capacity exhausted (GpuConfig.max_agents={self.gpu_config.max_agents})")
            self._agent_idx[agent_id] = i
            self._agent_name.append(agent_id)
        return i

    def agent_index(self, agent_id: str) -> int:
        """Dense device index of an agent id (assigned on first mention)."""
        return self._index(agent_id)

    def _make_id(self, seq: int) -> str:
        h = "%032x" % (self._id_hi | seq)                 # same text as str(uuid.UUID(int=...)), 4x cheaper
        return f"{h[:8]}-{h[8:12]}-{h[12:16]}-{h[16:20]}-{h[20:]}"

    def _count_tokens(self, content: Any) -> int:          # M:295-307
        if self.token_counter is None:
            return 0
        text = json.dumps(content) if isinstance(content, (dict, list)) else str(content)
        return self.token_counter(text)

    def _stage(self, kind: int, sender: int, target: int, prio: int, type_code: int, payload: bytes, ts: float) -> None:
        if len(payload) > self.gpu_config.max_payload_bytes:
            raise ValueError(f"encoded message is {len(payload)} bytes; GpuConfig.max_payload_bytes="
                             f"{self.gpu_config.max_payload_bytes}")
        self._b_sender.append(sender); self._b_kind.append(kind); self._b_target.append(target)
        self._b_prio.append(prio); self._b_type.append(type_code); self._b_len.append(len(payload))
        self._b_off.append(len(self._b_payload)); self._b_ts.append(ts)
        self._b_payload += payload
        self._b_payload += b"\0" * (pad32(len(payload)) - len(payload))

    def flush(self) -> None:
        """Push every buffered send to the device (one mixed batch, call order preserved)."""
        n = len(self._b_sender)
        if n == 0:
            return
        msgs = self._b_msgs
        try:
            payload = np.frombuffer(bytes(self._b_payload) + b"\0" * 32, dtype=np.uint8)
            base = self.shard.send_mixed_batch(
                np.asarray(self._b_sender, np.uint32), np.asarray(self._b_kind, np.uint8),
                np.asarray(self._b_target, np.uint32), np.asarray(self._b_list_off, np.uint64),
                np.asarray(self._b_list_idx, np.uint32), np.asarray(self._b_prio, np.uint8),
                np.asarray(self._b_type, np.uint8), np.asarray(self._b_len, np.uint16),
                np.asarray(self._b_off, np.uint64), payload, np.asarray(self._b_ts, np.float64))
            if base != self._b_first_seq:
                raise RuntimeError(f"sequence mirror out of step: device {base}, host {self._b_first_seq}")
            st = self.shard.stats()
            if st["ring_overflow"] > getattr(self, "_seen_overflow", 0):
                lost = st["ring_overflow"] - getattr(self, "_seen_overflow", 0)
                self._seen_overflow = st["ring_overflow"]
                err = RingOverflow(-4, f"{lost} message(s) not enqueued: a receiver's ring is full "
                                       f"(GpuConfig.ring_slots={self.gpu_config.ring_slots})")
                err.exact = self._mark_overflowed(msgs, str(err))
                raise err
        except Exception as e:                  # M:501-519: mark FAILED, keep the error, re-raise
            if not (isinstance(e, RingOverflow) and e.exact):      # the device said which records were lost: only those failed
                for m in msgs:
                    m.status = MessageStatus.FAILED
                    m.metadata["error"] = str(e)
            # ids already handed out stay used: the device counter moves up to the host's, never the other way
            self.shard.advance_seq(self._next_seq)
            self._reset_buffer()
            raise
        self._reset_buffer()

    def _mark_overflowed(self, msgs: List[Message], error: str) -> bool:
        """Delivery report for a batch that overflowed rings (M:374-391): the device logs (receiver, sequence number)
        of every dropped record (sdb_overflow_log).  Point-to-point and group messages have one record each (id <->
        sequence number): exactly those are marked FAILED, the rest of the batch stays DELIVERED.  A broadcast is one
        message with many copies: it stays DELIVERED and names the receivers that missed it in
        metadata["undelivered_to"].  Returns False when the log could not name every dropped record (more than its
        capacity in one batch): the caller then fails the whole batch, as before."""
        agents, seqs, dropped = self.shard.overflow_log()
        if dropped != len(seqs):
            return False
        in_batch = {id(m) for m in msgs}
        for a, sq in zip(agents.tolist(), seqs.tolist()):
            m = self.messages.get(self._make_id(int(sq)))
            if m is None or id(m) not in in_batch:
                return False
            if m.receiver_id is None:
                m.metadata.setdefault("undelivered_to", []).append(self._agent_name[a] if a < len(self._agent_name) else a)
            else:
                m.status = MessageStatus.FAILED
                m.metadata["error"] = error
        return True

    # ------------------------------------------------------------------ registry
    def register_agent(self, agent_id: str) -> None:
        if agent_id in self.registered_agents:
            return
        self.registered_agents.add(agent_id)
        self.agent_inbox.setdefault(agent_id, [])
        self.shard.register(self._index(agent_id))
        logger.debug(f"Agent {agent_id} registered with messaging system")

    def register_agents(self, agent_ids: Iterable[str]) -> np.ndarray:
        """Bulk registration; returns the dense indices."""
        idx = np.fromiter((self._index(a) for a in agent_ids), dtype=np.uint32)
        for a in agent_ids:
            if a not in self.registered_agents:
                self.registered_agents.add(a)
                self.agent_inbox.setdefault(a, [])
        if len(idx):
            self.shard.register(idx)
        return idx

    def deregister_agent(self, agent_id: str) -> None:
        if agent_id not in self.registered_agents:
            logger.warning(f"Agent {agent_id} not registered")
            return
        self.registered_agents.remove(agent_id)
        self.shard.deregister(self._index(agent_id))

    # ------------------------------------------------------------------ send
    def send_message(
        self,
        sender_id: str,
        content: Union[str, Dict[str, Any], List[Any]],
        receiver_id: Optional[str] = None,
        message_type: MessageType = MessageType.CHAT,
        priority: MessagePriority = MessagePriority.NORMAL,
        metadata: Optional[Dict[str, Any]] = None,
        visible_to: Optional[List[str]] = None,
        _group: int = NO_GROUP,
    ) -> str:
        if sender_id not in self.registered_agents:
            self.register_agent(sender_id)
        if receiver_id is not None and receiver_id not in self.registered_agents:
            self.register_agent(receiver_id)
        token_count = self._count_tokens(content) if self.token_counter else None

        seq = self._next_seq
        if not isinstance(content, (str, dict, list)):
            raise TypeError("content must be str, dict or list")
        # fields are normalised here, so pydantic's per-field validation (14 % of the reference's send cost) is skipped
        message = Message.model_construct(
            id=self._make_id(seq), sender_id=str(sender_id), receiver_id=receiver_id, content=content,
            type=MessageType(message_type), priority=MessagePriority(priority), timestamp=time.time(),
            status=MessageStatus.PENDING, metadata=dict(metadata or {}), token_count=token_count,
            visible_to=list(visible_to or []))
        if receiver_id is None and not message.visible_to:
            message.visible_to = list(self.registered_agents)           # M:449-450

        self.messages[message.id] = message
        self.message_count += 1
        if receiver_id is not None:
            if receiver_id in self.agent_inbox:
                self.agent_inbox[receiver_id].append(message.id)
        else:
            for agent_id in self.registered_agents:                      # M:461-463
                self.agent_inbox[agent_id].append(message.id)

        # wire form: content bytes (+ JSON extras for metadata / visible_to / token_count)
        body, flags = _encode_content(content)
        extras: Dict[str, Any] = {}
        md = dict(message.metadata)
        if _group != NO_GROUP:
            md.pop("group", None)                                        # travels as header.group
        if md:
            extras["m"] = md
        if message.visible_to:
            extras.update(self._encode_visibility(message.visible_to))
        if token_count is not None:
            extras["t"] = token_count
        if extras:
            payload = len(body).to_bytes(4, "little") + body + json.dumps(extras).encode("utf-8")
            flags |= TYPEF_EXTRAS
        else:
            payload = body
        type_code = _TYPE_CODE[MessageType(message_type)] | flags
        prio = int(MessagePriority(priority))
        sender = self._index(sender_id)
        try:
            if receiver_id is not None:
                deliverable = (not message.visible_to) or (receiver_id in message.visible_to)    # M:579-585
                if deliverable:
                    self._stage(0, sender, self._index(receiver_id), prio, type_code, payload, message.timestamp)
                else:      # in the log, matched by no consumer: an empty recipient list
                    self._stage_list(sender, [], prio, type_code, payload, message.timestamp)
            else:
                seen: Set[int] = set()
                recips = []
                for a in message.visible_to:
                    i = self._index(a)
                    if i not in seen:
                        seen.add(i); recips.append(i)
                self._stage_list(sender, recips, prio, type_code, payload, message.timestamp)
        except Exception as e:
            message.status = MessageStatus.FAILED
            message.metadata["error"] = str(e)
            raise
        self._next_seq += 1
        self._b_msgs.append(message)
        message.status = MessageStatus.DELIVERED                          # delivery report, M:374-391
        self._after_send()
        return message.id

    def _encode_visibility(self, visible_to: List[str]) -> Dict[str, Any]:
        """How `visible_to` (M:82) travels inside each copy's payload.  Short lists go as names ("v").  A broadcast
        to (nearly) everybody would put O(agents) bytes into every one of O(agents) copies, so long lists travel
        as their COMPLEMENT over the dense agent indices known at send time: "vc" = [n, [indices < n that are NOT
        visible]] (sender, excluded and not-yet/no-longer registered agents).  Indices agree on every rank of a
        sharded deployment (the registry is replicated), so any receiver reconstructs the exact set."""
        if len(visible_to) <= 32:
            return {"v": visible_to}
        n = len(self._agent_name)
        vis = {self._agent_idx[a] for a in visible_to if a in self._agent_idx}
        if len(vis) != len(set(visible_to)):                  # names the registry has never seen: keep them literal
            return {"v": visible_to}
        comp = [i for i in range(n) if i not in vis]
        if len(comp) >= len(visible_to):
            return {"v": visible_to}
        return {"vc": [n, comp]}

    def _decode_visibility(self, extras: Dict[str, Any]) -> List[str]:
        if "vc" in extras:
            n, comp = extras["vc"]
            skip = set(comp)
            return [self._agent_name[i] for i in range(min(n, len(self._agent_name))) if i not in skip]
        return list(extras.get("v", []))

    def _stage_list(self, sender: int, recips: List[int], prio: int, type_code: int, payload: bytes, ts: float) -> None:
        # recipient lists (and the one-entry lists of buffered point-to-point sends) share the device's per-batch
        # list pool: flush first if this list would not fit beside what is already buffered
        if self._b_sender and len(self._b_list_idx) + len(recips) + len(self._b_sender) + 1 > self._list_cap:
            self.flush()
        self._stage(2, sender, len(self._b_list_off) - 1, prio, type_code, payload, ts)
        self._b_list_idx.extend(recips)
        self._b_list_off.append(len(self._b_list_idx))

    def _after_send(self) -> None:
        if len(self._b_sender) >= self.gpu_config.flush_threshold or len(self._b_payload) >= (1 << 21):
            self.flush()
        if self.auto_save and (time.time() - self.last_save_time > self.save_interval
                               or self.message_count % self.max_messages_per_file == 0):
            self.save_message_history()

    def broadcast_message(
        self,
        sender_id: str,
        content: Union[str, Dict[str, Any], List[Any]],
        message_type: MessageType = MessageType.CHAT,
        priority: MessagePriority = MessagePriority.NORMAL,
        metadata: Optional[Dict[str, Any]] = None,
        exclude_agents: Optional[List[str]] = None,
    ) -> str:
        excluded = set(exclude_agents or [])
        visible = [a for a in self.registered_agents if a != sender_id and a not in excluded]
        return self.send_message(sender_id=sender_id, content=content, receiver_id=None, message_type=message_type,
                                 priority=priority, metadata=metadata, visible_to=visible)

    # ------------------------------------------------------------------ groups
    def add_agent_group(self, group_name: str, agent_ids: List[str]) -> None:
        groups = self.metadata.setdefault("agent_groups", {})
        groups[group_name] = agent_ids                    # the caller's list object, like M:1223
        self._sync_group(group_name)
        logger.debug(f"Agent group '{group_name}' created with {len(agent_ids)} agents")

    create_group = add_agent_group

    def _sync_group(self, group_name: str) -> int:
        """Upload the group's member list if it is new or was mutated by the caller."""
        members = self.metadata["agent_groups"][group_name]
        g = self._group_idx.get(group_name)
        if g is None:
            g = len(self._group_name)
            if g >= self.gpu_config.max_groups:
                raise RuntimeError("group capacity exhausted (GpuConfig.max_groups)")
            self._group_idx[group_name] = g
            self._group_name.append(group_name)
            self._group_snapshot.append(None)
        if self._group_snapshot[g] != members:
            self.flush()                                   # earlier sends used the old membership
            self.shard.create_group(g, np.fromiter((self._index(a) for a in members), dtype=np.uint32,
                                                   count=len(members)))
            self._group_snapshot[g] = list(members)
        return g

    def send_to_group(
        self,
        sender_id: str,
        group_name: str,
        content: Union[str, Dict[str, Any], List[Any]],
        message_type: MessageType = MessageType.CHAT,
        priority: MessagePriority = MessagePriority.NORMAL,
        metadata: Optional[Dict[str, Any]] = None,
    ) -> List[str]:
        groups = self.metadata.get("agent_groups", {})
        if group_name not in groups:
            logger.warning(f"Agent group '{group_name}' not found")
            return []
        members = groups[group_name]
        msg_metadata = metadata or {}
        msg_metadata["group"] = group_name               # mutates a non-empty caller dict, like M:1263-1264
        g = self._sync_group(group_name)

        if sender_id not in self.registered_agents:
            self.register_agent(sender_id)
        token_count = self._count_tokens(content) if self.token_counter else None
        body, flags = _encode_content(content)
        md = {k: v for k, v in msg_metadata.items() if k != "group"}
        extras: Dict[str, Any] = {}
        if md:
            extras["m"] = md
        if token_count is not None:
            extras["t"] = token_count
        if extras:
            payload = len(body).to_bytes(4, "little") + body + json.dumps(extras).encode("utf-8")
            flags |= TYPEF_EXTRAS
        else:
            payload = body
        base = self._next_seq
        now = time.time()
        mtype, mprio = MessageType(message_type), MessagePriority(priority)
        if not isinstance(content, (str, dict, list)):
            raise TypeError("content must be str, dict or list")
        ids: List[str] = []
        staged = False
        for j, agent_id in enumerate(members):
            if agent_id == sender_id:
                continue
            if agent_id not in self.registered_agents:
                self.register_agent(agent_id)
            m = Message.model_construct(
                id=self._make_id(base + j), sender_id=str(sender_id), receiver_id=agent_id, content=content,
                type=mtype, priority=mprio, metadata=dict(msg_metadata), token_count=token_count,
                timestamp=now, status=MessageStatus.DELIVERED, visible_to=[])
            self.messages[m.id] = m
            self.message_count += 1
            self.agent_inbox[agent_id].append(m.id)
            self._b_msgs.append(m)
            ids.append(m.id)
            staged = True
        if staged or members:
            self._stage(1, self._index(sender_id), g, int(MessagePriority(priority)),
                        _TYPE_CODE[MessageType(message_type)] | flags, payload, now)
            self._next_seq += len(members)
        if ids:
            self._after_send()
        return ids

    # ------------------------------------------------------------------ receive
    def receive_messages(self, agent_id: str, max_messages: int = 100, timeout: float = 1.0) -> List[Message]:
        """Pop up to `max_messages` pending messages for `agent_id` in stream order (or priority
        order with GpuConfig.priority_dequeue).  Never blocks: `timeout` is accepted for
        compatibility (the reference spins up to `timeout` when fewer are available, M:553-556)."""
        if agent_id not in self.registered_agents:
            logger.warning(f"Agent {agent_id} not registered, registering now")
            self.register_agent(agent_id)
        self.flush()
        return self._receive_local(agent_id, max_messages)

    def _receive_local(self, agent_id: str, max_messages: int) -> List[Message]:
        if max_messages <= 0:
            return []
        flags = RECV_PRIORITY if self.gpu_config.priority_dequeue else 0
        out: List[Message] = []
        a = self._index(agent_id)
        remaining = max_messages
        while remaining > 0:
            hdr, pay, k = self.shard.receive_one(a, remaining, flags)   # single-launch latency path of the C ABI
            if len(hdr) == 0:
                break
            out.extend(self._decode(hdr, pay, agent_id))
            remaining -= len(hdr)
            if len(hdr) < k:
                break
        return out

    def _pre_read(self) -> None:
        """Make buffered sends visible before a read (a sharded front-end reads only what earlier collective
        flushes delivered, so it overrides this with a no-op)."""
        self.flush()

    def peek_messages(self, agent_id: str, max_messages: int = 100) -> List[Message]:
        """What `receive_messages` would return next, without consuming it (device PEEK receive)."""
        self._pre_read()
        flags = RECV_PEEK | (RECV_PRIORITY if self.gpu_config.priority_dequeue else 0)
        hdr, pay, _ = self.shard.receive_one(self._index(agent_id), max_messages, flags)
        return self._decode(hdr, pay, agent_id, record=False, status=MessageStatus.DELIVERED)

    def pending_snapshot(self, per_agent: int = 1024) -> Dict[str, List[Message]]:
        """Every message still queued on the device, per agent, in delivery order (non-destructive).
        This is the D2H ring dump behind `save_message_history(include_device=True)`."""
        self._pre_read()
        n = len(self._agent_name)
        out: Dict[str, List[Message]] = {}
        if n == 0:
            return out
        g = self.gpu_config
        flags = RECV_PEEK | (RECV_PRIORITY if g.priority_dequeue else 0)
        # A receive call holds min(max_recv_records, max_recv_payload / padded payload) records and leaves out whole
        # agents beyond that; a PEEK consumes nothing, so a fixed chunking would silently skip them.  Chunks are packed
        # from the agents' pending counts instead (one device query), so that every chunk fits.
        max_recv_payload = g.max_recv_payload or g.max_recv_records * 256
        cap = max(1, min(g.max_recv_records, max_recv_payload // pad32(g.max_payload_bytes)))
        pending = np.minimum(self.shard.agent_loads(None, n)["pending"][:n].astype(np.int64), per_agent)
        busy = np.nonzero(pending)[0]
        b = 0
        while b < len(busy):
            e, room = b, cap
            while e < len(busy) and (pending[busy[e]] <= room or e == b):
                room -= int(pending[busy[e]]); e += 1
            part = busy[b:e].astype(np.uint32)
            if e == b + 1 and pending[busy[b]] > cap:
                logger.warning(f"pending_snapshot: agent {self._agent_name[int(busy[b])]} has more pending messages than one "
                               f"receive call holds ({cap}); the snapshot keeps the first {cap}")
            counts, hdr, pay = self.shard.receive_batch(part, min(per_agent, cap), flags)
            offs = _native.payload_offsets(hdr)
            pos = 0
            for a, c in zip(part, counts):
                c = int(c)
                if c:
                    name = self._agent_name[int(a)]
                    sub_pay = pay[int(offs[pos]):] if pos < len(offs) else pay[:0]
                    out[name] = self._decode(hdr[pos:pos + c], sub_pay, name, record=False, status=MessageStatus.DELIVERED)
                pos += c
            b = e
        return out

    def _decode(self, hdr: np.ndarray, pay: np.ndarray, agent_id: str, record: bool = True,
                status: MessageStatus = MessageStatus.READ) -> List[Message]:
        msgs: List[Message] = []
        n = len(hdr)
        if n == 0:
            return msgs
        offs = _native.payload_offsets(hdr).tolist()
        raw = pay.tobytes()
        # column-wise conversion once: per-record structured-array field access dominates otherwise
        seqs, lens, types, prios = hdr["seq"].tolist(), hdr["len"].tolist(), hdr["type"].tolist(), hdr["prio"].tolist()
        senders, receivers, groups, stamps = (hdr["sender"].tolist(), hdr["receiver"].tolist(), hdr["group"].tolist(),
                                              hdr["timestamp"].tolist())
        names = self._agent_name
        for k in range(n):
            blob = raw[offs[k]: offs[k] + lens[k]]
            t = types[k]
            extras: Dict[str, Any] = {}
            if t & TYPEF_EXTRAS:
                clen = int.from_bytes(blob[:4], "little")
                body, extras = blob[4:4 + clen], json.loads(blob[4 + clen:].decode("utf-8"))
            else:
                body = blob
            content: Any = json.loads(body.decode("utf-8")) if t & TYPEF_JSON else body.decode("utf-8")
            metadata = dict(extras.get("m", {})) if extras else {}
            grp = groups[k]
            if grp != NO_GROUP:
                metadata["group"] = self._group_name[grp]
            receiver = receivers[k]
            m = Message.model_construct(
                id=self._make_id(seqs[k]), sender_id=names[senders[k]],
                receiver_id=None if receiver == _native.NO_RECEIVER else names[receiver],
                content=content, type=_TYPE_BY_CODE[t & TYPE_MASK], priority=MessagePriority(prios[k]),
                timestamp=stamps[k], status=status, metadata=metadata,
                token_count=extras.get("t") if extras else None, visible_to=self._decode_visibility(extras) if extras else [])
            if record:
                self.messages[m.id] = m                    # M:587-588
            msgs.append(m)
        return msgs

    # ------------------------------------------------------------------ bulk (index-level) API
    def send_batch(self, sender_idx, receiver_idx, prio, typ, lens, payload_off, payload, ts=None) -> int:
        self.flush()
        base = self.shard.send_batch(sender_idx, receiver_idx, prio, typ, lens, payload_off, payload, ts)
        self._next_seq = base + len(np.atleast_1d(sender_idx)); self._b_first_seq = self._next_seq
        return base

    def send_to_group_batch(self, sender_idx, group_idx, prio, typ, lens, payload_off, payload, ts=None) -> int:
        self.flush()
        base = self.shard.send_group_batch(sender_idx, group_idx, prio, typ, lens, payload_off, payload, ts)
        self._next_seq = self.shard.stats()["next_seq"]; self._b_first_seq = self._next_seq
        return base

    def receive_batch(self, agent_idx=None, max_messages: int = 100, priority: Optional[bool] = None, **kw):
        self.flush()
        pr = self.gpu_config.priority_dequeue if priority is None else priority
        return self.shard.receive_batch(agent_idx, max_messages, RECV_PRIORITY if pr else 0, **kw)

    # ------------------------------------------------------------------ LLM backends (M:1281-1325 + real balancer)
    def set_llm_load_balancing(self, enabled: bool = True) -> None:
        self.llm_load_balancing = enabled

    def assign_llm_backend(self, agent_id: str, backend_id: str) -> None:
        self.metadata.setdefault("llm_backends", {})[agent_id] = backend_id
        b = self._backend_idx.get(backend_id)
        if b is not None and agent_id in self._agent_idx:      # device copy of the sticky map: backs queue-fed loads
            self.shard.assign_agent_backends([self._agent_idx[agent_id]], [b])

    def get_llm_backend(self, agent_id: str) -> Optional[str]:
        return self.metadata.get("llm_backends", {}).get(agent_id)

    def register_llm_backends(self, backend_ids: List[str], weights: Optional[List[int]] = None,
                              loads: Optional[List[int]] = None) -> None:
        """Declare the backend pool the balancer picks from (weight = relative capacity)."""
        self._backend_name = list(backend_ids)
        self._backend_idx = {b: i for i, b in enumerate(self._backend_name)}
        self.shard.set_backends(np.asarray(weights if weights is not None else [1] * len(backend_ids), np.uint32),
                                None if loads is None else np.asarray(loads, np.uint64))

    def select_llm_backends(self, n_requests: int, cost=None, mode: str = "least_load", seed: int = 0) -> np.ndarray:
        """Batched pick: backend index per request (see include/swarmdb_b200.h for the definition)."""
        return self.shard.select_backends(n_requests, cost, 0 if mode == "least_load" else 1, seed)

    def select_llm_backend(self, agent_id: Optional[str] = None, cost: int = 1) -> Optional[str]:
        """Sticky assignment if the agent has one (M:1303-1307), else a balanced pick when
        load balancing is enabled; the pick is remembered as the agent's assignment."""
        if agent_id is not None:
            b = self.get_llm_backend(agent_id)
            if b is not None:
                return b
        if not self.llm_load_balancing or not self._backend_name:
            return None
        b = self._backend_name[int(self.shard.select_backends(1, np.array([cost], np.uint32), 0, 0)[0])]
        if agent_id is not None:
            self.assign_llm_backend(agent_id, b)
        return b

    def release_llm_backend(self, backend_id: str, cost: int = 1) -> None:
        self.shard.release_backends([self._backend_idx[backend_id]], [cost])

    def refresh_llm_backend_loads(self) -> Dict[str, int]:
        """Set every backend's load to the backlog of the agents assigned to it - the reference's only load signal is
        get_agent_load (M:1049-1094); here it is one pass over the ring headers on the device."""
        self._pre_read()
        self.shard.backend_loads_from_queues()
        return self.llm_backend_loads()

    def llm_backend_loads(self) -> Dict[str, int]:
        return {b: int(l) for b, l in zip(self._backend_name, self.shard.backend_loads())}

    # ------------------------------------------------------------------ host-side bookkeeping (M:603-808, 973-1206)
    def get_message(self, message_id: str) -> Optional[Message]:
        return self.messages.get(message_id)

    def get_agent_messages(self, agent_id: str, status: Optional[MessageStatus] = None, limit: int = 100,
                           skip: int = 0) -> List[Message]:
        picked: List[Message] = []
        for pos, mid in enumerate(reversed(self.agent_inbox.get(agent_id, []))):
            if pos < skip:
                continue
            if len(picked) >= limit:
                break
            m = self.messages.get(mid)
            if m is not None and (status is None or m.status == status):
                picked.append(m)
        return picked

    def mark_message_as_processed(self, message_id: str) -> bool:
        m = self.messages.get(message_id)
        if m is None:
            return False
        m.status = MessageStatus.PROCESSED
        return True

    def query_messages(self, sender_id: Optional[str] = None, receiver_id: Optional[str] = None,
                       message_type: Optional[MessageType] = None, status: Optional[MessageStatus] = None,
                       after_timestamp: Optional[float] = None, before_timestamp: Optional[float] = None,
                       limit: int = 100) -> List[Message]:
        def keep(m: Message) -> bool:
            return ((sender_id is None or m.sender_id == sender_id)
                    and (receiver_id is None or m.receiver_id == receiver_id)
                    and (message_type is None or m.type == message_type)
                    and (status is None or m.status == status)
                    and (after_timestamp is None or m.timestamp > after_timestamp)
                    and (before_timestamp is None or m.timestamp < before_timestamp))
        hits: List[Message] = []
        for m in reversed(list(self.messages.values())):
            if len(hits) >= limit:
                break
            if keep(m):
                hits.append(m)
        return hits

    def search_messages(self, keyword: str, case_sensitive: bool = False, limit: int = 100) -> List[Message]:
        needle = keyword if case_sensitive else keyword.lower()
        hits: List[Message] = []
        for m in reversed(list(self.messages.values())):
            if len(hits) >= limit:
                break
            text = json.dumps(m.content) if isinstance(m.content, (dict, list)) else str(m.content)
            if needle in (text if case_sensitive else text.lower()):
                hits.append(m)
        return hits

    def get_conversation(self, agent_id_1: str, agent_id_2: str, limit: int = 100) -> List[Message]:
        half = limit // 2
        return (self.query_messages(sender_id=agent_id_1, receiver_id=agent_id_2, limit=half)
                + self.query_messages(sender_id=agent_id_2, receiver_id=agent_id_1, limit=half))

    def get_unread_message_count(self, agent_id: str) -> int:
        return sum(1 for mid in self.agent_inbox.get(agent_id, [])
                   if mid in self.messages and self.messages[mid].status == MessageStatus.DELIVERED)

    def get_agent_load(self, agent_id: str) -> Dict[str, Any]:
        if agent_id not in self.registered_agents:
            return {"registered": False, "message_count": 0, "inbox_size": 0, "unread_count": 0, "processing_rate": 0}
        horizon = time.time() - 60
        involved = recent = 0
        for m in self.messages.values():
            if m.receiver_id == agent_id or m.sender_id == agent_id:
                involved += 1
            if m.receiver_id == agent_id and m.timestamp > horizon:
                recent += 1
        return {"registered": True, "message_count": involved, "inbox_size": len(self.agent_inbox.get(agent_id, [])),
                "unread_count": self.get_unread_message_count(agent_id), "processing_rate": recent / 60}

    def get_agent_queue_load(self, agent_id: str) -> Dict[str, Any]:
        """get_agent_load's inbox_size / unread_count (M:1076-1080) answered by the device queue itself: works for
        traffic that never had a host-side record (bulk index-level sends, messages ingested by other ranks)."""
        if agent_id not in self._agent_idx:
            return {"registered": False, "inbox_size": 0, "unread_count": 0, "unread_by_priority": [0, 0, 0, 0], "unread_bytes": 0}
        self._pre_read()
        ld = self.shard.agent_loads([self._agent_idx[agent_id]])[0]
        return {"registered": agent_id in self.registered_agents, "inbox_size": int(ld["received"]),
                "unread_count": int(ld["pending"]), "unread_by_priority": [int(x) for x in ld["pending_by_prio"]],
                "unread_bytes": int(ld["pending_granules"]) * 32}

    def get_stats(self) -> Dict[str, Any]:
        by_type = {t.value: 0 for t in MessageType}
        by_status = {s.value: 0 for s in MessageStatus}
        sent: Dict[str, int] = {}
        recv: Dict[str, int] = {}
        for m in self.messages.values():
            by_type[m.type.value] += 1
            by_status[m.status.value] += 1
            sent[m.sender_id] = sent.get(m.sender_id, 0) + 1
            if m.receiver_id is not None:
                recv[m.receiver_id] = recv.get(m.receiver_id, 0) + 1
        by_agent = {a: {"sent": sent.get(a, 0), "received": recv.get(a, 0), "total": sent.get(a, 0) + recv.get(a, 0)}
                    for a in self.registered_agents}
        return {"total_messages": self.message_count, "active_agents": len(self.registered_agents),
                "messages_by_type": by_type, "messages_by_status": by_status, "messages_by_agent": by_agent,
                "last_save_time": self.last_save_time, "device": self.shard.stats(),
                "queue": self.shard.queue_stats()}            # pending totals / priority histogram / deepest inbox, computed on the device

    def resend_failed_messages(self) -> List[str]:
        resent: List[str] = []
        for m in [x for x in self.messages.values() if x.status == MessageStatus.FAILED]:
            new_id = self.send_message(sender_id=m.sender_id, content=m.content, receiver_id=m.receiver_id,
                                       message_type=m.type, priority=m.priority, metadata=m.metadata,
                                       visible_to=m.visible_to)
            self.messages[new_id].metadata["resent_from"] = m.id
            resent.append(new_id)
        return resent

    def delete_message(self, message_id: str) -> bool:
        if self.messages.pop(message_id, None) is None:
            return False
        for inbox in self.agent_inbox.values():
            if message_id in inbox:
                inbox.remove(message_id)
        return True

    def flush_old_messages(self, older_than: Optional[float] = None) -> int:
        cutoff = older_than if older_than is not None else time.time() - 7 * 24 * 3600
        old = [mid for mid, m in self.messages.items() if m.timestamp < cutoff]
        if old:
            path = self.save_dir / "archives" / f"archive_{int(time.time())}.json"
            path.parent.mkdir(parents=True, exist_ok=True)
            path.write_text(json.dumps({mid: self.messages[mid].to_dict() for mid in old}))
            for mid in old:
                self.delete_message(mid)
        return len(old)

    # ------------------------------------------------------------------ history (schema of M:878-884)
    def _history(self, include_device: bool = False) -> Dict[str, Any]:
        messages = {mid: m.to_dict() for mid, m in self.messages.items()}
        inbox = self.agent_inbox
        if include_device:
            # messages that only exist on the device (bulk index-level sends have no host record): add the
            # pending ones to the same schema (M:878-884) so a history file is complete
            inbox = {a: list(v) for a, v in self.agent_inbox.items()}
            for agent, pend in self.pending_snapshot().items():
                box = inbox.setdefault(agent, [])
                have = set(box)
                for m in pend:
                    if m.id not in messages:
                        messages[m.id] = m.to_dict()
                    if m.id not in have:
                        box.append(m.id); have.add(m.id)
        return {"messages": messages, "agent_inbox": inbox, "registered_agents": list(self.registered_agents),
                "timestamp": time.time(), "message_count": max(self.message_count, len(messages))}

    def save_message_history(self, filename: Optional[str] = None, include_device: bool = False) -> None:
        if not filename:
            stamp = datetime.datetime.now().strftime("%Y%m%d_%H%M%S")
            filename = f"message_history_{stamp}_{self.message_count}.json"
        try:
            with open(self.save_dir / filename, "w") as f:
                json.dump(self._history(include_device), f, indent=2)
            self.last_save_time = time.time()
        except Exception as e:
            logger.error(f"Failed to save message history: {e}")

    def load_message_history(self, filepath: Union[str, Path]) -> None:
        path = Path(filepath)
        if not path.exists():
            logger.error(f"Message history file {path} does not exist")
            return
        try:
            history = json.loads(path.read_text())
            self.messages = {mid: Message.from_dict(d) for mid, d in history["messages"].items()}
            self.agent_inbox = history["agent_inbox"]
            for a in history["registered_agents"]:
                self.register_agent(a)
            self.message_count = history["message_count"]
        except Exception as e:
            logger.error(f"Failed to load message history: {e}")

    def export_as_yaml(self, filepath: Union[str, Path]) -> None:
        try:
            import yaml
            with open(Path(filepath), "w") as f:
                yaml.dump(self._history(), f, sort_keys=False)
        except Exception as e:
            logger.error(f"Failed to export message history as YAML: {e}")

    def auto_scale_partitions(self) -> None:
        """Kafka partition growth (M:1327-1365) has no analogue: shards are the GPUs of the box."""
        return None

    # ------------------------------------------------------------------ lifecycle (M:1367-1394)
    def close(self) -> None:
        if self._closed:
            return
        try:
            self.flush()
            if self.auto_save:
                self.save_message_history()
        finally:
            self.shard.close()
            self._closed = True

    def __enter__(self) -> "SwarmsDB":
        return self

    def __exit__(self, exc_type, exc_val, exc_tb) -> None:
        self.close()
