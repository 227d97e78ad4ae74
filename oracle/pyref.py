"""Pure-Python restatement of the reference hot path (ORACLE - TEST INFRASTRUCTURE ONLY).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline/reference legs may
import this module; the shipped product (`swarmdb_b200/`) never does.

Restates, rule by rule, what /root/reference/swarmdb/" main.py" (tag `M:`) does on the path
register -> send / send_to_group / broadcast -> receive, over a Kafka topic with ONE
partition (the linearisation the build adopts, SURVEY.md Appendix A rule 7):

  * the topic is one append-only log; every agent has its own consumer group, i.e. its own
    read offset into the whole log (M:334-345), which survives deregistration (rule 11);
  * `receive_messages` walks the log from the agent's offset and keeps a record iff
    (receiver == agent or receiver is None) and (agent in visible_to or not visible_to)
    (M:579-585), stopping after `max_messages` matches (M:553-556).

PINNING: checked against the golden fixtures in tests/golden/*.json, which were produced by
the unmodified reference class (tests/golden/make_golden.py) - see tests/test_oracle_golden.py.
"""
from __future__ import annotations

import json
import time
import uuid
from enum import Enum
from typing import Any, Callable, Dict, List, Optional, Set


class MessageType(str, Enum):          # M:23-32
    CHAT = "chat"
    COMMAND = "command"
    FUNCTION_CALL = "function_call"
    FUNCTION_RESULT = "function_result"
    SYSTEM = "system"
    ERROR = "error"
    STATUS = "status"


class MessagePriority(int, Enum):      # M:35-41
    LOW = 0
    NORMAL = 1
    HIGH = 2
    CRITICAL = 3


class MessageStatus(str, Enum):        # M:44-51
    PENDING = "pending"
    DELIVERED = "delivered"
    READ = "read"
    PROCESSED = "processed"
    FAILED = "failed"


class Message:                          # fields and defaults of M:54-82
    __slots__ = ("id", "sender_id", "receiver_id", "content", "type", "priority", "timestamp",
                 "status", "metadata", "token_count", "visible_to")

    def __init__(self, **kw: Any) -> None:
        for k in self.__slots__:
            setattr(self, k, kw[k])

    def to_wire(self) -> str:           # M:466 (json.dumps(message.to_dict()))
        return json.dumps({
            "id": self.id, "sender_id": self.sender_id, "receiver_id": self.receiver_id,
            "content": self.content, "type": self.type.value, "priority": self.priority.value,
            "timestamp": self.timestamp, "status": self.status.value, "metadata": self.metadata,
            "token_count": self.token_count, "visible_to": self.visible_to})

    @classmethod
    def from_wire(cls, s: str) -> "Message":   # M:575-576, M:100-111
        d = json.loads(s)
        d["type"] = MessageType(d["type"])
        d["priority"] = MessagePriority(d["priority"])
        d["status"] = MessageStatus(d["status"])
        return cls(**d)


class OracleSwarmsDB:
    """Same method surface as the reference `SwarmsDB` for the hot path."""

    def __init__(self, id_factory: Optional[Callable[[], str]] = None,
                 clock: Callable[[], float] = time.time) -> None:
        self._log: List[str] = []                       # the 1-partition topic (wire records)
        self._offset: Dict[str, int] = {}               # consumer-group offsets, keyed by agent
        self.registered_agents: Set[str] = set()
        self.agent_inbox: Dict[str, List[str]] = {}
        self.messages: Dict[str, Message] = {}
        self.message_count = 0
        self.metadata: Dict[str, Any] = {}
        self._id_factory = id_factory or (lambda: str(uuid.uuid4()))
        self._clock = clock

    # M:314-349
    def register_agent(self, agent_id: str) -> None:
        if agent_id in self.registered_agents:
            return
        self.registered_agents.add(agent_id)
        self.agent_inbox.setdefault(agent_id, [])
        self._offset.setdefault(agent_id, 0)            # auto.offset.reset = earliest (M:338)

    # M:351-372: registry entry dropped, inbox and committed offset kept
    def deregister_agent(self, agent_id: str) -> None:
        if agent_id not in self.registered_agents:
            return
        self.registered_agents.remove(agent_id)

    # M:393-519
    def send_message(self, sender_id: str, content: Any, receiver_id: Optional[str] = None,
                     message_type: MessageType = MessageType.CHAT,
                     priority: MessagePriority = MessagePriority.NORMAL,
                     metadata: Optional[Dict[str, Any]] = None,
                     visible_to: Optional[List[str]] = None) -> str:
        if sender_id not in self.registered_agents:                      # M:419-420
            self.register_agent(sender_id)
        if receiver_id is not None and receiver_id not in self.registered_agents:   # M:423-427
            self.register_agent(receiver_id)
        msg = Message(id=self._id_factory(), sender_id=sender_id, receiver_id=receiver_id,
                      content=content, type=MessageType(message_type), priority=MessagePriority(priority),
                      timestamp=self._clock(), status=MessageStatus.PENDING,
                      metadata=dict(metadata or {}), token_count=None, visible_to=list(visible_to or []))
        if receiver_id is None and not msg.visible_to:                   # M:449-450
            msg.visible_to = list(self.registered_agents)
        self.messages[msg.id] = msg                                      # M:453-454
        self.message_count += 1
        if receiver_id is not None:                                      # M:457-463
            if receiver_id in self.agent_inbox:
                self.agent_inbox[receiver_id].append(msg.id)
        else:
            for a in self.registered_agents:
                self.agent_inbox[a].append(msg.id)
        self._log.append(msg.to_wire())                                  # M:466-482 (status PENDING on the wire)
        msg.status = MessageStatus.DELIVERED                             # M:484 -> M:387-391
        return msg.id

    # M:521-601
    def receive_messages(self, agent_id: str, max_messages: int = 100, timeout: float = 1.0) -> List[Message]:
        if agent_id not in self.registered_agents:                       # M:538-542
            self.register_agent(agent_id)
        out: List[Message] = []
        off = self._offset[agent_id]
        while len(out) < max_messages and off < len(self._log):          # M:553-563
            m = Message.from_wire(self._log[off])
            off += 1
            if (m.receiver_id == agent_id or m.receiver_id is None) and \
                    (agent_id in m.visible_to or len(m.visible_to) == 0):      # M:579-585
                m.status = MessageStatus.READ                            # M:587-588
                self.messages[m.id] = m
                out.append(m)
        self._offset[agent_id] = off
        return out

    # M:810-850
    def broadcast_message(self, sender_id: str, content: Any,
                          message_type: MessageType = MessageType.CHAT,
                          priority: MessagePriority = MessagePriority.NORMAL,
                          metadata: Optional[Dict[str, Any]] = None,
                          exclude_agents: Optional[List[str]] = None) -> str:
        excl = set(exclude_agents or [])
        vis = [a for a in self.registered_agents if a != sender_id and a not in excl]
        return self.send_message(sender_id, content, None, message_type, priority, metadata, vis)

    # M:1208-1227 (stores the caller's list object, overwrite)
    def add_agent_group(self, group_name: str, agent_ids: List[str]) -> None:
        self.metadata.setdefault("agent_groups", {})[group_name] = agent_ids

    create_group = add_agent_group

    # M:1229-1279
    def send_to_group(self, sender_id: str, group_name: str, content: Any,
                      message_type: MessageType = MessageType.CHAT,
                      priority: MessagePriority = MessagePriority.NORMAL,
                      metadata: Optional[Dict[str, Any]] = None) -> List[str]:
        groups = self.metadata.get("agent_groups", {})
        if group_name not in groups:                                     # M:1255-1257
            return []
        md = metadata or {}
        md["group"] = group_name                                         # M:1263-1264 (mutates caller's dict)
        ids = []
        for a in groups[group_name]:                                     # M:1267-1277
            if a != sender_id:
                ids.append(self.send_message(sender_id, content, a, message_type, priority, md))
        return ids

    # M:1281-1325
    def set_llm_load_balancing(self, enabled: bool = True) -> None:
        self.llm_load_balancing = enabled

    def assign_llm_backend(self, agent_id: str, backend_id: str) -> None:
        self.metadata.setdefault("llm_backends", {})[agent_id] = backend_id

    def get_llm_backend(self, agent_id: str) -> Optional[str]:
        return self.metadata.get("llm_backends", {}).get(agent_id)


def counter_ids(start: int = 1) -> Callable[[], str]:
    """Deterministic id factory equal to the reference under the uuid shim."""
    state = {"n": start - 1}

    def nxt() -> str:
        state["n"] += 1
        return str(uuid.UUID(int=state["n"]))
    return nxt
