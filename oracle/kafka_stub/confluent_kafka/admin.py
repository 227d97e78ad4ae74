"""Admin half of the in-memory `confluent_kafka` stand-in (TEST INFRASTRUCTURE ONLY).

Emulates the call sites AdminClient M:202-204, .list_topics M:241 / M:1332,
.create_topics M:277, .create_partitions M:1349, NewTopic M:249 / M:263,
NewPartitions M:1345 of the reference (/root/reference/swarmdb/" main.py").
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

from . import KafkaException, broker

__all__ = ["AdminClient", "NewTopic", "NewPartitions"]


class NewTopic:
    def __init__(self, topic: str, num_partitions: int = 1, replication_factor: int = 1,
                 config: Optional[Dict[str, str]] = None) -> None:
        self.topic = topic
        self.num_partitions = num_partitions
        self.replication_factor = replication_factor
        self.config = config or {}


class NewPartitions:
    def __init__(self, new_total_count: int) -> None:
        self.new_total_count = new_total_count


class _Future:
    def __init__(self, exc: Optional[Exception] = None) -> None:
        self._exc = exc

    def result(self, timeout: Optional[float] = None) -> None:
        if self._exc is not None:
            raise self._exc
        return None


class _TopicMetadata:
    def __init__(self, name: str, n: int) -> None:
        self.topic = name
        self.partitions = {i: object() for i in range(n)}


class _ClusterMetadata:
    def __init__(self) -> None:
        self.topics = {name: _TopicMetadata(name, len(parts))
                       for name, parts in broker().topics.items()}


class AdminClient:
    def __init__(self, conf: Dict[str, Any]) -> None:
        self.conf = dict(conf)

    def list_topics(self, topic: Optional[str] = None, timeout: float = -1) -> _ClusterMetadata:
        return _ClusterMetadata()

    def create_topics(self, new_topics: List[NewTopic], **_kw) -> Dict[str, _Future]:
        out = {}
        for t in new_topics:
            if t.topic in broker().topics:
                out[t.topic] = _Future(KafkaException(f"Topic '{t.topic}' already exists"))
            else:
                broker().ensure_topic(t.topic, t.num_partitions)
                out[t.topic] = _Future()
        return out

    def create_partitions(self, new_parts, **_kw) -> Dict[str, _Future]:
        out = {}
        items = new_parts.items() if isinstance(new_parts, dict) else [(p.topic, p) for p in new_parts]
        for topic, np_ in items:
            if topic not in broker().topics:
                out[topic] = _Future(KafkaException(f"unknown topic {topic}"))
            else:
                broker().ensure_topic(topic, np_.new_total_count)
                out[topic] = _Future()
        return out
