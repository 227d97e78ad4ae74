"""In-memory stand-in for the `confluent_kafka` client API (TEST INFRASTRUCTURE ONLY).

The reference's transport is the third-party `confluent_kafka` package (C librdkafka +
a real Kafka broker), unpinned in /root/reference/requirements.txt:1 and absent from this
image.  This stub emulates exactly the partition-log semantics the reference relies on, so
that the reference's own `SwarmsDB` class (swarmdb/" main.py") can be executed unmodified
to pin the oracle.  Call sites emulated (reference file:line):

  Producer(conf)            M:192-199     .produce  M:476-482, M:509-513
  .poll(0)                  M:484         .flush()  M:1386
  Consumer(conf)            M:334-343     .subscribe M:344   .poll M:557-559   .close M:367
  msg.key()/value()/error() M:379, M:565-575
  KafkaError._PARTITION_EOF M:566         KafkaException M:284

Semantics (Kafka protocol facts, not reference code):
  * a topic is a list of partitions, each an append-only log of (key, value);
  * produce with an explicit partition appends to that log; the delivery callback runs
    inside the next Producer.poll()/flush() on the caller's thread;
  * consumer offsets are kept per (group.id, topic, partition), start at 0 for
    auto.offset.reset == "earliest", survive Consumer.close() (committed offsets);
  * Consumer.poll returns the next unread record; when the group has caught up with every
    partition it returns a `_PARTITION_EOF` error event (librdkafka's
    `enable.partition.eof` behaviour), which the reference's drain loop handles by
    breaking (M:565-568) instead of spinning until `timeout` (M:553-563).  The returned
    message lists are identical either way; only the wait differs.
  For topics with more than one partition poll round-robins over partitions that have
  unread records - a legal interleave that preserves per-partition order.

Nothing in the shipped product imports this package.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Tuple

__all__ = ["Producer", "Consumer", "KafkaError", "KafkaException", "Message", "reset_broker"]


class _Broker:
    def __init__(self) -> None:
        self.topics: Dict[str, List[List[Tuple[Optional[bytes], bytes]]]] = {}
        # (group, topic, partition) -> next offset
        self.offsets: Dict[Tuple[str, str, int], int] = {}
        self.fail_next_produce: Optional[str] = None  # fault injection for tests

    def ensure_topic(self, name: str, num_partitions: int) -> None:
        parts = self.topics.setdefault(name, [])
        while len(parts) < num_partitions:
            parts.append([])


_BROKER = _Broker()


def reset_broker() -> None:
    """Drop all topics and offsets (a fresh cluster)."""
    global _BROKER
    _BROKER = _Broker()


def broker() -> _Broker:
    return _BROKER


class KafkaException(Exception):
    pass


class KafkaError:
    _PARTITION_EOF = -191
    _UNKNOWN_PARTITION = -190

    def __init__(self, code: int, reason: str = "") -> None:
        self._code = code
        self._reason = reason

    def code(self) -> int:
        return self._code

    def __str__(self) -> str:  # pragma: no cover - cosmetic
        return f"KafkaError{{code={self._code},str={self._reason!r}}}"


class Message:
    def __init__(self, topic: str, partition: int, offset: int,
                 key: Optional[bytes], value: Optional[bytes], error: Optional[KafkaError] = None):
        self._topic, self._partition, self._offset = topic, partition, offset
        self._key, self._value, self._error = key, value, error

    def key(self): return self._key
    def value(self): return self._value
    def error(self): return self._error
    def topic(self): return self._topic
    def partition(self): return self._partition
    def offset(self): return self._offset


def _as_bytes(x: Any) -> Optional[bytes]:
    if x is None:
        return None
    if isinstance(x, bytes):
        return x
    return str(x).encode("utf-8")


class Producer:
    def __init__(self, conf: Dict[str, Any]) -> None:
        self.conf = dict(conf)
        self._pending: List[Tuple[Callable, Optional[KafkaError], Message]] = []

    def produce(self, topic: str, value: Any = None, key: Any = None,
                partition: int = -1, callback: Optional[Callable] = None, **_kw) -> None:
        b = broker()
        if b.fail_next_produce is not None:
            reason, b.fail_next_produce = b.fail_next_produce, None
            raise KafkaException(reason)
        if topic not in b.topics:
            b.ensure_topic(topic, 1)           # auto.create.topics.enable default
        parts = b.topics[topic]
        if partition is None or partition < 0:
            partition = 0                      # keyless default partitioner, 1-partition topics only
        if partition >= len(parts):
            raise KafkaException(f"unknown partition {partition} for topic {topic}")
        log = parts[partition]
        msg = Message(topic, partition, len(log), _as_bytes(key), _as_bytes(value))
        log.append((msg.key(), msg.value()))
        if callback is not None:
            self._pending.append((callback, None, msg))

    def poll(self, timeout: float = 0) -> int:
        n = 0
        while self._pending:
            cb, err, msg = self._pending.pop(0)
            cb(err, msg)
            n += 1
        return n

    def flush(self, timeout: float = -1) -> int:
        self.poll(0)
        return 0


class Consumer:
    def __init__(self, conf: Dict[str, Any]) -> None:
        self.conf = dict(conf)
        self.group = conf["group.id"]
        self.topics: List[str] = []
        self._rr = 0
        self._closed = False

    def subscribe(self, topics: List[str]) -> None:
        self.topics = list(topics)

    def poll(self, timeout: Optional[float] = None) -> Optional[Message]:
        if self._closed:
            raise RuntimeError("Consumer closed")
        b = broker()
        for topic in self.topics:
            parts = b.topics.get(topic, [])
            n = len(parts)
            for k in range(n):
                p = (self._rr + k) % n
                off = b.offsets.get((self.group, topic, p), 0)
                if off < len(parts[p]):
                    key, value = parts[p][off]
                    b.offsets[(self.group, topic, p)] = off + 1
                    self._rr = (p + 1) % n
                    return Message(topic, p, off, key, value)
        if self.topics:
            return Message(self.topics[0], 0, -1, None, None,
                           KafkaError(KafkaError._PARTITION_EOF, "caught up"))
        return None

    def close(self) -> None:
        self._closed = True
