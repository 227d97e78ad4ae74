"""swarmdb_b200 - B200-native agent message queue + LLM-backend balancer.

Drop-in for the hot path of The-Swarm-Corporation/SwarmDB: the same Python surface
(`SwarmsDB`, `KafkaConfig`, `Message`, `MessageType`, `MessagePriority`, `MessageStatus`,
reference swarmdb/" main.py" and the import at api.py:29-36), with the Kafka
producer/consumer layer replaced by GPU-resident per-agent rings driven through the C ABI in
include/swarmdb_b200.h.  No CPU fallback.
"""
from .core import (GpuConfig, KafkaConfig, Message, MessagePriority, MessageStatus, MessageType,  # noqa: F401
                   RingOverflow, SwarmsDB)

__all__ = ["SwarmsDB", "KafkaConfig", "GpuConfig", "Message", "MessageType", "MessagePriority", "MessageStatus",
           "RingOverflow"]
