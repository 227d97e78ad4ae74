"""Drop-in import path: `from swarmdb import SwarmsDB, KafkaConfig, Message, ...` (the import the
reference's api.py:29-36 performs) resolves to the B200-native implementation."""
from swarmdb_b200 import (GpuConfig, KafkaConfig, Message, MessagePriority, MessageStatus,  # noqa: F401
                          MessageType, SwarmsDB)

__all__ = ["SwarmsDB", "KafkaConfig", "GpuConfig", "Message", "MessageType", "MessagePriority", "MessageStatus"]
