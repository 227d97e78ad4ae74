/*
 * swarmdb_b200.h - C ABI of the B200-native agent message queue + backend balancer.
 *
 * This is the drop-in boundary for the hot path of The-Swarm-Corporation/SwarmDB
 * (reference class `SwarmsDB` in /root/reference/swarmdb/" main.py", tag `M:`).  The
 * reference has no FFI of its own: its `SwarmsDB` methods call the third-party
 * `confluent_kafka` client (librdkafka) directly.  Each entry point below replaces one
 * reference method body plus the Kafka client calls it makes; the file:line it replaces is
 * cited on every declaration.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - plain C: opaque handle, POD structs, pointers and sizes; no C++/torch types.
 *   - every function returns 0 (SDB_OK) or a negative sdb_status; sdb_last_error(h)
 *     returns a human-readable string for the last failure on that handle.
 *   - a handle is NOT thread-safe (the reference class has no locks either);
 *     one handle drives one GPU (one shard).  One process per GPU.
 *   - pointer arguments are caller-owned HOST buffers unless the name ends in `_dev`.
 *   - calls are stream-ordered on the handle's CUDA stream; functions that return
 *     data to the host synchronise that stream before returning, others may return early.
 *   - agents, groups and backends are dense uint32 indices; the string ids of the Python
 *     surface are mapped by the caller (swarmdb_b200/core.py keeps the dict, as the
 *     reference keeps `registered_agents`, M:233).
 *   - there is NO CPU fallback: without a CUDA device sdb_create fails with SDB_ECUDA.
 */
#ifndef SWARMDB_B200_H
#define SWARMDB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDB_ABI_VERSION 1u
#define SDB_GRANULE 32u                 /* arena allocation unit, bytes (one DRAM sector) */
#define SDB_NO_GROUP 0xFFFFFFFFu        /* header.group when the message was not a group send */
#define SDB_NO_RECEIVER 0xFFFFFFFFu     /* header.receiver for broadcast records (receiver_id=None, M:61) */

typedef struct sdb_ctx* sdb_handle;

typedef enum sdb_status {
  SDB_OK = 0,
  SDB_EINVAL = -1,         /* bad argument (index out of range, misaligned offset, ...) */
  SDB_ECUDA = -2,          /* CUDA runtime failure (incl. no device) */
  SDB_ENOMEM = -3,         /* host or device allocation failed */
  SDB_ERING_OVERFLOW = -4, /* some receiver's ring was full: those records were NOT enqueued (counted) */
  SDB_EARENA_FULL = -5,    /* message arena cannot hold the batch even after reclaiming consumed space */
  SDB_ECAPACITY = -6,      /* batch larger than the configured staging capacity */
  SDB_ENOTFOUND = -7,      /* unknown group / backend */
  SDB_EOUTPUT = -8         /* caller's output buffer too small */
} sdb_status;

/* Message vocabularies, M:23-51.  Device codes; names in swarmdb_b200/core.py. */
enum { SDB_PRIO_LOW = 0, SDB_PRIO_NORMAL = 1, SDB_PRIO_HIGH = 2, SDB_PRIO_CRITICAL = 3 };
enum { SDB_TYPE_CHAT = 0, SDB_TYPE_COMMAND, SDB_TYPE_FUNCTION_CALL, SDB_TYPE_FUNCTION_RESULT,
       SDB_TYPE_SYSTEM, SDB_TYPE_ERROR, SDB_TYPE_STATUS };
#define SDB_TYPE_MASK 0x07u
#define SDB_TYPEF_JSON 0x08u            /* payload content is JSON (dict/list content, M:75) */
#define SDB_TYPEF_EXTRAS 0x10u          /* payload = u32 content_len | content | JSON extras */

/* Replaces KafkaConfig (M:114-127) + SwarmsDB.__init__ kwargs (M:156-165) for the device side. */
typedef struct sdb_config {
  uint32_t struct_bytes;        /* sizeof(sdb_config), ABI check */
  int32_t  device;              /* CUDA ordinal */
  uint32_t shard_id;            /* this process's shard (== rank) */
  uint32_t num_shards;          /* replaces num_partitions (M:121): 1, 2, 4 or 8 */
  uint32_t max_agents;          /* size of the global agent index space */
  uint32_t ring_slots;          /* per-agent ring capacity in messages, power of two */
  uint64_t arena_bytes;         /* message arena, power of two */
  uint32_t max_payload_bytes;   /* largest payload accepted (<= 65504) */
  uint32_t max_groups;
  uint64_t member_pool_entries; /* total group-member slots (groups may be re-created) */
  uint32_t max_backends;
  uint32_t max_batch_sends;     /* staging capacity: sends per batch */
  uint64_t max_batch_payload;   /* staging capacity: payload bytes per batch */
  uint64_t max_recv_records;    /* receive output capacity, records per call */
  uint64_t max_recv_payload;    /* receive output capacity, payload bytes per call (0: records x 256) */
  uint64_t list_pool_entries;   /* per-batch broadcast recipient-list capacity (0: 2 x max_agents) */
  uint32_t fanout_variant;      /* 2 (recommended): warp-per-send, TMA-in / coalesced stores - imports of wire batches use the deep-prefetch form;
                                   3: deep-prefetch warp-per-send everywhere (payloads <= 480 B); 0: CTA-per-send; 1: TMA-in / TMA-out */
  uint32_t flags;               /* reserved, 0 */
} sdb_config;

/* One message as stored in the arena and returned by receive: 32-byte header followed by
 * the payload padded to SDB_GRANULE.  Field set = the device-relevant subset of `Message`
 * (M:54-82): id <-> seq, sender_id, receiver_id, type, priority, timestamp, metadata["group"]. */
typedef struct sdb_msg_header {
  uint64_t seq;        /* message identity; ids are derived from it (replaces uuid4, M:72) */
  double   timestamp;  /* M:78, carried, never compared */
  uint32_t sender;     /* agent index */
  uint32_t receiver;   /* agent index this copy was delivered to */
  uint32_t group;      /* group index (M:1264 metadata["group"]) or SDB_NO_GROUP */
  uint16_t len;        /* payload bytes */
  uint8_t  prio;       /* 0..3, M:35-41 */
  uint8_t  type;       /* SDB_TYPE_* | SDB_TYPEF_* */
} sdb_msg_header;

typedef struct sdb_stats {
  uint64_t next_seq;            /* messages created so far (M:454 message_count analogue) */
  uint64_t enqueued;            /* records appended to rings */
  uint64_t delivered;           /* records returned by receive */
  uint64_t ring_overflow;       /* records dropped because a ring was full (reported, never silent) */
  uint64_t skipped_sender;      /* group members skipped because member == sender (M:1268) */
  uint64_t arena_tail_bytes;    /* monotonic arena write position */
  uint64_t arena_floor_bytes;   /* everything below is reclaimed */
  uint64_t n_agents;            /* registered-index watermark */
  uint64_t kernel_launches;     /* CUDA kernels launched by this handle so far */
  uint64_t backend_picks;
} sdb_stats;

/* ---- lifecycle: SwarmsDB.__init__ (M:156-237), close (M:1367-1394) ---------------------- */
int sdb_abi_version(void);
int sdb_create(const sdb_config* cfg, sdb_handle* out);
int sdb_destroy(sdb_handle h);
/* Run on a caller-provided CUDA stream (e.g. torch's current stream) instead of the internal one. */
int sdb_set_stream(sdb_handle h, void* cuda_stream);
int sdb_sync(sdb_handle h);
const char* sdb_last_error(sdb_handle h);
int sdb_get_stats(sdb_handle h, sdb_stats* out);
/* Test hook: move the (empty) arena's write position, e.g. next to the 2^32-granule boundary where the
 * 32-bit ring handles wrap.  Only valid while no message is pending. */
int sdb_debug_set_arena_pos(sdb_handle h, uint64_t granules);
/* Move the sequence counter FORWARD to `next_seq` (no-op when it is already there or beyond).  A front-end that
 * hands out message ids at send time (ids are derived from sequence numbers, replacing uuid4 M:72) calls this when a
 * batch was refused, so the ids of the failed messages (M:501-519: marked FAILED, kept) are never issued twice. */
int sdb_advance_seq(sdb_handle h, uint64_t next_seq);

/* Per-kernel device timing (CUDA events on the handle's stream), used by bench.py for the
 * roofline: enable, run, then read accumulated milliseconds and launch counts per kernel class. */
enum { SDB_PK_P2P = 0, SDB_PK_FANOUT, SDB_PK_COMMIT, SDB_PK_RECV_COUNT, SDB_PK_RECV_SCAN, SDB_PK_RECV_SELECT,
       SDB_PK_RECV_GATHER, SDB_PK_ARENA_FLOOR, SDB_PK_PICK, SDB_PK_XSHARD, SDB_PK_INDEX, SDB_PK_XWAIT, SDB_PK_N = 16 };
int sdb_profile(sdb_handle h, int enable);
int sdb_profile_read(sdb_handle h, double* ms_out /* [SDB_PK_N] */, uint64_t* count_out /* [SDB_PK_N] */);

/* ---- registry: register_agent (M:314-349) / deregister_agent (M:351-372) ------------------
 * Rings exist for every index < max_agents; registration only moves the watermark that
 * bounds the per-batch commit sweep.  Deregistration keeps the ring and its read position
 * (the Kafka consumer-group offset survives Consumer.close(), SURVEY App. A rule 11). */
int sdb_register_agents(sdb_handle h, uint32_t n, const uint32_t* agent_idx);
int sdb_deregister_agents(sdb_handle h, uint32_t n, const uint32_t* agent_idx);

/* ---- groups: add_agent_group (M:1208-1227), overwrite semantics, duplicates kept --------- */
int sdb_create_group(sdb_handle h, uint32_t group_idx, uint32_t n_members, const uint32_t* member_idx);

/* ---- enqueue --------------------------------------------------------------------------
 * All three take a batch of `n` sends in call order (struct-of-arrays).  payload_off[i] is a
 * byte offset into `payload`, 16-byte aligned, with len[i] bytes valid and readable up to
 * the next multiple of 16.  Sequence numbers are assigned in array order starting at the
 * handle's next_seq; *seq_base_out (nullable) receives the first one.
 *
 * sdb_send_batch        point-to-point: send_message (M:393-519) -> Producer.produce (M:476-482);
 *                       send i gets seq_base + i.
 * sdb_send_group_batch  send_to_group (M:1229-1279): one copy per member != sender, in member
 *                       order; member j of send i gets seq_base + sum(size of earlier groups) + j
 *                       (ids of skipped members are left unused).
 * sdb_send_list_batch   broadcast (M:449-463, M:810-850): recipient list i is
 *                       list_idx[list_off[i] .. list_off[i+1]); every copy shares ONE seq
 *                       (one Message, one id), receiver field = SDB_NO_RECEIVER.
 */
int sdb_send_batch(sdb_handle h, uint32_t n,
                   const uint32_t* sender, const uint32_t* receiver,
                   const uint8_t* prio, const uint8_t* type, const uint16_t* len,
                   const uint64_t* payload_off, const uint8_t* payload, uint64_t payload_bytes,
                   const double* timestamp /* [n] or NULL = 0.0 */, uint64_t* seq_base_out);
int sdb_send_group_batch(sdb_handle h, uint32_t n,
                         const uint32_t* sender, const uint32_t* group_idx,
                         const uint8_t* prio, const uint8_t* type, const uint16_t* len,
                         const uint64_t* payload_off, const uint8_t* payload, uint64_t payload_bytes,
                         const double* timestamp, uint64_t* seq_base_out);
int sdb_send_list_batch(sdb_handle h, uint32_t n,
                        const uint32_t* sender, const uint64_t* list_off /* [n+1] */, const uint32_t* list_idx,
                        const uint8_t* prio, const uint8_t* type, const uint16_t* len,
                        const uint64_t* payload_off, const uint8_t* payload, uint64_t payload_bytes,
                        const double* timestamp, uint64_t* seq_base_out);

/* Mixed batch in call order - what a host-side send buffer flushes (the reference's producer
 * also batches, linger.ms = 10, M:197).  kind[i]: 0 = point-to-point (target = receiver index),
 * 1 = group send (target = group index), 2 = broadcast list (target = list number t, recipients
 * list_idx[list_off[t] .. list_off[t+1])).  Sequence numbers advance by 1, group size, 1. */
int sdb_send_mixed_batch(sdb_handle h, uint32_t n,
                         const uint32_t* sender, const uint8_t* kind, const uint32_t* target,
                         uint32_t n_lists, const uint64_t* list_off, const uint32_t* list_idx,
                         const uint8_t* prio, const uint8_t* type, const uint16_t* len,
                         const uint64_t* payload_off, const uint8_t* payload, uint64_t payload_bytes,
                         const double* timestamp, uint64_t* seq_base_out);

/* Staged variant used to keep inputs resident in HBM (bench `value` leg): stage copies the
 * batch to device memory once; submit enqueues it (may be repeated; each submit takes fresh
 * sequence numbers).  kind: 0 = p2p (second index array = receiver), 1 = group. */
typedef struct sdb_staged* sdb_staged_t;
int sdb_stage_batch(sdb_handle h, uint32_t kind, uint32_t n,
                    const uint32_t* sender, const uint32_t* receiver_or_group,
                    const uint8_t* prio, const uint8_t* type, const uint16_t* len,
                    const uint64_t* payload_off, const uint8_t* payload, uint64_t payload_bytes,
                    const double* timestamp, sdb_staged_t* out);
int sdb_submit_staged(sdb_handle h, sdb_staged_t s, uint64_t* seq_base_out);
int sdb_free_staged(sdb_handle h, sdb_staged_t s);

/* ---- dequeue: receive_messages (M:521-601) -> Consumer.poll (M:557-559) ------------------
 * For each listed agent (agent_idx == NULL: every index below the registration watermark)
 * remove up to max_messages pending records and return them, agents in list order, records
 * in delivery order:
 *   flags == 0                 stream (arrival) order - the reference's behaviour;
 *   flags & SDB_RECV_PRIORITY  (priority desc, arrival asc) - the extension of SURVEY
 *                              App. A rule 9; equals stream order when priorities are equal.
 * Outputs (host pointers, each nullable to skip the copy): count_out[n_agents];
 * hdr_out[total]; payload_out = payloads back to back, each padded to SDB_GRANULE, in the
 * same order (offset of record r = sum of pad32(hdr_out[0..r).len)).  *total_out /
 * *payload_bytes_out receive the totals.  Capacity: hdr_cap records / payload_cap bytes;
 * agents that would exceed max_recv_records are truncated (their remaining records stay
 * queued), never dropped.
 */
#define SDB_RECV_PRIORITY 1u
#define SDB_RECV_PEEK 2u                /* return what a receive would deliver, retire nothing (history snapshots, M:852-892) */
#define SDB_RECV_ASYNC 4u               /* device-resident consumers: enqueue the receive and return at once; all host output
                                           pointers must be NULL, results stay in the buffers of sdb_last_receive_dev (valid in
                                           stream order), totals through sdb_last_receive_totals */
#define SDB_RECV_OWNED 8u               /* sharded handles, agent_idx == NULL: the agents THIS shard owns (ascending index; count_out has
                                           one entry per owned agent) instead of every index below the watermark - the other agents'
                                           rings are empty here by construction, so nothing is lost and 1/num_shards of the work is done */
int sdb_receive_batch(sdb_handle h, uint32_t n_agents, const uint32_t* agent_idx,
                      uint32_t max_messages, uint32_t flags,
                      uint32_t* count_out, sdb_msg_header* hdr_out, uint64_t hdr_cap,
                      uint8_t* payload_out, uint64_t payload_cap,
                      uint64_t* total_out, uint64_t* payload_bytes_out);
/* Low-latency dequeue server (opt-in): receive_messages (M:521-601) for ONE agent in a few microseconds.
 * enable = 1 starts a persistent single-CTA kernel on its own stream that polls a mailbox in mapped pinned host memory;
 * sdb_receive_batch calls naming exactly one agent with host outputs (what the Python surface's receive_messages makes)
 * then post a request there and read the answer - count, headers and payloads - from pinned host memory the kernel wrote
 * directly: no kernel launch, no cudaStreamSynchronize, no cudaMemcpy on the path.  Results are identical to the
 * ordinary path (same selection code).  enable = 0 stops it (also done by sdb_destroy).  While it runs the handle
 * stops it around its own calls that free device memory (cudaFree waits for running kernels) and restarts it on the
 * next single-agent receive; callers must likewise not call cudaDeviceSynchronize / cudaFree from outside while it runs. */
int sdb_latency_server(sdb_handle h, int enable);

/* Device-resident results of the LAST receive call (valid until the next one). */
int sdb_last_receive_dev(sdb_handle h, const uint32_t** count_dev, const sdb_msg_header** hdr_dev,
                         const uint8_t** payload_dev);
/* Records and payload bytes of the LAST receive call; waits for it (the host side of SDB_RECV_ASYNC). */
int sdb_last_receive_totals(sdb_handle h, uint64_t* total_out, uint64_t* payload_bytes_out);

/* ---- stream digests: per-agent delivery order + content, checkable at any scale ---------------------------------
 * What the reference promises a consumer is the ORDER and CONTENT of its own stream (drain loop M:553-601, filter
 * M:579-585).  A digest per agent captures exactly that without moving the records off the device:
 *   rec_hash = sum over the 64-bit little-endian words w_k (k = 0, 1, ..) of [sdb_msg_header | payload padded to
 *              SDB_GRANULE] of fmix64(w_k ^ ((k + 1) * 0x9E3779B97F4A7C15))  mod 2^64   (fmix64 = MurmurHash3 finaliser)
 *   chain    : d[a] <- (rotl64(d[a], 5) ^ rec_hash) * 0x9E3779B97F4A7C15, for every record delivered to agent a, in
 *              delivery order; d[a] starts at 0.
 * sdb_digest_fold folds the device-resident results of the LAST sdb_receive_batch call (bulk path; not the
 * <= 8-agent latency path) into d[]; it is stream-ordered and does not synchronise.  sdb_digest_read copies
 * d[agent_idx[i]] (agent_idx == NULL: agents 0..n-1) to the host.  Used by tests and by bench.py's parity check
 * (1M agents, 1..8 GPUs) against the same definition in oracle/cpu_ref.c. */
int sdb_digest_reset(sdb_handle h);
int sdb_digest_fold(sdb_handle h);
int sdb_digest_read(sdb_handle h, uint32_t n, const uint32_t* agent_idx, uint64_t* digest_out);

/* ---- cross-shard delivery (one handle per GPU, one process per GPU) ----------------------------
 * Replaces the partitioned topic: _get_partition (M:309-312, salted hash() % num_partitions) and
 * the explicit-partition produce (M:469-482).  Agents are hash-partitioned over num_shards GPUs;
 * sdb_set_agent_shards declares the owner of every agent index (the Python layer uses
 * fnv1a64(utf8(agent_id)) % num_shards).  Group tables are replicated: every shard calls
 * sdb_create_group with the FULL member list and keeps the members it owns plus their positions.
 *
 * A rank exports the group sends it ingested as ONE wire batch in device memory - descriptor +
 * payload per SEND, not per recipient.  The caller moves wire batches between ranks (NCCL
 * all-gather over NVLink through torch.distributed) and every rank imports the wire batches of
 * all ranks, in rank order: each send is expanded into the copies for the members this shard owns.
 * Sequence numbers are global (rank r's sends follow rank r-1's), so per-agent streams are
 * identical to a single-shard run over the concatenated batch.  With num_shards == 1 the same
 * two calls work without any collective.
 */
uint64_t sdb_wire_bytes(sdb_handle h, uint32_t max_sends, uint64_t max_payload_bytes);
int sdb_set_agent_shards(sdb_handle h, uint32_t n, const uint8_t* shard_of);
int sdb_export_group_batch(sdb_handle h, uint32_t n,
                           const uint32_t* sender, const uint32_t* group_idx,
                           const uint8_t* prio, const uint8_t* type, const uint16_t* len,
                           const uint64_t* payload_off, const uint8_t* payload, uint64_t payload_bytes,
                           const double* timestamp, void* wire_dev, uint64_t wire_cap);
/* Same, for a mixed batch in call order (kinds as in sdb_send_mixed_batch): point-to-point sends are
 * delivered only by the shard that owns the receiver, broadcast lists by every shard for the
 * recipients it owns.  Recipient lists travel inside the wire batch (count them in max_payload_bytes). */
int sdb_export_mixed_batch(sdb_handle h, uint32_t n,
                           const uint32_t* sender, const uint8_t* kind, const uint32_t* target,
                           uint32_t n_lists, const uint64_t* list_off, const uint32_t* list_idx,
                           const uint8_t* prio, const uint8_t* type, const uint16_t* len,
                           const uint64_t* payload_off, const uint8_t* payload, uint64_t payload_bytes,
                           const double* timestamp, void* wire_dev, uint64_t wire_cap);
/* Same, but the exporter numbers its sends itself, starting at `seq_base` (non-zero).  Front-ends that must
 * return message ids at send time - before the collective flush - use a composite number
 * (round << 40 | rank << 32 | local index), which keeps the global (round, rank, local) delivery order. */
int sdb_export_mixed_batch_seq(sdb_handle h, uint64_t seq_base, uint32_t n,
                               const uint32_t* sender, const uint8_t* kind, const uint32_t* target,
                               uint32_t n_lists, const uint64_t* list_off, const uint32_t* list_idx,
                               const uint8_t* prio, const uint8_t* type, const uint16_t* len,
                               const uint64_t* payload_off, const uint8_t* payload, uint64_t payload_bytes,
                               const double* timestamp, void* wire_dev, uint64_t wire_cap);
int sdb_import_wire_batches(sdb_handle h, uint32_t n_src, const void* wire_dev_all, uint64_t wire_stride,
                            uint64_t* seq_base_out);
/* Peer-memory transport (no collective): each rank exports into a buffer allocated with
 * sdb_wire_alloc, hands the 64-byte CUDA-IPC handle to the other ranks (any side channel), which
 * map it with sdb_wire_open; after a cross-rank barrier every rank imports from the table of
 * pointers (its own buffer + the mapped peer buffers, rank order).  The import kernels and the
 * fan-out kernel's TMA loads then read descriptors and payloads directly out of the exporting
 * GPU's memory over NVLink, overlapped with the fan-out itself. */
int sdb_wire_alloc(sdb_handle h, uint64_t bytes, void** dev_out, void* ipc_handle_out /* 64 bytes, nullable */);
int sdb_wire_open(sdb_handle h, const void* ipc_handle, void** dev_out);
int sdb_wire_close(sdb_handle h, void* dev, int opened /* 1: from sdb_wire_open, 0: from sdb_wire_alloc */);
int sdb_import_wire_ptrs(sdb_handle h, uint32_t n_src, const void* const* wire_ptrs, uint64_t* seq_base_out);
/* Flag-synchronised, host-asynchronous form of the peer-memory transport: no collective and no host round trip.
 * Every export buffer of `wire_bytes` (as sized by sdb_wire_bytes) ends in a 128-byte control block with two
 * counters: `ready` (last step whose export into the buffer is complete) and `done` (last step for which the
 * buffer's OWNER finished importing every rank's buffer of that parity).  Steps count 1, 2, 3 ...; with two
 * alternating buffers per rank, buffer step & 1 is used for step.
 *   sdb_wire_wait_done     stream-ordered wait until every listed buffer's owner reports done >= step; call it with
 *                          (step - 2) before exporting `step` into a buffer that peers read two steps ago
 *   sdb_wire_publish       stream-ordered: ready = step (after the export's copies, which run on the same stream)
 *   sdb_import_wire_ptrs_async
 *                          waits ON THE DEVICE for ready >= step of every source (peer flags are polled over NVLink),
 *                          then places and expands the import entirely on the device - arena position, sequence base
 *                          and totals never visit the host (a device-resident cursor is the authority until the next
 *                          call that needs them on the host) - and finally reports done = step in this rank's own
 *                          buffer (wire_ptrs[shard_id]).  Three kernels do the work: one fused localize (wire headers,
 *                          descriptors and group buckets read straight out of the exporting GPUs, decoupled
 *                          look-back scan, placement), the fan-out (TMA pulls of the payloads over NVLink), the
 *                          group-parallel index build.  An import that does not fit (arena / list pool) is dropped
 *                          WHOLE and reported by the next host-synchronising call (SDB_EARENA_FULL / SDB_ECAPACITY).
 *                          Traffic outside the fast shape (payloads above 512 bytes, agents in several groups) takes
 *                          the synchronous import between the same flags. */
int sdb_wire_wait_done(sdb_handle h, uint32_t n_src, const void* const* wire_ptrs, uint64_t wire_bytes, uint32_t step);
int sdb_wire_publish(sdb_handle h, void* wire_dev, uint64_t wire_bytes, uint32_t step);
int sdb_import_wire_ptrs_async(sdb_handle h, uint32_t n_src, const void* const* wire_ptrs, uint64_t wire_bytes, uint32_t step);
/* Pipelining across steps: runs the flag wait and the localize pass of step `step` on a second stream of the handle,
 * ordered after everything enqueued on the shard's stream so far and BESIDE what the caller enqueues next (typically
 * the receive of the previous step, which is bound by HBM while the localize pass is bound by NVLink latency).  The
 * following sdb_import_wire_ptrs_async(.., step) finds the batch localized and only places it (one single-thread
 * kernel: arena base and sequence base still come from the device cursor, after the previous step's commit), then
 * fans out and indexes as usual: delivery order and content are those of the unpipelined sequence.  Two buffer sets
 * alternate by step parity.  The export buffers of `step` must not be rewritten and group membership / ownership must
 * not change between the two calls.  A no-op for traffic outside the fast shape. */
int sdb_import_prefetch(sdb_handle h, uint32_t n_src, const void* const* wire_ptrs, uint64_t wire_bytes, uint32_t step);

/* ---- which records were lost to full rings (delivery report, M:374-391 / M:501-519)
 * `ring_overflow` in sdb_stats says how many records found their receiver's ring full; this call says WHICH: up to
 * `cap` (receiver agent, sequence number) pairs of records dropped since the previous call, in no particular order
 * (the device logs the first 4096 per interval; *dropped_out = how many were dropped in the interval).  Reading
 * clears the log.  Call it right after the batch that overflowed: the sequence number is read from the dropped
 * record's image in the arena log, which is reclaimed like any other record. */
int sdb_overflow_log(sdb_handle h, uint32_t cap, uint32_t* agent_out, uint64_t* seq_out, uint32_t* n_out, uint64_t* dropped_out);

/* ---- inbox / load queries on the device: get_agent_load, get_unread_message_count (M:1026-1094), get_stats (M:973-1024)
 * The reference answers these by walking its host dictionaries (`messages`, `agent_inbox`); the queue itself now lives
 * in HBM, so the same questions are answered from the rings.  Mapping of the reference's fields:
 *   inbox_size    (len(agent_inbox[a]), M:1076)            = received      records ever enqueued for the agent
 *   unread_count  (status == DELIVERED in the inbox, M:1040) = pending       records still queued (not yet received)
 *   READ / PROCESSED side                                   = received - pending
 * sdb_agent_loads fills one sdb_agent_load per listed agent (agent_idx == NULL: agents 0..n-1).
 * sdb_queue_stats reduces over every agent below the watermark: totals, the pending records per priority level (a
 * histogram over the ring metadata), the deepest queue.  Both synchronise (they return data). */
typedef struct sdb_agent_load {
  uint32_t received;             /* records enqueued for the agent so far (mod 2^32) */
  uint32_t pending;              /* records waiting in its ring */
  uint32_t pending_by_prio[4];   /* ... by priority level (M:35-41) */
  uint32_t pending_granules;     /* 32-byte granules of pending payload (without headers) */
  uint32_t reserved;
} sdb_agent_load;
typedef struct sdb_queue_summary {
  uint64_t agents_with_pending;
  uint64_t pending;
  uint64_t pending_by_prio[4];
  uint64_t pending_granules;
  uint64_t received;             /* sum over agents of `received` */
  uint64_t max_pending;          /* deepest queue */
  uint64_t max_pending_agent;
} sdb_queue_summary;
int sdb_agent_loads(sdb_handle h, uint32_t n, const uint32_t* agent_idx, sdb_agent_load* out);
int sdb_queue_stats(sdb_handle h, sdb_queue_summary* out);

/* ---- LLM backend balancer: set_llm_load_balancing / assign_llm_backend / get_llm_backend
 * (M:1281-1325).  The reference stores a flag and a dict and has NO pick algorithm
 * (SURVEY 0.5); the per-agent sticky map stays in the Python layer, and these entry points
 * add the batched pick over live load counters:
 *   mode 0  weighted least-load: requests are served in index order, each goes to
 *           argmin_b load[b]/weight[b] (exact rational compare, ties -> lowest index) and adds
 *           cost[i] (NULL = 1) to that backend's load;
 *   mode 1  weighted random (one-item weighted reservoir, Efraimidis-Spirakis exponential
 *           race in integer fixed point): P(b) proportional to weight[b], keyed by
 *           (seed, request index); loads are then incremented atomically.
 */
int sdb_set_backends(sdb_handle h, uint32_t n, const uint32_t* weight, const uint64_t* load0);
int sdb_get_backend_loads(sdb_handle h, uint32_t n, uint64_t* load_out);
int sdb_release_backends(sdb_handle h, uint32_t n, const uint32_t* backend, const uint32_t* cost /* NULL = 1 */);
int sdb_select_backend_batch(sdb_handle h, uint32_t n_req, const uint32_t* cost, uint32_t mode,
                             uint64_t seed, uint32_t* backend_out);
/* "get_agent_load is the only load signal" (SURVEY 8a R10, M:1049-1094): feed the balancer from the queue itself.
 * sdb_assign_agent_backends is the device copy of the sticky agent -> backend map (assign_llm_backend, M:1293-1311;
 * backend index, or 0xFFFFFFFF = none).  sdb_backend_loads_from_queues sets load[b] = sum of the pending records of the
 * agents assigned to b (one pass over the ring headers, stream-ordered, no host round trip); picks then see real
 * backlog.  Call it before sdb_select_backend_batch whenever the loads should track the queues. */
int sdb_assign_agent_backends(sdb_handle h, uint32_t n, const uint32_t* agent_idx, const uint32_t* backend_idx);
int sdb_backend_loads_from_queues(sdb_handle h);

#ifdef __cplusplus
}
#endif
#endif /* SWARMDB_B200_H */
