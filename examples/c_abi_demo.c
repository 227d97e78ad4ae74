/* Minimal C caller of the drop-in boundary (include/swarmdb_b200.h): create a shard, make a group,
 * fan one message out to it, drain one member.  Build:
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -Lswarmdb_b200/csrc -lswarmdb_b200 -Wl,-rpath,$PWD/swarmdb_b200/csrc -o c_abi_demo
 * Without a CUDA device sdb_create fails with SDB_ECUDA and a message - there is no CPU fallback. */
#include <stdio.h>
#include <string.h>

#include "swarmdb_b200.h"

int main(void) {
  sdb_config cfg;
  sdb_handle h = NULL;
  memset(&cfg, 0, sizeof cfg);
  cfg.struct_bytes = (uint32_t)sizeof cfg;
  cfg.num_shards = 1; cfg.max_agents = 1024; cfg.ring_slots = 64; cfg.arena_bytes = 1ull << 24;
  cfg.max_payload_bytes = 256; cfg.max_groups = 8; cfg.member_pool_entries = 4096; cfg.fanout_variant = 2;
  printf("abi %d sizeof(sdb_config)=%u sizeof(sdb_msg_header)=%u sizeof(sdb_stats)=%u\n", sdb_abi_version(),
         (unsigned)sizeof(sdb_config), (unsigned)sizeof(sdb_msg_header), (unsigned)sizeof(sdb_stats));
  int rc = sdb_create(&cfg, &h);
  if (rc != SDB_OK) {
    printf("sdb_create failed: rc=%d (%s)\n", rc, sdb_last_error(h));
    if (h) sdb_destroy(h);
    return 3;
  }
  {
    const uint32_t members[3] = {1, 2, 3};
    const uint32_t sender[1] = {7}, group[1] = {0};
    const uint16_t len[1] = {5};
    const uint64_t off[1] = {0};
    uint8_t payload[32] = "hello";
    uint64_t seq = 0, total = 0, bytes = 0;
    const uint32_t who[1] = {2};
    uint32_t count[1];
    sdb_msg_header hdr[4];
    uint8_t out[4 * 256];
    if ((rc = sdb_create_group(h, 0, 3, members)) != SDB_OK) goto fail;
    if ((rc = sdb_send_group_batch(h, 1, sender, group, NULL, NULL, len, off, payload, sizeof payload, NULL, &seq)) != SDB_OK) goto fail;
    if ((rc = sdb_receive_batch(h, 1, who, 100, 0, count, hdr, 4, out, sizeof out, &total, &bytes)) != SDB_OK) goto fail;
    printf("agent 2 got %u message(s): seq=%llu sender=%u len=%u \"%.*s\"\n", count[0], (unsigned long long)hdr[0].seq,
           hdr[0].sender, hdr[0].len, (int)hdr[0].len, (const char*)out);
    sdb_destroy(h);
    return (total == 1 && hdr[0].seq == seq + 1 && memcmp(out, "hello", 5) == 0) ? 0 : 1;
  }
fail:
  printf("call failed: rc=%d (%s)\n", rc, sdb_last_error(h));
  sdb_destroy(h);
  return 2;
}
