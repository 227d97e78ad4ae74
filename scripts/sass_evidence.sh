#!/bin/bash
# SASS evidence that the TMA / mbarrier path is what the built library contains (B200_PROFILING.md: "What proves a
# Blackwell-native kernel").  Per kernel: bulk-copy (UBLKCP = cp.async.bulk, TMA engine), mbarrier (SYNCS) and 256-bit
# memory instruction counts.  Usage: scripts/sass_evidence.sh > profiles/r2_sass_tma.txt
lib=swarmdb_b200/csrc/libswarmdb_b200.so
echo "# cuobjdump -sass $lib (sm_100a), instruction counts per kernel; built $(date -u +%Y-%m-%dT%H:%MZ)"
/usr/local/cuda/bin/cuobjdump -sass $lib | awk '
  /Function :/ { name=$3; next }
  /UBLKCP\.S\.G/ { lg[name]++ } /UBLKCP\.G\.S/ { sg[name]++ } /SYNCS/ { sy[name]++ }
  /LDG\.E\.(ENL2\.)?256|STG\.E\.(ENL2\.)?256/ { w256[name]++ }
  /REDG|ATOMG|RED\.E/ { at[name]++ }
  /UTCMMA|HMMA|LDTM/ { tc[name]++ }
  END { printf "%-70s %8s %8s %6s %6s %6s %6s\n", "kernel", "UBLKCP.S.G", "UBLKCP.G.S", "SYNCS", "256bit", "atomic", "tensor";
        for (n in sy) seen[n]=1; for (n in lg) seen[n]=1; for (n in sg) seen[n]=1; for (n in at) seen[n]=1; for (n in w256) seen[n]=1;
        for (n in seen) printf "%-70s %8d %8d %6d %6d %6d %6d\n", substr(n,1,70), lg[n], sg[n], sy[n], w256[n], at[n], tc[n] }' | sort
