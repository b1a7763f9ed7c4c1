// Build identity of libsqcircuit_b200.so: SHA-256 over the kernel sources and the public header, passed in by
// the build (__graft_entry__.build) and compared by the loader (sqcircuit_b200/_lib.py) with a hash of the
// sources it finds next to the library: a stale shipped binary is refused instead of silently benchmarked.
#include "../../include/sqcircuit_b200.h"

#ifndef SQC_SOURCE_HASH
#define SQC_SOURCE_HASH "unknown"
#endif

// the tag makes the hash findable in the file without loading the library (build staleness check)
static const char kTag[] = "SQC_SOURCE_HASH=" SQC_SOURCE_HASH;

extern "C" const char* sqc_source_hash(void) { return kTag + 16; }
