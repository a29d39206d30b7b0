/* include/csfm.h — C ABI of the B200-native FM-index query engine (libcsfm.so).
 *
 * This is the drop-in boundary for the reference's hot path. The reference has no FFI layer:
 * its boundary is the C++ class cs::FMIndex (/root/reference/src/api/fm_index.hpp:9-68). Each
 * entry point below names the reference interface it replaces; the C++ class that mirrors
 * cs::FMIndex on top of this ABI lives in
 * compressed-fm-index-implementation-with-learned-optimizations_b200/host/src/api/fm_index.hpp.
 *
 * Conventions
 *  - plain pointers and sizes only; nothing throws across the boundary: every function returns
 *    a csfm_status (0 = ok) and csfm_last_error() gives the message for the calling thread;
 *  - "host" functions take host pointers (pageable or pinned; pinned buffers from
 *    csfm_host_alloc avoid a staging copy) and return when the results are in the out buffers;
 *  - "_device" functions take device pointers on the index's device and a cudaStream_t passed
 *    as void*; they only enqueue work (results are ready when the stream reaches that point),
 *    except where a host-visible total is returned (documented per function);
 *  - patterns travel packed: `bytes` holds all pattern bytes back to back, `offs[npat+1]`
 *    (uint64) the start of each pattern, offs[npat] = total bytes. Patterns are borrowed. Offsets must not
 *    decrease: the host-pointer entry points check that (CSFM_ERR_INVALID); for the "_device" forms it is a
 *    PRECONDITION (the offsets live on the device and the kernels trust them);
 *  - all results are bit-exact with the reference on the same text and patterns, including its
 *    quirks: count("") == n, cyclic over-count without a terminator, locate
 *    positions in SA-row order, an error status where the reference throws;
 *  - n must be < 2^32 - 1, as in the reference (uint32 SA / C array, fm_index.hpp:43-44);
 *  - threads: every call on a handle is serialised by that handle's mutex (safe, not concurrent); host threads that
 *    want to query one index concurrently take an alias each (csfm_alias: a second handle over the same device blob,
 *    own streams and workspaces, no copy);
 *  - a single pattern (npat == 1, up to 64 bytes, layout 2) takes a one-launch path: the pattern rides in the kernel
 *    parameters and the result comes back through mapped pinned memory, about 10 us per call;
 *  - there is no CPU fallback: without a usable CUDA device every call fails with
 *    CSFM_ERR_CUDA.
 */
#ifndef CSFM_H
#define CSFM_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CSFM_API __attribute__((visibility("default")))
#else
#define CSFM_API
#endif

typedef struct csfm_index csfm_index; /* opaque; one handle <-> one device */

/* Library status codes. */
typedef enum {
  CSFM_OK = 0,
  CSFM_ERR_INVALID = 1,   /* bad argument */
  CSFM_ERR_CUDA = 2,      /* CUDA runtime / driver error, or no device */
  CSFM_ERR_NOMEM = 3,     /* host or device allocation failed */
  CSFM_ERR_TOO_LARGE = 4, /* n >= 2^32 - 1 */
  CSFM_ERR_CAPACITY = 5,  /* locate output buffer too small; *total tells the need */
  CSFM_ERR_FORMAT = 6     /* malformed index blob */
} csfm_status;

/* Per-query status of locate (where cs::FMIndex::locate throws, fm_index.cpp:136-146). */
typedef enum {
  CSFM_Q_OK = 0,
  CSFM_Q_LF_WALK_EXCEEDED = 1, /* "locate: LF walk exceeded text length" (fm_index.cpp:136-138) */
  CSFM_Q_SSA_OOB = 2           /* "locate: SSA sample index out of range" (fm_index.cpp:141-146) */
} csfm_query_status;

/* Mirrors cs::BuildParams (fm_index.hpp:11-14) field for field. As in the reference, only
 * ssa_stride is honoured (fm_index.cpp:58); S, s and eps are accepted and ignored. */
typedef struct {
  uint32_t S, s, ssa_stride;
  double eps;
} csfm_params;

/* Build flags. */
#define CSFM_BUILD_DEFAULT 0u
#define CSFM_BUILD_NO_COMPACT 1u /* index raw byte values (8 bits) even if the text uses few symbols */
#define CSFM_BUILD_KEEP_SA 2u    /* keep the full suffix array on the device for csfm_get_sa */
#define CSFM_BUILD_LAYOUT_BINARY64 4u /* layout 1: binary wavelet matrix in 64-byte lines (one bit
                                         per line fetch) instead of the default layout 2: 16-ary
                                         levels in 128-byte lines (four bits per line fetch) */
#define CSFM_BUILD_NO_TEXT_CHECK 16u /* do not keep the text + suffix array for the verification
                                        shortcut (kept by default when the text ends in a unique
                                        smallest byte: a query whose interval has shrunk to one row
                                        then compares its remaining characters with the text) */
#define CSFM_BUILD_FORCE_TEXT_CHECK 32u /* keep them for every text with such a terminator. By default the
                                           shortcut is built only for two-level indexes (more than 16
                                           distinct bytes) whose levels exceed 96 MB: below that, stepping
                                           through L2-resident lines, or through one-fetch steps, is faster
                                           than the two HBM fetches of a verification */
#define CSFM_BUILD_LARGE_TABLE 64u /* spend device memory on the k-mer table: budget = a third of the free
                                      device memory, at most 40 GiB, instead of a quarter of the level bytes,
                                      at most 1 GiB. The lookup then leaves only a few rows and the query goes
                                      straight to the text verification (implies CSFM_BUILD_FORCE_TEXT_CHECK):
                                      three dependent fetches per query. Same results, 1.5-1.9x the count
                                      throughput, an index of 10-25 GB */
#define CSFM_BUILD_LAYOUT_NIBBLE128 128u /* keep layout 2 (16-ary levels, 128-byte lines) for a text that qualifies for
                                           layout 3: two-bit symbols in 64-byte lines, one fetch per rank by a two-lane
                                           sub-warp, chosen by default when the text has at most four distinct bytes, or
                                           five of which one occurs exactly once (DNA + terminator): a third of the
                                           bytes of layout 2, so configs[1] and configs[3] stay in the L2 */
#define CSFM_BUILD_NO_KMER_TABLE 8u /* do not build the k-mer jump table (layout 2 builds one by
                                       default: the first k steps of a query become one lookup) */
#define CSFM_BUILD_ROW_SAMPLES 256u /* layout 3: locate walks LF to a row-sampled suffix-array entry like the
                                       reference (fm_index.cpp:125-153). By default, when the text's last byte is the
                                       symbol that occurs once (DNA + terminator), csfm_build_from_text samples the
                                       suffix array by TEXT POSITION as well — the same ceil(n / stride) samples, found
                                       through one mark bit per row kept inside the level lines (128 rows per 64-byte
                                       line instead of 192) — so a walk takes at most stride - 1 steps, half as many on
                                       average; positions are SA[row] either way */

typedef struct {
  uint64_t n;          /* text length (cs::IndexMeta::n, fm_index.hpp:15) */
  uint32_t sigma;      /* distinct byte values present in the text */
  uint32_t levels;     /* levels stored = dependent line fetches per rank: layout 2: 1 (sigma <= 16)
                          or 2; layout 1: ceil(log2 sigma), or 8 with CSFM_BUILD_NO_COMPACT */
  uint32_t ssa_stride; /* SA sample stride (cs::SSA::stride) */
  uint32_t device;     /* CUDA device ordinal */
  uint64_t nsamp;      /* number of SA samples = ceil(n / stride) */
  uint64_t blocks_per_level; /* 64-byte lines per level */
  uint64_t blob_bytes; /* size of the device-resident index */
  uint32_t has_sa;     /* full SA still resident (CSFM_BUILD_KEEP_SA) */
  uint32_t layout;     /* 1 = binary / 64-byte lines, 2 = 16-ary / 128-byte lines, 3 = two-bit symbols / 64-byte lines */
  uint32_t line_bytes; /* bytes fetched per rank per level: 64 or 128 */
  uint32_t kmer_k;     /* length of the k-mer jump table's keys, 0 = no table */
  uint32_t text_check; /* 1 = text + suffix array resident for the verification shortcut */
  uint32_t half_table; /* 1 = half-step table resident (two-level indexes: the first rank step of a
                          query needs only its second level) */
  uint32_t sa_rounds;  /* build_from_text: sorting rounds of the suffix sort (1 = the packed first
                          symbols already separated every suffix), 0 if the index was not built here */
  uint32_t sa_radix_passes; /* 8-bit radix passes of the rounds that sorted ALL n (u64 key, u32 suffix) pairs */
  uint64_t sa_pair_passes;  /* sum over all rounds of pairs sorted x 8-bit passes: the sort moved about
                               sa_pair_passes * 2 * 12 bytes. Rounds after the first sort only the suffixes whose
                               group still has more than one member once those are at most half of the text */
  uint32_t position_samples; /* 1 = layout 3 in its marked form: locate walks to a suffix-array sample taken by text position */
  uint32_t pad0;
} csfm_index_info;

/* Counters describing the most recent query call on this handle (for bench accounting). */
typedef struct {
  uint64_t kernel_launches; /* CUDA kernels this library launched in the call */
  uint64_t h2d_bytes, d2h_bytes;
  uint64_t search_steps;    /* backward-search steps executed through rank (incl. the free first
                               step of queries that did not use the table); 0 unless asked */
  uint64_t lf_steps;        /* LF steps walked by locate; 0 unless asked */
  float kernel_ms;          /* device time of the dominant kernel (CUDA events), 0 unless asked */
  uint32_t table_lookups;   /* queries that started from the k-mer table (their first k steps are
                               not in search_steps); 0 unless asked */
  uint32_t text_checks;     /* queries finished by comparing their remaining characters with the
                               text (those characters are not in search_steps); 0 unless asked */
  uint32_t half_steps;      /* first steps taken from the half-step table: one level (2 lines) instead
                               of two; not in search_steps; 0 unless asked */
  uint32_t pad0;
  uint64_t line_fetches;    /* level lines the rank steps of the call actually loaded (sp and ep share one
                               load when they fall in the same line); layout 2 count kernels; 0 unless asked */
} csfm_call_stats;

CSFM_API const char* csfm_last_error(void);
CSFM_API const char* csfm_version(void);
CSFM_API int csfm_device_count(int* count);

/* ---- construction: replaces cs::FMIndex::build_from_text (fm_index.cpp:16-69) ------------- */

/* text: n host bytes. Runs SA -> BWT -> C -> wavelet matrix -> SSA entirely on the device
 * (prefix-doubling radix sort; order of build_sa_naive, sais.hpp:8-16). */
CSFM_API int csfm_build_from_text(const uint8_t* text, uint64_t n, const csfm_params* params,
                                  int device, uint32_t flags, csfm_index** out);
/* Same, text already resident on `device`. */
CSFM_API int csfm_build_from_text_device(const uint8_t* d_text, uint64_t n,
                                         const csfm_params* params, int device, uint32_t flags,
                                         csfm_index** out);
/* From build products computed elsewhere (host pointers): BWT (bwt.hpp:7-15) and the sampled SA
 * (samples[k] = SA[k*stride], fm_index.cpp:57-65). C and the wavelet matrix are derived here. */
CSFM_API int csfm_build_from_parts(const uint8_t* bwt, uint64_t n, const uint32_t* ssa,
                                   uint64_t nsamp, uint32_t ssa_stride, int device,
                                   uint32_t flags, csfm_index** out);
CSFM_API void csfm_destroy(csfm_index* idx);
CSFM_API int csfm_info(const csfm_index* idx, csfm_index_info* out);

/* ---- build products back to the host (verification, .csidx writer, CPU baseline) ---------- */
CSFM_API int csfm_get_C(const csfm_index* idx, uint32_t C[257]);           /* fm_index.cpp:36-47 */
CSFM_API int csfm_get_ssa(const csfm_index* idx, uint32_t* out /*nsamp*/); /* fm_index.cpp:57-65 */
CSFM_API int csfm_get_sa(const csfm_index* idx, uint32_t* out /*n*/);      /* needs KEEP_SA */
/* Device pointer of the resident suffix array (n x uint32), for on-device certification. */
CSFM_API int csfm_sa_device(const csfm_index* idx, const uint32_t** d_sa);
CSFM_API int csfm_release_sa(csfm_index* idx);
/* BWT re-derived from the wavelet matrix by an access kernel (wavelet.cpp:102-128). */
CSFM_API int csfm_extract_bwt(const csfm_index* idx, uint8_t* out /*n*/);

/* cs::FMIndex::extract (fm_index.cpp:163-167) from the index itself, for handles that have no host copy of the text
 * (an index loaded from a .csidx without its TEXT section, an attached blob): out receives T[pos, pos + *got),
 * *got = min(len, n - pos), 0 when pos >= n. Served from the blob's text section when it has one, else from a device
 * copy of the text rebuilt once out of the index (n LF steps in all, layouts 2 and 3). CSFM_ERR_INVALID when the
 * text cannot be rebuilt: it must end in a unique smallest byte (otherwise the reference's BWT is not a rotation
 * BWT and LF does not walk the text backwards; locate() over-counts or throws on such texts too). */
CSFM_API int csfm_extract(csfm_index* idx, uint64_t pos, uint64_t len, uint8_t* out, uint64_t* got);

/* ---- replication: the whole index is one contiguous device blob --------------------------- */
/* Pointer/size of the blob on the index's device. Broadcast it (one ncclBroadcast) and attach. */
CSFM_API int csfm_blob(const csfm_index* idx, const void** d_blob, uint64_t* bytes);
/* Wraps a blob already resident on `device` (e.g. the receive buffer of the broadcast). With
 * take_ownership == 0 the caller keeps the memory alive for the life of the handle. */
CSFM_API int csfm_attach_blob(void* d_blob, uint64_t bytes, int device, int take_ownership,
                              csfm_index** out);
/* A second handle over the SAME blob on the same device, without a copy: its own streams, workspaces, statistics and
 * mutex. Calls on ONE handle are serialised by its mutex; host threads that want to query one index concurrently each
 * take an alias (the C++ class cs::FMIndex does that by itself for every thread other than the one that built the index).
 * The alias borrows the blob: destroy it before the handle that owns the blob. */
CSFM_API int csfm_alias(const csfm_index* idx, csfm_index** out);
/* A second handle over a copy of the blob on `device` (the same device, or a peer: the copy goes over
 * NVLink when peer access can be enabled, else through the host). For single-process callers that want
 * one replica per GPU without a communicator; the new handle owns its copy. */
CSFM_API int csfm_replicate(const csfm_index* idx, int device, csfm_index** out);
/* Host round trip of the same blob (checkpoint / .csidx device-layout section). */
CSFM_API int csfm_blob_to_host(const csfm_index* idx, void* out, uint64_t bytes);
CSFM_API int csfm_from_host_blob(const void* blob, uint64_t bytes, int device, csfm_index** out);

/* ---- queries: replace cs::FMIndex::count (fm_index.cpp:79-101) ----------------------------- */
/* counts[q] = count(pattern q). sp_ep (nullable) receives the interval [sp,ep) after the last
 * executed step, (0,0) for empty patterns and empty results. */
CSFM_API int csfm_count_batch(csfm_index* idx, const uint8_t* bytes, const uint64_t* offs,
                              uint64_t npat, uint64_t* counts, uint64_t* sp_ep);
/* Asynchronous form for streaming callers: submit enqueues host->device copies, the kernel and
 * the device->host copy of one batch on one of CSFM_ASYNC_SLOTS internal streams and returns a
 * ticket; wait blocks until that batch's results are in `counts` / `sp_ep`. Keeping two or three
 * batches in flight overlaps the PCIe copies of one batch with the kernel of another. The host
 * buffers must stay valid (and should be pinned, csfm_host_alloc) until the ticket is waited on;
 * a slot is reused after CSFM_ASYNC_SLOTS submits (submit waits for it if still busy). */
#define CSFM_ASYNC_SLOTS 3
CSFM_API int csfm_count_batch_submit(csfm_index* idx, const uint8_t* bytes, const uint64_t* offs,
                                     uint64_t npat, uint64_t* counts, uint64_t* sp_ep,
                                     uint64_t* ticket);
/* Compact form of submit for PCIe-bound streams: 32-bit offsets (offs32[npat] < 2^32) and 32-bit
 * counts (a count never exceeds n < 2^32). Same tickets, same slots, waited on with
 * csfm_count_batch_wait; the offsets are widened and the counts narrowed on the device. */
CSFM_API int csfm_count_batch_submit32(csfm_index* idx, const uint8_t* bytes, const uint32_t* offs32,
                                       uint64_t npat, uint32_t* counts32, uint64_t* ticket);
/* Most compact form: one LENGTH byte per pattern instead of an offset (patterns of at most 255
 * bytes, packed back to back in `bytes`; nbytes must equal the sum of the lengths, which is
 * checked: CSFM_ERR_INVALID otherwise). The offsets are rebuilt by a prefix sum on the device, so
 * 1 + m bytes per pattern cross the bus instead of 4 + m. */
CSFM_API int csfm_count_batch_submit_len8(csfm_index* idx, const uint8_t* bytes, uint64_t nbytes,
                                          const uint8_t* lens8, uint64_t npat, uint32_t* counts32,
                                          uint64_t* ticket);
/* Smallest form on the wire, for small alphabets (DNA: 3 bits per symbol instead of 8): patterns travel as the index's
 * WIRE CODES — csfm_pattern_codes gives code_of_byte[] (0 .. sigma-1 for the bytes that occur, in byte order; 255 for
 * bytes that do not) and `bits` = max(1, ceil(log2 sigma)) — packed LSB-first back to back across pattern boundaries:
 * symbol i of the batch occupies bits [i * bits, (i + 1) * bits) of `packed` (ceil(nsyms * bits / 8) bytes). One length
 * byte per pattern as in the len8 form; nsyms must equal the sum of the lengths. A code >= sigma stands for a symbol
 * that does not occur (the pattern counts 0). The codes are unpacked to bytes on the device by one extra kernel. */
CSFM_API int csfm_pattern_codes(const csfm_index* idx, uint8_t code_of_byte[256], uint32_t* bits);
CSFM_API int csfm_count_batch_submit_packed(csfm_index* idx, const uint8_t* packed, uint64_t nsyms, const uint8_t* lens8,
                                            uint64_t npat, uint32_t* counts32, uint64_t* ticket);
CSFM_API int csfm_count_batch_wait(csfm_index* idx, uint64_t ticket);
CSFM_API int csfm_count_batch_device(csfm_index* idx, const uint8_t* d_bytes,
                                     const uint64_t* d_offs, uint64_t npat, uint64_t* d_counts,
                                     uint64_t* d_sp_ep, void* stream);

/* ---- queries: replace cs::FMIndex::locate (fm_index.cpp:107-157) --------------------------- */
/* For query q the positions (SA-row order, at most `limit`) are out_pos[out_offs[q] ..
 * out_offs[q+1]). status[q] is a csfm_query_status; positions of a failed query are
 * unspecified. *total = out_offs[npat]. If cap < total nothing is written to out_pos and
 * CSFM_ERR_CAPACITY is returned with out_offs and *total filled (call again with room);
 * out_pos == NULL && cap == 0 is the sizing call and returns CSFM_OK. */
CSFM_API int csfm_locate_batch(csfm_index* idx, const uint8_t* bytes, const uint64_t* offs,
                               uint64_t npat, uint64_t limit, uint64_t* out_offs,
                               uint64_t* out_pos, uint64_t cap, int32_t* status,
                               uint64_t* total);
/* Device variant. Synchronises `stream` once (the total is needed to check cap). */
CSFM_API int csfm_locate_batch_device(csfm_index* idx, const uint8_t* d_bytes,
                                      const uint64_t* d_offs, uint64_t npat, uint64_t limit,
                                      uint64_t* d_out_offs, uint64_t* d_out_pos, uint64_t cap,
                                      int32_t* d_status, uint64_t* total, void* stream);

/* ---- accounting -------------------------------------------------------------------------- */
/* Enable per-call instrumentation (bit 0: count executed steps; bit 1: time the dominant
 * kernel with CUDA events). Off by default; costs one extra reduction / two events. */
CSFM_API int csfm_set_instrumentation(csfm_index* idx, uint32_t mask);
CSFM_API int csfm_last_call_stats(const csfm_index* idx, csfm_call_stats* out);

/* ---- pinned host memory for batch buffers ------------------------------------------------- */
CSFM_API int csfm_host_alloc(void** p, uint64_t bytes);
CSFM_API int csfm_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* CSFM_H */
