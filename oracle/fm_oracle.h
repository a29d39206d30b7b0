/* oracle/fm_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's FM-index hot path
 * (cs::FMIndex::{build_from_text,count,locate} -> cs::WaveletTree::rank -> cs::BitVector::rank1).
 * It is the checker for the CUDA product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load it. The product (libcsfm.so) never links or
 * calls it. Parity status: PINNED — tests/test_oracle_vs_reference.py checks every function here
 * against the compiled reference (oracle/_ref/libcsref.so) and against the reference's own
 * known-answer vectors (tests/golden/).
 *
 * All file:line citations are relative to /root/reference/.
 */
#ifndef FM_ORACLE_H
#define FM_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/core/bitvector.hpp:94-99 — packed words + two-level directory (2048 / 256 bits). */
typedef struct {
  uint64_t nbits, nwords, nsuper, nsub;
  uint64_t* bits;   /* LSB-first packing, bitvector.cpp:24-33 */
  uint32_t* super_; /* absolute rank1 every 2048 bits, bitvector.cpp:49-51 */
  uint16_t* sub;    /* relative rank1 every 256 bits, bitvector.cpp:62-63 */
  uint64_t ones;    /* == count_ones(); cached so rank1(i>=nbits) is O(1) (same value) */
} orc_bitvec;

/* src/core/wavelet.hpp:55-58 — 8 levels, MSB first (a wavelet MATRIX, wavelet.cpp:23-52). */
typedef struct {
  uint64_t n;
  orc_bitvec lv[8];
} orc_wavelet;

/* src/api/fm_index.hpp:40-46 */
typedef struct {
  uint64_t n;
  uint8_t* text; /* may be NULL (from_bwt) */
  uint8_t* bwt;
  uint32_t* sa; /* may be NULL (from_bwt) */
  uint32_t C[257];
  orc_wavelet wt;
  uint32_t stride;
  uint32_t* ssa;
  uint64_t nsamp;
} orc_index;

/* status codes shared with include/csfm.h */
#define ORC_OK 0
#define ORC_LF_WALK_EXCEEDED 1 /* fm_index.cpp:136-138 */
#define ORC_SSA_OOB 2          /* fm_index.cpp:141-146 */

/* ---- construction -------------------------------------------------------------------- */
/* Suffix array in the order of build_sa_naive (src/core/sais.hpp:8-16): unsigned-byte
 * lexicographic, a proper prefix sorts before the longer suffix. Returns 0, or -1 on OOM. */
int orc_sa_build(const uint8_t* text, uint64_t n, uint32_t* sa);
/* O(n) certificate that sa is THE suffix array of text in that order. 0 = valid. */
int orc_sa_check(const uint8_t* text, uint64_t n, const uint32_t* sa);
/* src/core/bwt.hpp:7-15 */
void orc_bwt_from_sa(const uint8_t* text, uint64_t n, const uint32_t* sa, uint8_t* bwt);
/* src/api/fm_index.cpp:36-47 */
void orc_build_C(const uint8_t* bwt, uint64_t n, uint32_t C[257]);

void orc_bv_build(orc_bitvec* bv, const uint8_t* bits01, uint64_t n);             /* bitvector.cpp:14-92 */
void orc_bv_build_from_words(orc_bitvec* bv, const uint64_t* w, uint64_t nw, uint64_t nbits); /* :98-159 */
void orc_bv_free(orc_bitvec* bv);
uint64_t orc_bv_rank1(const orc_bitvec* bv, uint64_t i); /* bitvector.cpp:165-230 */
uint8_t orc_bv_get(const orc_bitvec* bv, uint64_t i);    /* bitvector.hpp:45-50 */

void orc_wt_build(orc_wavelet* wt, const uint8_t* seq, uint64_t n); /* wavelet.cpp:14-53 */
void orc_wt_free(orc_wavelet* wt);
uint64_t orc_wt_rank(const orc_wavelet* wt, uint8_t c, uint64_t i); /* wavelet.cpp:59-96 */
uint8_t orc_wt_access(const orc_wavelet* wt, uint64_t i);           /* wavelet.cpp:102-128 */

orc_index* orc_index_build(const uint8_t* text, uint64_t n, uint32_t stride); /* fm_index.cpp:16-69 */
orc_index* orc_index_from_sa(const uint8_t* text, uint64_t n, const uint32_t* sa, uint32_t stride);
orc_index* orc_index_from_bwt(const uint8_t* bwt, uint64_t n, const uint32_t* ssa, uint64_t nsamp,
                              uint32_t stride);
void orc_index_free(orc_index* idx);

/* ---- queries --------------------------------------------------------------------------- */
/* fm_index.cpp:79-101. Optional outputs: interval after the last executed step (normalised to
 * (0,0) for empty patterns / empty results, SURVEY §8a note 8) and number of executed
 * backward-search steps (the S of the roofline model). */
uint64_t orc_count(const orc_index* idx, const uint8_t* pat, uint64_t m, uint64_t* sp_out,
                   uint64_t* ep_out, uint64_t* steps_out);
/* fm_index.hpp:62-66 */
uint64_t orc_LF(const orc_index* idx, uint64_t i);
/* fm_index.cpp:107-157. Returns number of positions (SA-row order) written to out (cap must be
 * >= min(count,limit)); *status = ORC_* (on error the reference throws: result count 0).
 * lf_steps_out (optional) accumulates LF steps walked. */
uint64_t orc_locate(const orc_index* idx, const uint8_t* pat, uint64_t m, uint64_t limit,
                    uint64_t* out, uint64_t cap, int32_t* status, uint64_t* lf_steps_out);

/* Packed batches: bytes + offs[npat+1]. nthreads <= 0 -> all cores (OpenMP). */
void orc_count_batch(const orc_index* idx, const uint8_t* bytes, const uint64_t* offs,
                     uint64_t npat, uint64_t* counts, uint64_t* sp_ep /*2*npat or NULL*/,
                     uint64_t* steps /*npat or NULL*/, int nthreads);
/* out_offs[npat+1] = exclusive prefix of min(count,limit) (0 for failed queries' *written*
 * positions is NOT applied: slots are reserved by count, like the device engine); returns
 * total slots. out_pos may be NULL to size the buffer first. */
uint64_t orc_locate_batch(const orc_index* idx, const uint8_t* bytes, const uint64_t* offs,
                          uint64_t npat, uint64_t limit, uint64_t* out_offs, uint64_t* out_pos,
                          uint64_t cap, int32_t* status, uint64_t* lf_steps_total, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
