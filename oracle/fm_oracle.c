/* oracle/fm_oracle.c — TEST INFRASTRUCTURE ONLY (see fm_oracle.h for the rules).
 *
 * Plain-C restatement of the reference algorithm; each function cites the reference lines it
 * follows (paths relative to /root/reference/). Differences from the reference are limited to:
 *   - rank1(i >= nbits) returns a cached total instead of rescanning the level
 *     (bitvector.cpp:167-170 -> count_ones(), :236-248): same value, O(1);
 *   - the suffix array is built by prefix doubling instead of std::sort on substr copies
 *     (sais.hpp:12-14): same total order, hence the same array (suffixes are distinct).
 * Both are pinned against the compiled reference in tests/test_oracle_vs_reference.py.
 */
#include "fm_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define SUPER 2048u /* include/cs/config.hpp:56-58 CS_SUPER_BLOCK_SIZE */
#define SUB 256u    /* include/cs/config.hpp:61-63 CS_SUB_BLOCK_SIZE */

static inline uint32_t popcount64(uint64_t x) { /* src/util/bitops.hpp:12-26 */
  return (uint32_t)__builtin_popcountll(x);
}

/* ======================================================================================
 * Suffix array (order of src/core/sais.hpp:8-16)
 * ====================================================================================== */

/* Stable LSD radix sort of (key,val) pairs by 16-bit digits; digits on which all keys agree
 * are skipped. Result ends in (k,v); (kt,vt) are scratch. */
static void radix_sort_pairs(uint64_t** k, uint32_t** v, uint64_t** kt, uint32_t** vt, uint64_t n) {
  uint64_t* cnt = (uint64_t*)malloc(65536 * sizeof(uint64_t));
  for (int pass = 0; pass < 4; ++pass) {
    const int sh = pass * 16;
    memset(cnt, 0, 65536 * sizeof(uint64_t));
    const uint64_t* K = *k;
    for (uint64_t i = 0; i < n; ++i) cnt[(K[i] >> sh) & 0xFFFF]++;
    if (n && cnt[(K[0] >> sh) & 0xFFFF] == n) continue; /* constant digit */
    uint64_t sum = 0;
    for (int d = 0; d < 65536; ++d) { uint64_t c = cnt[d]; cnt[d] = sum; sum += c; }
    const uint32_t* V = *v;
    uint64_t* KO = *kt;
    uint32_t* VO = *vt;
    for (uint64_t i = 0; i < n; ++i) {
      const uint64_t p = cnt[(K[i] >> sh) & 0xFFFF]++;
      KO[p] = K[i];
      VO[p] = V[i];
    }
    uint64_t* t1 = *k; *k = *kt; *kt = t1;
    uint32_t* t2 = *v; *v = *vt; *vt = t2;
  }
  free(cnt);
}

int orc_sa_build(const uint8_t* T, uint64_t n, uint32_t* sa_out) {
  if (n == 0) return 0;
  uint64_t* key = (uint64_t*)malloc(n * 8);
  uint64_t* key2 = (uint64_t*)malloc(n * 8);
  uint32_t* sa = (uint32_t*)malloc(n * 4);
  uint32_t* sa2 = (uint32_t*)malloc(n * 4);
  uint32_t* rank = (uint32_t*)malloc(n * 4);
  if (!key || !key2 || !sa || !sa2 || !rank) {
    free(key); free(key2); free(sa); free(sa2); free(rank);
    return -1;
  }
  /* Round 0: first 7 symbols, 9 bits each: code = byte+1, 0 = past the end. The pad code
   * makes a suffix that is a proper prefix of another sort first, which is what
   * std::string::operator< does in sais.hpp:13; bytes compare unsigned (char_traits<char>). */
  const uint64_t K0 = 7;
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t kk = 0;
    for (uint64_t j = 0; j < K0; ++j) {
      const uint64_t code = (i + j < n) ? (uint64_t)T[i + j] + 1 : 0;
      kk = (kk << 9) | code;
    }
    key[i] = kk;
    sa[i] = (uint32_t)i;
  }
  uint64_t h = K0;
  for (;;) {
    radix_sort_pairs(&key, &sa, &key2, &sa2, n);
    /* rank = index of the first member of the equal-key group */
    uint64_t groups = 0;
    uint32_t head = 0;
    for (uint64_t j = 0; j < n; ++j) {
      if (j == 0 || key[j] != key[j - 1]) { head = (uint32_t)j; ++groups; }
      rank[sa[j]] = head;
    }
    if (groups == n) break;
    for (uint64_t j = 0; j < n; ++j) {
      const uint64_t s = sa[j];
      const uint64_t r2 = (s + h < n) ? (uint64_t)rank[s + h] + 1 : 0; /* shorter first */
      key[j] = ((uint64_t)rank[s] << 32) | r2;
    }
    h *= 2;
  }
  memcpy(sa_out, sa, n * 4);
  free(key); free(key2); free(sa); free(sa2); free(rank);
  return 0;
}

int orc_sa_check(const uint8_t* T, uint64_t n, const uint32_t* sa) {
  if (n == 0) return 0;
  uint32_t* isa = (uint32_t*)malloc(n * 4);
  if (!isa) return -1;
  memset(isa, 0xFF, n * 4);
  int bad = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if (sa[i] >= n || isa[sa[i]] != 0xFFFFFFFFu) { bad = 1; break; } /* not a permutation */
    isa[sa[i]] = (uint32_t)i;
  }
  /* n == 2^32-1 would collide with the 0xFFFFFFFF marker; the reference caps n < 2^32 and
   * this oracle is used far below that. */
  for (uint64_t i = 0; !bad && i + 1 < n; ++i) {
    const uint64_t a = sa[i], b = sa[i + 1];
    if (T[a] < T[b]) continue;
    if (T[a] > T[b]) { bad = 2; break; }
    /* equal first byte: compare the suffixes that follow; past-the-end is smallest */
    if (a + 1 == n) continue;          /* a's remainder is empty -> smaller */
    if (b + 1 == n) { bad = 3; break; } /* b's remainder empty but sorted after a */
    if (isa[a + 1] > isa[b + 1]) { bad = 4; break; }
  }
  free(isa);
  return bad;
}

void orc_bwt_from_sa(const uint8_t* T, uint64_t n, const uint32_t* sa, uint8_t* bwt) {
  /* src/core/bwt.hpp:10-13 — cyclic predecessor, no sentinel appended */
  for (uint64_t i = 0; i < n; ++i) {
    const uint32_t idx = sa[i];
    bwt[i] = (idx == 0) ? T[n - 1] : T[idx - 1];
  }
}

void orc_build_C(const uint8_t* bwt, uint64_t n, uint32_t C[257]) {
  /* src/api/fm_index.cpp:36-47 */
  uint32_t freq[256];
  memset(freq, 0, sizeof freq);
  for (uint64_t i = 0; i < n; ++i) freq[bwt[i]]++;
  uint32_t cum = 0;
  for (int c = 0; c < 256; ++c) { C[c] = cum; cum += freq[c]; }
  C[256] = cum;
}

/* ======================================================================================
 * BitVector (src/core/bitvector.cpp)
 * ====================================================================================== */

static void bv_build_directory(orc_bitvec* bv) {
  /* bitvector.cpp:35-92 (identical loop in :108-159) */
  const uint64_t nbits = bv->nbits;
  const uint64_t num_supers = (nbits + SUPER - 1) / SUPER;
  const uint64_t num_subs = (nbits + SUB - 1) / SUB;
  bv->super_ = (uint32_t*)malloc((num_supers ? num_supers : 1) * 4);
  bv->sub = (uint16_t*)malloc((num_subs ? num_subs : 1) * 2);
  bv->nsuper = 0;
  bv->nsub = 0;
  uint64_t running = 0;
  for (uint64_t s = 0; s < num_supers; ++s) {
    bv->super_[bv->nsuper++] = (uint32_t)running;
    const uint64_t s0 = s * SUPER;
    const uint64_t s1 = (s0 + SUPER < nbits) ? s0 + SUPER : nbits;
    uint64_t local = 0;
    for (uint64_t k = 0; k < SUPER / SUB; ++k) {
      const uint64_t b0 = s0 + k * SUB;
      if (b0 >= nbits) break;
      bv->sub[bv->nsub++] = (uint16_t)local;
      const uint64_t b1 = (b0 + SUB < s1) ? b0 + SUB : s1;
      const uint64_t w0 = b0 / 64, w1 = (b1 + 63) / 64;
      for (uint64_t w = w0; w < w1; ++w) {
        uint64_t word = bv->bits[w];
        const uint64_t wb0 = w * 64, wb1 = wb0 + 64;
        if (wb0 < b0) word &= (~0ULL << (b0 - wb0));
        if (wb1 > b1) word &= (~0ULL >> (wb1 - b1));
        const uint32_t pop = popcount64(word);
        local += pop;
        running += pop;
      }
    }
  }
  /* cached count_ones() (bitvector.cpp:236-248) */
  uint64_t total = 0;
  for (uint64_t w = 0; w < bv->nwords; ++w) {
    uint64_t word = bv->bits[w];
    if ((w + 1) * 64 > nbits) {
      const uint64_t valid = nbits - w * 64;
      word &= ((1ULL << valid) - 1);
    }
    total += popcount64(word);
  }
  bv->ones = total;
}

void orc_bv_build(orc_bitvec* bv, const uint8_t* bits01, uint64_t n) {
  memset(bv, 0, sizeof *bv);
  bv->nbits = n;
  if (n == 0) return; /* bitvector.cpp:16-21 */
  bv->nwords = (n + 63) / 64;
  bv->bits = (uint64_t*)calloc(bv->nwords, 8);
  for (uint64_t i = 0; i < n; ++i) /* bitvector.cpp:26-32, LSB = bit 0 */
    if (bits01[i]) bv->bits[i / 64] |= (1ULL << (i % 64));
  bv_build_directory(bv);
}

void orc_bv_build_from_words(orc_bitvec* bv, const uint64_t* w, uint64_t nw, uint64_t nbits) {
  memset(bv, 0, sizeof *bv);
  bv->nbits = nbits;
  const uint64_t need = (nbits + 63) / 64; /* bitvector.cpp:102-106 */
  bv->nwords = nw > need ? nw : need;
  bv->bits = (uint64_t*)calloc(bv->nwords ? bv->nwords : 1, 8);
  if (nw) memcpy(bv->bits, w, nw * 8);
  bv_build_directory(bv);
}

void orc_bv_free(orc_bitvec* bv) {
  free(bv->bits); free(bv->super_); free(bv->sub);
  memset(bv, 0, sizeof *bv);
}

uint8_t orc_bv_get(const orc_bitvec* bv, uint64_t i) {
  if (i >= bv->nbits) return 0; /* bitvector.hpp:46 */
  return (uint8_t)((bv->bits[i / 64] >> (i % 64)) & 1u);
}

uint64_t orc_bv_rank1(const orc_bitvec* bv, uint64_t i) {
  /* bitvector.cpp:165-230 */
  if (i == 0) return 0;
  if (i >= bv->nbits) return bv->ones; /* :167-170, count_ones() value */
  const uint64_t super_idx = i / SUPER;
  uint64_t rank = bv->super_[super_idx];
  const uint64_t super_start = super_idx * SUPER;
  const uint64_t off = i - super_start;
  if (off == 0) return rank; /* :183-185 */
  const uint64_t sub_off = off / SUB;
  const uint64_t block_idx = super_idx * (SUPER / SUB) + sub_off;
  if (block_idx < bv->nsub) rank += bv->sub[block_idx]; /* :193-195 */
  const uint64_t sub_start = super_start + sub_off * SUB;
  if (i == sub_start) return rank; /* :200-202 */
  const uint64_t w0 = sub_start / 64, w1 = (i - 1) / 64;
  for (uint64_t w = w0; w <= w1 && w < bv->nwords; ++w) { /* :208-227 */
    uint64_t word = bv->bits[w];
    const uint64_t wb0 = w * 64;
    if (wb0 < sub_start) word &= (~0ULL << (sub_start - wb0));
    if (wb0 + 64 > i) {
      const uint64_t keep = i - wb0;
      if (keep < 64) word &= ((1ULL << keep) - 1);
    }
    rank += popcount64(word);
  }
  return rank;
}

/* ======================================================================================
 * WaveletTree (src/core/wavelet.cpp) — 8-level wavelet matrix
 * ====================================================================================== */

void orc_wt_build(orc_wavelet* wt, const uint8_t* seq, uint64_t n) {
  memset(wt, 0, sizeof *wt);
  wt->n = n;
  if (n == 0) return; /* wavelet.cpp:16 */
  uint8_t* cur = (uint8_t*)malloc(n);
  uint8_t* nxt = (uint8_t*)malloc(n);
  uint8_t* bitvec = (uint8_t*)malloc(n);
  memcpy(cur, seq, n);
  for (int bit = 7; bit >= 0; --bit) { /* wavelet.cpp:23-52 */
    const int level = 7 - bit;
    uint64_t nz = 0;
    for (uint64_t i = 0; i < n; ++i) {
      bitvec[i] = (cur[i] >> bit) & 1;
      nz += !bitvec[i];
    }
    orc_bv_build(&wt->lv[level], bitvec, n);
    if (bit > 0) { /* stable split: zeros (left) then ones (right), :47-51 */
      uint64_t l = 0, r = nz;
      for (uint64_t i = 0; i < n; ++i) {
        if (bitvec[i]) nxt[r++] = cur[i]; else nxt[l++] = cur[i];
      }
      uint8_t* t = cur; cur = nxt; nxt = t;
    }
  }
  free(cur); free(nxt); free(bitvec);
}

void orc_wt_free(orc_wavelet* wt) {
  for (int l = 0; l < 8; ++l) orc_bv_free(&wt->lv[l]);
  wt->n = 0;
}

uint64_t orc_wt_rank(const orc_wavelet* wt, uint8_t c, uint64_t i) {
  /* wavelet.cpp:59-96 */
  if (i == 0 || i > wt->n) return 0; /* :60 */
  uint64_t start = 0, end = i;
  for (int level = 0; level < 8; ++level) {
    const int bit = 7 - level;
    const orc_bitvec* bv = &wt->lv[level];
    if (((c >> bit) & 1) == 0) { /* :73-78 */
      start = start - orc_bv_rank1(bv, start);
      end = end - orc_bv_rank1(bv, end);
    } else { /* :81-87 */
      const uint64_t r1s = orc_bv_rank1(bv, start), r1e = orc_bv_rank1(bv, end);
      const uint64_t zeros_total = bv->nbits - orc_bv_rank1(bv, bv->nbits);
      start = zeros_total + r1s;
      end = zeros_total + r1e;
    }
    if (start >= end) return 0; /* :91 */
  }
  return end - start;
}

uint8_t orc_wt_access(const orc_wavelet* wt, uint64_t i) {
  /* wavelet.cpp:102-128 */
  if (i >= wt->n) return 0;
  uint8_t symbol = 0;
  uint64_t pos = i;
  for (int level = 0; level < 8; ++level) {
    const int bit = 7 - level;
    const orc_bitvec* bv = &wt->lv[level];
    const uint8_t b = orc_bv_get(bv, pos);
    symbol |= (uint8_t)(b << bit);
    if (b == 0) {
      pos = pos - orc_bv_rank1(bv, pos);
    } else {
      const uint64_t zeros_total = bv->nbits - orc_bv_rank1(bv, bv->nbits);
      pos = zeros_total + orc_bv_rank1(bv, pos);
    }
  }
  return symbol;
}

/* ======================================================================================
 * FMIndex (src/api/fm_index.cpp)
 * ====================================================================================== */

static void fill_ssa(orc_index* idx, const uint32_t* sa) {
  /* fm_index.cpp:57-65 — samples are taken at SA-ROW multiples of the stride */
  idx->nsamp = (idx->n + idx->stride - 1) / idx->stride;
  idx->ssa = (uint32_t*)malloc((idx->nsamp ? idx->nsamp : 1) * 4);
  for (uint64_t i = 0; i < idx->n; ++i)
    if (i % idx->stride == 0) idx->ssa[i / idx->stride] = sa[i];
}

orc_index* orc_index_from_sa(const uint8_t* text, uint64_t n, const uint32_t* sa, uint32_t stride) {
  orc_index* idx = (orc_index*)calloc(1, sizeof *idx);
  idx->n = n;
  idx->stride = stride;
  idx->text = (uint8_t*)malloc(n ? n : 1);
  idx->bwt = (uint8_t*)malloc(n ? n : 1);
  idx->sa = (uint32_t*)malloc((n ? n : 1) * 4);
  if (n) { memcpy(idx->text, text, n); memcpy(idx->sa, sa, n * 4); }
  orc_bwt_from_sa(text, n, sa, idx->bwt);
  orc_build_C(idx->bwt, n, idx->C);
  orc_wt_build(&idx->wt, idx->bwt, n);
  fill_ssa(idx, sa);
  return idx;
}

orc_index* orc_index_build(const uint8_t* text, uint64_t n, uint32_t stride) {
  uint32_t* sa = (uint32_t*)malloc((n ? n : 1) * 4);
  if (orc_sa_build(text, n, sa) != 0) { free(sa); return NULL; }
  orc_index* idx = orc_index_from_sa(text, n, sa, stride);
  free(sa);
  return idx;
}

orc_index* orc_index_from_bwt(const uint8_t* bwt, uint64_t n, const uint32_t* ssa, uint64_t nsamp,
                              uint32_t stride) {
  orc_index* idx = (orc_index*)calloc(1, sizeof *idx);
  idx->n = n;
  idx->stride = stride;
  idx->bwt = (uint8_t*)malloc(n ? n : 1);
  if (n) memcpy(idx->bwt, bwt, n);
  orc_build_C(idx->bwt, n, idx->C);
  orc_wt_build(&idx->wt, idx->bwt, n);
  idx->nsamp = nsamp;
  idx->ssa = (uint32_t*)malloc((nsamp ? nsamp : 1) * 4);
  if (ssa && nsamp) memcpy(idx->ssa, ssa, nsamp * 4);
  return idx;
}

void orc_index_free(orc_index* idx) {
  if (!idx) return;
  free(idx->text); free(idx->bwt); free(idx->sa); free(idx->ssa);
  orc_wt_free(&idx->wt);
  free(idx);
}

uint64_t orc_count(const orc_index* idx, const uint8_t* pat, uint64_t m, uint64_t* sp_out,
                   uint64_t* ep_out, uint64_t* steps_out) {
  /* fm_index.cpp:79-101 */
  uint64_t sp = 0, ep = idx->n, steps = 0, result;
  if (sp_out) *sp_out = 0;
  if (ep_out) *ep_out = 0;
  if (steps_out) *steps_out = 0;
  if (m == 0) return idx->n; /* :80 */
  if (idx->n == 0) return 0; /* :81 */
  result = 0;
  int empty = 0;
  for (uint64_t k = m; k-- > 0;) { /* right to left, :88 */
    const uint8_t c = pat[k];
    sp = idx->C[c] + orc_wt_rank(&idx->wt, c, sp); /* :92 */
    ep = idx->C[c] + orc_wt_rank(&idx->wt, c, ep); /* :93 */
    ++steps;
    if (sp >= ep) { empty = 1; break; } /* :96 */
  }
  if (steps_out) *steps_out = steps;
  if (!empty) {
    result = ep - sp; /* :100 */
    if (sp_out) *sp_out = sp;
    if (ep_out) *ep_out = ep;
  }
  return result;
}

uint64_t orc_LF(const orc_index* idx, uint64_t i) {
  /* fm_index.hpp:62-66 — reads the plain BWT byte */
  if (i >= idx->n) return 0;
  const uint8_t c = idx->bwt[i];
  return idx->C[c] + orc_wt_rank(&idx->wt, c, i);
}

uint64_t orc_locate(const orc_index* idx, const uint8_t* pat, uint64_t m, uint64_t limit,
                    uint64_t* out, uint64_t cap, int32_t* status, uint64_t* lf_steps_out) {
  /* fm_index.cpp:107-157 */
  if (status) *status = ORC_OK;
  if (m == 0 || idx->n == 0) return 0; /* :109 */
  uint64_t sp, ep;
  const uint64_t cnt = orc_count(idx, pat, m, &sp, &ep, NULL); /* same loop as :112-120 */
  if (cnt == 0) return 0;
  uint64_t nout = 0, lf_steps = 0;
  for (uint64_t i = sp; i < ep && nout < limit; ++i) { /* :125 */
    uint64_t pos = i, steps = 0;
    while (pos % idx->stride != 0 && steps < idx->n) { /* :130-133 */
      pos = orc_LF(idx, pos);
      ++steps;
    }
    lf_steps += steps;
    if (steps >= idx->n) { /* :136-138 throws */
      if (status) *status = ORC_LF_WALK_EXCEEDED;
      if (lf_steps_out) *lf_steps_out += lf_steps;
      return 0;
    }
    const uint64_t sample_idx = pos / idx->stride;
    if (sample_idx >= idx->nsamp) { /* :141-146 throws */
      if (status) *status = ORC_SSA_OOB;
      if (lf_steps_out) *lf_steps_out += lf_steps;
      return 0;
    }
    const uint64_t text_pos = ((uint64_t)idx->ssa[sample_idx] + steps) % idx->n; /* :152 */
    if (nout < cap) out[nout] = text_pos;
    ++nout;
  }
  if (lf_steps_out) *lf_steps_out += lf_steps;
  return nout;
}

/* ---- packed batches, sliced over pthreads (queries are independent, idx is read-only) ---- */
typedef struct {
  const orc_index* idx;
  const uint8_t* bytes;
  const uint64_t* offs;
  uint64_t npat, limit, cap;
  uint64_t *counts, *sp_ep, *steps, *out_offs, *out_pos;
  int32_t* status;
  atomic_ullong next, lf_total;
  int mode; /* 0 = count, 1 = locate */
} batch_job;

static void* batch_worker(void* arg) {
  batch_job* j = (batch_job*)arg;
  const uint64_t chunk = j->mode ? 4 : 64;
  for (;;) {
    const uint64_t q0 = atomic_fetch_add(&j->next, chunk);
    if (q0 >= j->npat) break;
    const uint64_t q1 = (q0 + chunk < j->npat) ? q0 + chunk : j->npat;
    for (uint64_t q = q0; q < q1; ++q) {
      const uint8_t* pat = j->bytes + j->offs[q];
      const uint64_t m = j->offs[q + 1] - j->offs[q];
      if (j->mode == 0) {
        uint64_t sp, ep, st;
        j->counts[q] = orc_count(j->idx, pat, m, &sp, &ep, &st);
        if (j->sp_ep) { j->sp_ep[2 * q] = sp; j->sp_ep[2 * q + 1] = ep; }
        if (j->steps) j->steps[q] = st;
      } else {
        const uint64_t o = j->out_offs[q];
        const uint64_t room = (o < j->cap) ? j->cap - o : 0;
        uint64_t lf = 0;
        int32_t st = ORC_OK;
        orc_locate(j->idx, pat, m, j->limit, j->out_pos + o, room, &st, &lf);
        if (j->status) j->status[q] = st;
        atomic_fetch_add(&j->lf_total, lf);
      }
    }
  }
  return NULL;
}

static void run_batch(batch_job* j, int nthreads) {
  if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  int started = 0;
  for (int t = 1; t < nthreads; ++t)
    if (pthread_create(&th[started], NULL, batch_worker, j) == 0) ++started;
  batch_worker(j);
  for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
}

void orc_count_batch(const orc_index* idx, const uint8_t* bytes, const uint64_t* offs,
                     uint64_t npat, uint64_t* counts, uint64_t* sp_ep, uint64_t* steps,
                     int nthreads) {
  batch_job j;
  memset(&j, 0, sizeof j);
  j.idx = idx; j.bytes = bytes; j.offs = offs; j.npat = npat;
  j.counts = counts; j.sp_ep = sp_ep; j.steps = steps; j.mode = 0;
  atomic_init(&j.next, 0);
  atomic_init(&j.lf_total, 0);
  run_batch(&j, nthreads);
}

uint64_t orc_locate_batch(const orc_index* idx, const uint8_t* bytes, const uint64_t* offs,
                          uint64_t npat, uint64_t limit, uint64_t* out_offs, uint64_t* out_pos,
                          uint64_t cap, int32_t* status, uint64_t* lf_steps_total, int nthreads) {
  /* pass 1: slots per query = min(count, limit); locate("") is empty (fm_index.cpp:109) */
  uint64_t* counts = (uint64_t*)malloc((npat ? npat : 1) * 8);
  orc_count_batch(idx, bytes, offs, npat, counts, NULL, NULL, nthreads);
  uint64_t total = 0;
  for (uint64_t q = 0; q < npat; ++q) {
    uint64_t c = (offs[q + 1] == offs[q]) ? 0 : counts[q];
    if (c > limit) c = limit;
    out_offs[q] = total;
    total += c;
  }
  out_offs[npat] = total;
  free(counts);
  if (!out_pos) return total;
  batch_job j;
  memset(&j, 0, sizeof j);
  j.idx = idx; j.bytes = bytes; j.offs = offs; j.npat = npat; j.limit = limit; j.cap = cap;
  j.out_offs = out_offs; j.out_pos = out_pos; j.status = status; j.mode = 1;
  atomic_init(&j.next, 0);
  atomic_init(&j.lf_total, 0);
  run_batch(&j, nthreads);
  if (lf_steps_total) *lf_steps_total = atomic_load(&j.lf_total);
  return total;
}
