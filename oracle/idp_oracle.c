/*
 * idp_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of fg-idprimer's IdentifyPrimers matching chain.  See
 * idp_oracle.h for the parity status ("UNPINNED" items) and citation keys.
 * Every function names the reference lines it follows.  It is written to be
 * read next to the Scala, not to be fast: strings stay ASCII, k-mers stay
 * strings, the aligner fills three score + three trace matrices and walks
 * the traceback exactly as described for fgbio's Aligner.
 */
#define _GNU_SOURCE
#include "idp_oracle.h"
#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * htsjdk 2.16.0 SequenceUtil (un-vendored; build.sbt:128).  Pinned by PrimerMatcherTest.scala:161-185.
 * ---------------------------------------------------------------------------------------------- */
static unsigned char IUPAC_MASK[256];
static pthread_once_t tables_once = PTHREAD_ONCE_INIT;

static void init_tables(void) {
    memset(IUPAC_MASK, 0, sizeof IUPAC_MASK); /* NON_IUPAC_CODE */
    const int A = 1, C = 2, G = 4, T = 8;
    IUPAC_MASK['A'] = A; IUPAC_MASK['C'] = C; IUPAC_MASK['G'] = G; IUPAC_MASK['T'] = T;
    IUPAC_MASK['M'] = A | C; IUPAC_MASK['R'] = A | G; IUPAC_MASK['W'] = A | T;
    IUPAC_MASK['S'] = C | G; IUPAC_MASK['Y'] = C | T; IUPAC_MASK['K'] = G | T;
    IUPAC_MASK['V'] = A | C | G; IUPAC_MASK['H'] = A | C | T; IUPAC_MASK['D'] = A | G | T;
    IUPAC_MASK['B'] = C | G | T; IUPAC_MASK['N'] = A | C | G | T;
    for (int c = 'A'; c <= 'Z'; ++c) IUPAC_MASK[c + 32] = IUPAC_MASK[c];
    IUPAC_MASK['.'] = A | C | G | T;
}

/* SequenceUtil.readBaseMatchesRefBaseWithAmbiguity (call site PM:143): read mask must be a subset of ref mask */
static inline int read_base_matches_ref_base_with_ambiguity(unsigned char read_base, unsigned char ref_base) {
    return (IUPAC_MASK[read_base] & IUPAC_MASK[ref_base]) == IUPAC_MASK[read_base];
}

/* SequenceUtil.complement: htsjdk 2.16 complements only ACGTacgt, everything else is returned unchanged */
static inline char complement(char b) {
    switch (b) {
        case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        default: return b;
    }
}
static void reverse_complement(const char *in, int n, char *out) {
    for (int i = 0; i < n; ++i) out[i] = complement(in[n - 1 - i]);
}

/* ------------------------------------------------------------------------------------------------
 * Scala 2.12 collection-order machinery behind PM:78-91 / PM:112-114 (UNPINNED).
 * ---------------------------------------------------------------------------------------------- */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static uint32_t java_string_hash(const char *s) {
    uint32_t h = 0;
    for (; *s; ++s) h = 31u * h + (unsigned char)*s;
    return h;
}
/* scala.runtime.Statics.mix / finalizeHash (MurmurHash3) as used by the synthetic case-class hashCode */
static uint32_t statics_mix(uint32_t hash, uint32_t data) {
    uint32_t k = data * 0xcc9e2d51u;
    k = rotl32(k, 15) * 0x1b873593u;
    uint32_t h = rotl32(hash ^ k, 13);
    return h * 5u + 0xe6546b64u;
}
static uint32_t statics_finalize(uint32_t h, uint32_t len) {
    h ^= len; h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
/* Primer case class (PR:182-188): fields pair_id, primer_id, sequence, positive_strand, ref_name, start, end */
static uint32_t primer_case_class_hash(const orc_primer *p) {
    uint32_t acc = 0xcafebabeu;
    acc = statics_mix(acc, java_string_hash(p->pair_id));
    acc = statics_mix(acc, java_string_hash(p->primer_id));
    acc = statics_mix(acc, java_string_hash(p->sequence));
    acc = statics_mix(acc, p->positive_strand ? 1231u : 1237u);
    acc = statics_mix(acc, java_string_hash(p->ref_name));
    acc = statics_mix(acc, (uint32_t)p->start);
    acc = statics_mix(acc, (uint32_t)p->end);
    return statics_finalize(acc, 7);
}
/* scala.util.hashing.byteswap32 */
static uint32_t byteswap32(uint32_t v) {
    uint32_t hc = v * 0x9e3775cdu;
    hc = __builtin_bswap32(hc);
    return hc * 0x9e3775cdu;
}
/* mutable.FlatHashTable.index (2.12): improve(hcode, seed) then take the top bits */
static uint32_t flat_index(uint32_t hcode, uint32_t table_len) {
    uint32_t ones = table_len - 1;
    int seed = __builtin_popcount(ones);
    uint32_t improved = byteswap32(hcode);
    int rotation = seed % 32;
    uint32_t rotated = rotation ? ((improved >> rotation) | (improved << (32 - rotation))) : improved;
    return (rotated >> (32 - __builtin_popcount(ones))) & ones;
}
/* immutable.HashSet.improve (2.12) */
static uint32_t trie_improve(uint32_t hcode) {
    uint32_t h = hcode + ~(hcode << 9);
    h ^= h >> 14;
    h += h << 4;
    return h ^ (h >> 10);
}
/* lexicographic order on 5-bit digits taken from the low end = depth-first order of the hash trie */
static int trie_order_less(uint32_t a, uint32_t b) {
    for (int lvl = 0; lvl < 35; lvl += 5) {
        uint32_t da = (a >> lvl) & 31u, db = (b >> lvl) & 31u;
        if (da != db) return da < db;
    }
    return 0;
}
/* Given primers in insertion order (ids[0..n)), produce the iteration order of
 * `mutable.HashSet[Primer]` filled in that order and then `.toSet` (PM:86-90). */
static void scala_set_iteration_order(const uint32_t *hashes, int *ids, int n) {
    if (n <= 1) return;
    /* 1. mutable.HashSet = FlatHashTable: initial 32 slots, load factor 450/1000 */
    int cap = 32, size = 0;
    int threshold = (int)((long long)cap * 450 / 1000);
    int *table = malloc(sizeof(int) * cap);
    for (int i = 0; i < cap; ++i) table[i] = -1;
    for (int e = 0; e < n; ++e) {
        uint32_t h = flat_index(hashes[ids[e]], cap);
        while (table[h] >= 0) h = (h + 1) % cap;
        table[h] = ids[e];
        if (++size >= threshold) { /* growTable: re-add in old slot order */
            int ncap = cap * 2;
            int *nt = malloc(sizeof(int) * ncap);
            for (int i = 0; i < ncap; ++i) nt[i] = -1;
            for (int i = 0; i < cap; ++i) if (table[i] >= 0) {
                uint32_t g = flat_index(hashes[table[i]], ncap);
                while (nt[g] >= 0) g = (g + 1) % ncap;
                nt[g] = table[i];
            }
            free(table); table = nt; cap = ncap;
            threshold = (int)((long long)cap * 450 / 1000);
        }
    }
    int k = 0;
    for (int i = 0; i < cap; ++i) if (table[i] >= 0) ids[k++] = table[i];
    free(table);
    /* 2. .toSet: Set1..Set4 keep insertion order; the 5th element converts to an immutable.HashSet trie */
    if (n >= 5) { /* stable insertion sort by trie order (full collisions keep insertion order) */
        for (int i = 1; i < n; ++i) {
            int v = ids[i]; uint32_t hv = trie_improve(hashes[v]);
            int j = i - 1;
            while (j >= 0 && trie_order_less(hv, trie_improve(hashes[ids[j]]))) { ids[j + 1] = ids[j]; --j; }
            ids[j + 1] = v;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Panel: primers + the three matchers' immutable state (IP:223-234)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { char *kmer; int n; int *ids; } kmer_entry;
typedef struct { int k; int cap; kmer_entry *slots; } kmer_map; /* PM:78-91 kmerToPrimers */

struct orc_panel {
    int n;
    orc_primer *primers;     /* deep copies */
    int *seq_len;            /* sequence.length */
    int *length;             /* Primer.length (PR:206) */
    uint32_t *hash;
    int max_primer_length;   /* PM:75 */
    orc_params params;
    int *score_matrix;
    kmer_map maps[2];        /* 0: ungapped matcher (-k), 1: gapped matcher (-K) */
    int *pair_of;            /* dense pair index per primer (groupBy(_.pair_id), IP:233) */
};

static uint64_t fnv1a(const char *s, int n) {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < n; ++i) { h ^= (unsigned char)s[i]; h *= 1099511628211ull; }
    return h;
}
static kmer_entry *kmer_lookup(const kmer_map *m, const char *s, int n, int create) {
    if (m->cap == 0) return NULL;
    uint64_t h = fnv1a(s, n) & (uint64_t)(m->cap - 1);
    for (;;) {
        kmer_entry *e = &m->slots[h];
        if (!e->kmer) {
            if (!create) return NULL;
            e->kmer = malloc(n + 1); memcpy(e->kmer, s, n); e->kmer[n] = 0;
            return e;
        }
        if ((int)strlen(e->kmer) == n && memcmp(e->kmer, s, n) == 0) return e;
        h = (h + 1) & (uint64_t)(m->cap - 1);
    }
}
/* PM:78-91 */
static void build_kmer_map(orc_panel *p, kmer_map *m, int k) {
    m->k = k; m->cap = 0; m->slots = NULL;
    if (k <= 0) return; /* case None => Map.empty */
    long total = 0;
    for (int i = 0; i < p->n; ++i) total += p->seq_len[i] >= k ? p->seq_len[i] - k + 1 : 1;
    int cap = 64; while (cap < 2 * total + 2) cap *= 2;
    m->cap = cap; m->slots = calloc(cap, sizeof(kmer_entry));
    for (int i = 0; i < p->n; ++i) {
        const char *s = p->primers[i].sequence; int L = p->seq_len[i];
        /* String.sliding(len): a string shorter than len yields itself once; an empty one yields nothing */
        int nw = L >= k ? L - k + 1 : (L > 0 ? 1 : 0);
        for (int w = 0; w < nw; ++w) {
            int wl = L >= k ? k : L;
            kmer_entry *e = kmer_lookup(m, s + w, wl, 1);
            int dup = 0; /* HashSet.add of an element already present */
            for (int j = 0; j < e->n; ++j) if (e->ids[j] == i) { dup = 1; break; }
            if (!dup) { e->ids = realloc(e->ids, sizeof(int) * (e->n + 1)); e->ids[e->n++] = i; }
        }
    }
    for (int s = 0; s < cap; ++s) if (m->slots[s].kmer) scala_set_iteration_order(p->hash, m->slots[s].ids, m->slots[s].n);
}

orc_panel *orc_panel_create(const orc_primer *primers, int n, const orc_params *params) {
    pthread_once(&tables_once, init_tables);
    orc_panel *p = calloc(1, sizeof *p);
    p->n = n; p->params = *params;
    p->primers = calloc(n, sizeof(orc_primer));
    p->seq_len = calloc(n, sizeof(int)); p->length = calloc(n, sizeof(int));
    p->hash = calloc(n, sizeof(uint32_t)); p->pair_of = calloc(n, sizeof(int));
    int npairs = 0;
    for (int i = 0; i < n; ++i) {
        p->primers[i] = primers[i];
        p->primers[i].pair_id = strdup(primers[i].pair_id);
        p->primers[i].primer_id = strdup(primers[i].primer_id);
        p->primers[i].sequence = strdup(primers[i].sequence);
        p->primers[i].ref_name = strdup(primers[i].ref_name ? primers[i].ref_name : "");
        p->seq_len[i] = (int)strlen(primers[i].sequence);
        /* PR:206 def length = if (ref_name.nonEmpty) end - start + 1 else sequence.length */
        p->length[i] = p->primers[i].ref_name[0] ? primers[i].end - primers[i].start + 1 : p->seq_len[i];
        p->hash[i] = primer_case_class_hash(&p->primers[i]);
        if (i == 0 || p->length[i] > p->max_primer_length) p->max_primer_length = p->length[i];
        int found = -1;
        for (int j = 0; j < i; ++j) if (strcmp(p->primers[j].pair_id, p->primers[i].pair_id) == 0) { found = p->pair_of[j]; break; }
        p->pair_of[i] = found >= 0 ? found : npairs++;
    }
    if (params->score_matrix) {
        p->score_matrix = malloc(sizeof(int) * 128 * 128);
        memcpy(p->score_matrix, params->score_matrix, sizeof(int) * 128 * 128);
        p->params.score_matrix = p->score_matrix;
    }
    build_kmer_map(p, &p->maps[0], params->k_ungapped);
    build_kmer_map(p, &p->maps[1], params->k_gapped);
    return p;
}
void orc_panel_destroy(orc_panel *p) {
    if (!p) return;
    for (int i = 0; i < p->n; ++i) {
        free((char *)p->primers[i].pair_id); free((char *)p->primers[i].primer_id);
        free((char *)p->primers[i].sequence); free((char *)p->primers[i].ref_name);
    }
    for (int m = 0; m < 2; ++m) {
        for (int s = 0; s < p->maps[m].cap; ++s) { free(p->maps[m].slots[s].kmer); free(p->maps[m].slots[s].ids); }
        free(p->maps[m].slots);
    }
    free(p->primers); free(p->seq_len); free(p->length); free(p->hash); free(p->pair_of); free(p->score_matrix); free(p);
}
int orc_primer_length(const orc_panel *p, int i) { return p->length[i]; }
uint32_t orc_primer_scala_hash(const orc_panel *p, int i) { return p->hash[i]; }

/* ------------------------------------------------------------------------------------------------
 * PM:39-45 basesInSequencingOrder
 * ---------------------------------------------------------------------------------------------- */
static void bases_in_sequencing_order(const char *bases, int len, uint16_t flags, char *out) {
    int unmapped = (flags & ORC_F_UNMAPPED) != 0, positive = (flags & ORC_F_REVERSE) == 0;
    if (unmapped || positive) memcpy(out, bases, len);
    else reverse_complement(bases, len, out);
}

/* ------------------------------------------------------------------------------------------------
 * PM:98-116 candidate primers
 * ---------------------------------------------------------------------------------------------- */
int orc_candidates_with_common_kmer(const orc_panel *p, int which, const char *bases, int len, int kmer_len, int *out, int cap) {
    const kmer_map *m = &p->maps[which];
    if (len < kmer_len) return 0;                                        /* PM:109 */
    int end = (p->max_primer_length < len ? p->max_primer_length : len) - kmer_len + 1; /* PM:110 */
    int n_windows = len - kmer_len + 1;                                  /* sliding(k) */
    if (end > n_windows) end = n_windows;
    int n = 0;
    for (int i = 0; i < end; ++i) {                                       /* .take(end) */
        const kmer_entry *e = kmer_lookup(m, bases + i, kmer_len, 0);
        if (!e) continue;
        for (int j = 0; j < e->n; ++j) {                                  /* flatMap in Set iteration order */
            int id = e->ids[j], seen = 0;
            for (int q = 0; q < n; ++q) if (out[q] == id) { seen = 1; break; } /* .distinct keeps first occurrence */
            if (!seen && n < cap) out[n++] = id;
        }
    }
    return n;
}
/* PM:98-103 */
static int primers_for_alignment_tasks(const orc_panel *p, int which, const char *so_bases, int len, int *out) {
    int k = which == 0 ? p->params.k_ungapped : p->params.k_gapped;
    if (k <= 0) { for (int i = 0; i < p->n; ++i) out[i] = i; return p->n; }
    return orc_candidates_with_common_kmer(p, which, so_bases, len, k, out, p->n);
}

/* ------------------------------------------------------------------------------------------------
 * PM:121-148 UngappedAligner
 * ---------------------------------------------------------------------------------------------- */
int orc_num_mismatches(const char *bases, int bases_len, const char *primer, int primer_len, int start, double max_mismatches) {
    pthread_once(&tables_once, init_tables);
    int length = bases_len - start < primer_len ? bases_len - start : primer_len; /* PM:139 */
    int count = 0, primer_index = 0;
    while (primer_index < length && count <= max_mismatches) {                      /* PM:142 */
        if (!read_base_matches_ref_base_with_ambiguity((unsigned char)bases[primer_index + start], (unsigned char)primer[primer_index])) count += 1;
        primer_index += 1;
    }
    return count;
}
int orc_mismatch_align(const orc_panel *p, const char *so_bases, int len, int primer, int *mm_out) {
    int length = len < p->length[primer] ? len : p->length[primer];                /* PM:130 */
    double max_mismatches = length * p->params.max_mismatch_rate;                  /* PM:131 */
    int mm = orc_num_mismatches(so_bases, len, p->primers[primer].sequence, p->seq_len[primer], 0, max_mismatches);
    double mismatch_rate = mm / (double)length;                                    /* PM:133 */
    *mm_out = mm;
    return mismatch_rate <= p->params.max_mismatch_rate;                           /* PM:134 (NaN -> false) */
}

/* ------------------------------------------------------------------------------------------------
 * PM:156-214 LocationBasedPrimerMatcher
 * ---------------------------------------------------------------------------------------------- */
int orc_location_overlaps(const orc_panel *p, int ref_idx, int negative, int ustart, int uend, int *out, int cap) {
    int slop = p->params.slop, n = 0;
    int pos = negative ? uend : ustart;                                            /* PM:180, 186 */
    int lo = pos - slop > 1 ? pos - slop : 1, hi = pos + slop;                     /* PM:181, 187 */
    for (int i = 0; i < p->n; ++i) {
        const orc_primer *pr = &p->primers[i];
        if (!pr->ref_name[0]) continue;                                            /* p.mapped (PM:164) */
        if ((pr->positive_strand != 0) == (negative != 0)) continue;               /* per-strand detector */
        if (pr->start > pr->end) continue;        /* htsjdk OverlapDetector.addLhs drops empty intervals */
        if (pr->ref_idx != ref_idx || ref_idx < 0) continue;                       /* Interval contig */
        if (!(pr->start <= hi && pr->end >= lo)) continue;                         /* getOverlaps */
        int key = negative ? pr->end : pr->start;
        if (abs(key - pos) > slop) continue;                                       /* PM:183, 189 */
        if (n < cap) out[n] = i;
        ++n;
    }
    return n;
}
/* PM:194-213; so_bases are the read's bases in sequencing order */
static void location_find(const orc_panel *p, const char *so_bases, int len, int ref_idx, int ustart, int uend, uint16_t flags, orc_result *out) {
    memset(out, 0, sizeof *out); out->primer = -1;
    if (flags & ORC_F_UNMAPPED) return;
    int negative = (flags & ORC_F_REVERSE) != 0;
    int hits[8];
    int n = orc_location_overlaps(p, ref_idx, negative, ustart, uend, hits, 8);
    int chosen = -1;
    if (n == 1) chosen = hits[0];
    else if (n > 1) { /* PM:201-205: filter(_.positiveStrand == rec.positiveStrand) must leave exactly one */
        int cnt = 0, last = -1;
        int m = n < 8 ? n : 8;
        for (int i = 0; i < m; ++i) if ((p->primers[hits[i]].positive_strand != 0) == !negative) { ++cnt; last = hits[i]; }
        if (n > 8) cnt = 2; /* every hit is same-strand by construction; more than one => give up */
        if (cnt == 1) chosen = last;
    }
    if (chosen < 0) return;
    int mm;
    if (orc_mismatch_align(p, so_bases, len, chosen, &mm)) { out->primer = chosen; out->type = ORC_LOCATION; out->mm = mm; out->next_mm = 0; }
}

/* ------------------------------------------------------------------------------------------------
 * PM:222-249 UngappedAlignmentBasedPrimerMatcher
 * ---------------------------------------------------------------------------------------------- */
void orc_best_ungapped(const orc_panel *p, const int *primers, const int *mms, int n, orc_result *out) {
    memset(out, 0, sizeof *out); out->primer = -1;
    if (n == 0) return;
    /* stable sortBy(mm / primer.length.toDouble).take(2) (PM:240-242) */
    int best = -1, second = -1; double kb = 0, ks = 0;
    for (int i = 0; i < n; ++i) {
        double key = mms[i] / (double)p->length[primers[i]];
        if (best < 0 || key < kb) { second = best; ks = kb; best = i; kb = key; }
        else if (second < 0 || key < ks) { second = i; ks = key; }
    }
    out->primer = primers[best]; out->type = ORC_UNGAPPED; out->mm = mms[best];
    out->next_mm = second >= 0 ? mms[second] : ORC_NEXT_NA;                        /* PM:244-245 */
    if (!(out->mm <= out->next_mm)) out->status = ORC_REQ_MM_LE_NEXT;              /* PMa:102 */
}
static void ungapped_find(const orc_panel *p, const char *so_bases, int len, int *cand, int *mms, orc_result *out) {
    int nc = primers_for_alignment_tasks(p, 0, so_bases, len, cand);               /* PM:232 */
    int na = 0;
    for (int i = 0; i < nc; ++i) { int mm; if (orc_mismatch_align(p, so_bases, len, cand[i], &mm)) { cand[na] = cand[i]; mms[na] = mm; ++na; } }
    orc_best_ungapped(p, cand, mms, na, out);                                      /* PM:233 */
}

/* ------------------------------------------------------------------------------------------------
 * fgbio Aligner (un-vendored, fgbio 0.8.0-8d0176b-SNAPSHOT; call sites IP:364-372, AA:69, PM:275).
 * UNPINNED restatement: three score matrices + three trace matrices indexed [query i][target j].
 * The tie policy is isolated here so it can be corrected the day the jar can be run.
 * ---------------------------------------------------------------------------------------------- */
enum { LEFT = 0, UP = 1, DIAG = 2, DONE = 3 };
#define MIN_START_SCORE (INT_MIN / 2)
/* The switches below are the UNPINNED choices (nothing in the reference tree decides them: its tests only check accept /
 * reject at this boundary).  Defaults = the published fgbio algorithm as restated in SURVEY 3.5.  The product carries the very
 * same switches (fg_idprimer_b200/csrc/idp_internal.h, AlignPolicy); tests/test_align_policy.py builds both with the
 * non-default values and checks parity again, so that correcting a choice is a one-line change on both sides. */
#ifndef IDP_POLICY_LEFT_FROM_UP
#define IDP_POLICY_LEFT_FROM_UP 0      /* 1: a Left gap may also open from the Up matrix */
#endif
#ifndef IDP_POLICY_UP_FROM_LEFT
#define IDP_POLICY_UP_FROM_LEFT 0      /* 1: an Up gap may also open from the Left matrix */
#endif
#ifndef IDP_POLICY_EXTEND_OVER_OPEN
#define IDP_POLICY_EXTEND_OVER_OPEN 0  /* 1: equal scores prefer extending a gap over opening one from Diagonal */
#endif
#ifndef IDP_POLICY_TIE_UP_BEFORE_LEFT
#define IDP_POLICY_TIE_UP_BEFORE_LEFT 0 /* 1: equal scores are tried Diagonal, Up, Left (diagonal move and end cell); default Diagonal, Left, Up */
#endif
static const struct {
    int left_from_up, up_from_left, extend_over_open, tie_up_before_left;
} ALIGN_POLICY = { IDP_POLICY_LEFT_FROM_UP, IDP_POLICY_UP_FROM_LEFT, IDP_POLICY_EXTEND_OVER_OPEN, IDP_POLICY_TIE_UP_BEFORE_LEFT };
int orc_align_policy_bits(void) { return IDP_POLICY_LEFT_FROM_UP | IDP_POLICY_UP_FROM_LEFT << 1 | IDP_POLICY_EXTEND_OVER_OPEN << 2 | IDP_POLICY_TIE_UP_BEFORE_LEFT << 3; }

static __thread int *tl_S[3]; static __thread unsigned char *tl_T[3]; static __thread size_t tl_cells;

static inline int score_pair(const orc_params *pa, unsigned char q, unsigned char t, int *missing) {
    if (pa->score_matrix) { /* IP:252-278: (primer_base, read_base) -> score */
        int s = (q < 128 && t < 128) ? pa->score_matrix[q * 128 + t] : INT_MIN;
        if (s == INT_MIN) { *missing = 1; return 0; } /* JVM: NoSuchElementException */
        return s;
    }
    return q == t ? pa->match_score : pa->mismatch_score; /* Aligner.simpleScoringFunction: byte equality */
}

static int align_impl(const orc_params *pa, int mode, const char *query, int n, const char *target, int m, orc_alignment *out) {
    size_t cells = (size_t)(n + 1) * (m + 1);
    if (cells > tl_cells) {
        for (int d = 0; d < 3; ++d) { free(tl_S[d]); free(tl_T[d]); tl_S[d] = malloc(cells * sizeof(int)); tl_T[d] = malloc(cells); }
        tl_cells = cells;
    }
    int *S[3] = { tl_S[0], tl_S[1], tl_S[2] }; unsigned char *T[3] = { tl_T[0], tl_T[1], tl_T[2] };
    const int W = m + 1, go = pa->gap_open, ge = pa->gap_extend;
    int missing = 0;
#define AT(i, j) ((size_t)(i) * W + (j))
    for (int d = 0; d < 3; ++d) { S[d][AT(0, 0)] = 0; T[d][AT(0, 0)] = DONE; }
    for (int i = 1; i <= n; ++i) { /* down the side: target is empty */
        if (mode == ORC_MODE_LOCAL) { for (int d = 0; d < 3; ++d) { S[d][AT(i, 0)] = 0; T[d][AT(i, 0)] = DONE; } }
        else {
            S[UP][AT(i, 0)] = go + i * ge; T[UP][AT(i, 0)] = UP;
            S[LEFT][AT(i, 0)] = MIN_START_SCORE; T[LEFT][AT(i, 0)] = DONE;
            S[DIAG][AT(i, 0)] = MIN_START_SCORE; T[DIAG][AT(i, 0)] = DONE;
        }
    }
    for (int j = 1; j <= m; ++j) { /* across the top: query is empty */
        if (mode == ORC_MODE_GLOBAL) {
            S[LEFT][AT(0, j)] = go + j * ge; T[LEFT][AT(0, j)] = LEFT;
            S[UP][AT(0, j)] = MIN_START_SCORE; T[UP][AT(0, j)] = DONE;
            S[DIAG][AT(0, j)] = MIN_START_SCORE; T[DIAG][AT(0, j)] = DONE;
        } else { for (int d = 0; d < 3; ++d) { S[d][AT(0, j)] = 0; T[d][AT(0, j)] = DONE; } } /* free leading target gap */
    }
    for (int i = 1; i <= n; ++i) {
        for (int j = 1; j <= m; ++j) {
            { /* Diagonal: from the best of (Diagonal, Left, Up) at (i-1, j-1); ties prefer Diagonal, then Left, then Up */
                int addend = score_pair(pa, (unsigned char)query[i - 1], (unsigned char)target[j - 1], &missing);
                int best = S[DIAG][AT(i - 1, j - 1)], dir = DIAG;
                const int second = ALIGN_POLICY.tie_up_before_left ? UP : LEFT, third = ALIGN_POLICY.tie_up_before_left ? LEFT : UP;
                if (S[second][AT(i - 1, j - 1)] > best) { best = S[second][AT(i - 1, j - 1)]; dir = second; }
                if (S[third][AT(i - 1, j - 1)] > best) { best = S[third][AT(i - 1, j - 1)]; dir = third; }
                int s = best + addend;
                if (mode == ORC_MODE_LOCAL && s < 0) { s = 0; dir = DONE; }
                S[DIAG][AT(i, j)] = s; T[DIAG][AT(i, j)] = (unsigned char)dir;
            }
            { /* Up (consumes a query base): open from Diagonal or extend Up; ties prefer Diagonal */
                int best = S[DIAG][AT(i - 1, j)] + go + ge, dir = DIAG;
                int ext = S[UP][AT(i - 1, j)] + ge;
                if (ext > best || (ALIGN_POLICY.extend_over_open && ext == best)) { best = ext; dir = UP; }
                if (ALIGN_POLICY.up_from_left) { int x = S[LEFT][AT(i - 1, j)] + go + ge; if (x > best) { best = x; dir = LEFT; } }
                S[UP][AT(i, j)] = best; T[UP][AT(i, j)] = (unsigned char)dir;
            }
            { /* Left (consumes a target base): open from Diagonal or extend Left; ties prefer Diagonal */
                int best = S[DIAG][AT(i, j - 1)] + go + ge, dir = DIAG;
                int ext = S[LEFT][AT(i, j - 1)] + ge;
                if (ext > best || (ALIGN_POLICY.extend_over_open && ext == best)) { best = ext; dir = LEFT; }
                if (ALIGN_POLICY.left_from_up) { int x = S[UP][AT(i, j - 1)] + go + ge; if (x > best) { best = x; dir = UP; } }
                S[LEFT][AT(i, j)] = best; T[LEFT][AT(i, j)] = (unsigned char)dir;
            }
        }
    }
    /* end cell: strict '>' scanning j (and i for Local) ascending, directions in (Diagonal, Left, Up) order */
    const int DIR_ORDER[3] = { DIAG, ALIGN_POLICY.tie_up_before_left ? UP : LEFT, ALIGN_POLICY.tie_up_before_left ? LEFT : UP };
    int ci = n, cj = m, cd = DIAG, max_score = MIN_START_SCORE;
    if (mode == ORC_MODE_GLOBAL) {
        for (int k = 0; k < 3; ++k) { int d = DIR_ORDER[k]; if (S[d][AT(n, m)] > max_score) { max_score = S[d][AT(n, m)]; cd = d; } }
    } else if (mode == ORC_MODE_GLOCAL) {
        cj = 1;
        for (int j = 1; j <= m; ++j) for (int k = 0; k < 3; ++k) {
            int d = DIR_ORDER[k];
            if (S[d][AT(n, j)] > max_score) { max_score = S[d][AT(n, j)]; cj = j; cd = d; }
        }
    } else {
        ci = 1; cj = 1;
        for (int i = 1; i <= n; ++i) for (int j = 1; j <= m; ++j) for (int k = 0; k < 3; ++k) {
            int d = DIR_ORDER[k];
            if (S[d][AT(i, j)] > max_score) { max_score = S[d][AT(i, j)]; ci = i; cj = j; cd = d; }
        }
    }
    if (n == 0 || m == 0) { /* degenerate: nothing to scan */
        ci = n; cj = m; cd = DIAG; max_score = 0;
        if (mode != ORC_MODE_LOCAL && m == 0 && n > 0) { cd = UP; max_score = S[UP][AT(n, 0)]; }
        if (mode == ORC_MODE_GLOBAL && n == 0 && m > 0) { cd = LEFT; max_score = S[LEFT][AT(0, m)]; }
    }
    /* traceback: count what the CIGAR would hold */
    int i = ci, j = cj, d = cd, len_t = 0, len_q = 0;
    while (T[d][AT(i, j)] != DONE) {
        int nd = T[d][AT(i, j)];
        if (d == DIAG) { --i; --j; ++len_t; ++len_q; }
        else if (d == LEFT) { --j; ++len_t; }
        else { --i; ++len_q; }
        d = nd;
    }
#undef AT
    out->score = max_score;
    out->query_start = i + 1; out->target_start = j + 1;            /* 1-based */
    out->query_end = out->query_start + len_q - 1;
    out->target_end = out->target_start + len_t - 1;                 /* Alignment.targetEnd */
    return missing;
}
void orc_align(const orc_params *pa, int mode, const char *query, int qlen, const char *target, int tlen, orc_alignment *out) {
    align_impl(pa, mode, query, qlen, target, tlen, out);
}

/* ------------------------------------------------------------------------------------------------
 * PM:305-328 getBestGappedAlignment
 * ---------------------------------------------------------------------------------------------- */
void orc_best_gapped(const int *primers, const int *scores, const int *qlens, const int *tstarts, const int *tends, int n, orc_result *out) {
    memset(out, 0, sizeof *out); out->primer = -1;
    if (n == 0) return;
    int best = -1, second = -1; double kb = 0, ks = 0;
    for (int i = 0; i < n; ++i) { /* stable sortBy(-normalizedScore).take(2); normalizedScore = score / query.length.toDouble */
        double key = -(scores[i] / (double)qlens[i]);
        if (best < 0 || key < kb) { second = best; ks = kb; best = i; kb = key; }
        else if (second < 0 || key < ks) { second = i; ks = key; }
    }
    if (second < 0) { /* PM:313-316 */
        if (scores[best] < 0) return;
        out->multi = 0; out->raw_next = 0; out->q_len_next = 0;
    } else { /* PM:317-325 */
        out->multi = 1; out->raw_next = scores[second]; out->q_len_next = (int16_t)qlens[second];
    }
    out->primer = primers[best]; out->type = ORC_GAPPED;
    out->raw_score = scores[best]; out->q_len = (int16_t)qlens[best];
    out->t_start = (int16_t)tstarts[best]; out->t_end = (int16_t)tends[best];
    if (!(orc_result_score(out) >= orc_result_second(out))) out->status = ORC_REQ_SCORE_GE_SECOND; /* PMa:113 */
    else if (!(tstarts[best] <= tends[best])) out->status = ORC_REQ_START_LE_END;                  /* PMa:114 */
}
double orc_result_score(const orc_result *r) {
    if (r->type != ORC_GAPPED) return 0.0;
    return r->multi ? r->raw_score / (double)r->q_len : (double)r->raw_score;      /* PM:315 vs PM:321 */
}
double orc_result_second(const orc_result *r) {
    if (r->type != ORC_GAPPED || !r->multi) return 0.0;
    return r->raw_next / (double)r->q_len_next;                                    /* PM:322 */
}
/* IP:459-471: filter by minAlignmentScoreRate */
static void finalize_gapped(const orc_panel *p, orc_result *r) {
    if (r->type != ORC_GAPPED) return;
    double rate = orc_result_score(r) / (double)p->length[r->primer];              /* IP:467 */
    if (!(rate >= p->params.min_alignment_score_rate)) { /* IP:468 */
        uint8_t st = r->status; memset(r, 0, sizeof *r); r->primer = -1; r->status = st;
    }
}

/* ------------------------------------------------------------------------------------------------
 * IP:441-452 toPrimeMatchFuture, IP:483-516 matchThreePrime
 * ---------------------------------------------------------------------------------------------- */
typedef struct { /* per-thread scratch */
    char *so; char *rc; int cap;
    int *cand, *mms, *scores, *qlens, *ts, *te;
} scratch;
static void scratch_init(scratch *s, int n_primers) {
    memset(s, 0, sizeof *s);
    s->cand = malloc(sizeof(int) * (n_primers + 1)); s->mms = malloc(sizeof(int) * (n_primers + 1));
    s->scores = malloc(sizeof(int) * (n_primers + 1)); s->qlens = malloc(sizeof(int) * (n_primers + 1));
    s->ts = malloc(sizeof(int) * (n_primers + 1)); s->te = malloc(sizeof(int) * (n_primers + 1));
}
static void scratch_free(scratch *s) { free(s->so); free(s->rc); free(s->cand); free(s->mms); free(s->scores); free(s->qlens); free(s->ts); free(s->te); }
static void scratch_fit(scratch *s, int len) {
    if (len + 1 > s->cap) { s->cap = len + 64; s->so = realloc(s->so, s->cap); s->rc = realloc(s->rc, s->cap); }
}

#define ORC_ERR_SCORE_MISSING 4

/* align a list of primers against one target and pick the best (PM:305-328 + IP:459-471) */
static void gapped_over(const orc_panel *p, int mode, const int *cand, int nc, const char *so, int len, int five_prime,
                        scratch *s, orc_result *out, int64_t *n_align) {
    int missing = 0;
    for (int c = 0; c < nc; ++c) {
        int pr = cand[c];
        const char *target; int tlen;
        if (five_prime) { /* PM:292-293 */
            tlen = p->seq_len[pr] + p->params.slop < len ? p->seq_len[pr] + p->params.slop : len;
            target = so;
        } else { /* IP:502-512: toTarget ignores targetLength; reverse complement of the entire read */
            tlen = len; target = s->rc;
        }
        orc_alignment a;
        missing |= align_impl(&p->params, mode, p->primers[pr].sequence, p->seq_len[pr], target, tlen, &a);
        s->scores[c] = a.score; s->qlens[c] = p->seq_len[pr]; s->ts[c] = a.target_start; s->te[c] = a.target_end;
        if (n_align) ++*n_align;                                                    /* AA:70; IP:424 counts 5' only */
    }
    orc_best_gapped(cand, s->scores, s->qlens, s->ts, s->te, nc, out);
    finalize_gapped(p, out);
    if (missing) { memset(out, 0, sizeof *out); out->primer = -1; out->status = ORC_ERR_SCORE_MISSING; } /* the JVM throws NoSuchElementException (IP:277) */
}

static void five_prime_match(const orc_panel *p, const orc_reads *rd, int r, scratch *s, orc_result *out, int64_t *n_align) {
    const char *bases = rd->bases + rd->off[r]; int len = (int)(rd->off[r + 1] - rd->off[r]);
    uint16_t fl = rd->flags[r];
    scratch_fit(s, len);
    bases_in_sequencing_order(bases, len, fl, s->so);
    location_find(p, s->so, len, rd->ref_idx[r], rd->unclipped_start[r], rd->unclipped_end[r], fl, out);   /* IP:442 */
    if (out->type != ORC_NONE) return;
    ungapped_find(p, s->so, len, s->cand, s->mms, out);                                                      /* .orElse */
    if (out->type != ORC_NONE) return;
    uint8_t st = out->status;
    int nc = primers_for_alignment_tasks(p, 1, s->so, len, s->cand);                                          /* PM:290 */
    gapped_over(p, ORC_MODE_GLOCAL, s->cand, nc, s->so, len, 1, s, out, n_align);                            /* IP:445-450 */
    if (st && !out->status) out->status = st;
}

static void three_prime_match(const orc_panel *p, const orc_reads *rd, int r, const orc_result *five, const orc_result *other_five,
                              scratch *s, orc_result *out) {
    memset(out, 0, sizeof *out); out->primer = -1;
    if (p->params.three_prime < 0) return;                                          /* IP:487 */
    const char *bases = rd->bases + rd->off[r]; int len = (int)(rd->off[r + 1] - rd->off[r]);
    scratch_fit(s, len);
    bases_in_sequencing_order(bases, len, rd->flags[r], s->so);
    reverse_complement(s->so, len, s->rc);                                          /* IP:502-506 */
    int nc = 0;
    if (other_five && other_five->type != ORC_NONE) s->cand[nc++] = other_five->primer;             /* IP:495 */
    else if (five->type != ORC_NONE) {                                              /* IP:496 */
        int me = five->primer;
        for (int i = 0; i < p->n; ++i)
            if (p->pair_of[i] == p->pair_of[me] && (p->primers[i].positive_strand != 0) == (p->primers[me].positive_strand == 0)) s->cand[nc++] = i;
    } else { for (int i = 0; i < p->n; ++i) s->cand[nc++] = i; }                    /* IP:497 */
    gapped_over(p, ORC_MODE_LOCAL, s->cand, nc, s->so, len, 0, s, out, NULL);      /* IP:509-515 */
}

int orc_metric_bin(int template_type, int is_fr, int r1_type, int r2_type) {
    return ((template_type * 2 + (is_fr ? 1 : 0)) * 4 + r1_type) * 4 + r2_type;
}
/* TemplateType.scala:52-63 */
static int template_type(uint16_t f1, int has_r2, uint16_t f2) {
    int m1 = !(f1 & ORC_F_UNMAPPED);
    if (!has_r2) return m1 ? 3 : 4;
    int m2 = !(f2 & ORC_F_UNMAPPED);
    if (m1 && m2) return 0;
    if (m1 != m2) return 1;
    return 2;
}

typedef struct {
    const orc_panel *p; const orc_reads *rd; int32_t n_reads;
    orc_result *five, *three;
    const int32_t *tmpl_start; int32_t n_tmpl; /* first read index of each template */
    volatile int32_t *next_chunk;
    int64_t metrics[160]; int64_t n_align;
} worker_arg;

#define TEMPLATES_PER_THREAD 5000 /* IP:189 */

static void *worker(void *vp) {
    worker_arg *a = vp;
    scratch s; scratch_init(&s, a->p->n);
    for (;;) {
        int32_t c = __sync_fetch_and_add(a->next_chunk, 1);
        int64_t t0 = (int64_t)c * TEMPLATES_PER_THREAD;
        if (t0 >= a->n_tmpl) break;
        int64_t t1 = t0 + TEMPLATES_PER_THREAD < a->n_tmpl ? t0 + TEMPLATES_PER_THREAD : a->n_tmpl;
        for (int64_t t = t0; t < t1; ++t) { /* IP:388-418 */
            int r1 = a->tmpl_start[t];
            int has_r2 = (a->rd->flags[r1] & ORC_F_MATE_NEXT) && r1 + 1 < a->n_reads;
            int r2 = r1 + 1;
            five_prime_match(a->p, a->rd, r1, &s, &a->five[r1], &a->n_align);       /* IP:390 */
            if (has_r2) five_prime_match(a->p, a->rd, r2, &s, &a->five[r2], &a->n_align); /* IP:391 */
            if (a->three) {
                three_prime_match(a->p, a->rd, r1, &a->five[r1], has_r2 ? &a->five[r2] : NULL, &s, &a->three[r1]); /* IP:396 */
                if (has_r2) three_prime_match(a->p, a->rd, r2, &a->five[r2], &a->five[r1], &s, &a->three[r2]);    /* IP:397 */
            }
            int tt = template_type(a->rd->flags[r1], has_r2, has_r2 ? a->rd->flags[r2] : 0);
            int bin = orc_metric_bin(tt, (a->rd->flags[r1] & ORC_F_FR_PAIR) != 0, a->five[r1].type, has_r2 ? a->five[r2].type : ORC_NONE); /* IP:413 */
            a->metrics[bin] += 1;
        }
    }
    scratch_free(&s);
    return NULL;
}

int orc_process_batch(const orc_panel *p, const orc_reads *rd, int32_t n_reads, orc_result *five, orc_result *three,
                      int64_t *metrics160, int64_t *n_align, int n_threads) {
    int32_t *tmpl = malloc(sizeof(int32_t) * (n_reads + 1)); int32_t nt = 0;
    for (int32_t r = 0; r < n_reads;) { tmpl[nt++] = r; r += ((rd->flags[r] & ORC_F_MATE_NEXT) && r + 1 < n_reads) ? 2 : 1; }
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    volatile int32_t next = 0;
    worker_arg *args = calloc(n_threads, sizeof(worker_arg));
    pthread_t *th = malloc(sizeof(pthread_t) * n_threads);
    for (int i = 0; i < n_threads; ++i) {
        args[i].p = p; args[i].rd = rd; args[i].n_reads = n_reads; args[i].five = five;
        args[i].three = p->params.three_prime >= 0 ? three : NULL;
        args[i].tmpl_start = tmpl; args[i].n_tmpl = nt; args[i].next_chunk = &next;
    }
    if (n_threads == 1) worker(&args[0]);
    else { for (int i = 0; i < n_threads; ++i) pthread_create(&th[i], NULL, worker, &args[i]); for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL); }
    for (int i = 0; i < n_threads; ++i) { /* IP:424-426 */
        if (metrics160) for (int b = 0; b < 160; ++b) metrics160[b] += args[i].metrics[b];
        if (n_align) *n_align += args[i].n_align;
    }
    free(args); free(th); free(tmpl);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * IdentifyPrimersMetric.scala:34-89
 * ---------------------------------------------------------------------------------------------- */
void orc_summary_metrics(const int64_t *m, int64_t *o) {
    int64_t by_type[5] = {0}, by_match[4] = {0}, ends[3] = {0}, fr = 0, total = 0, attempts = 0;
    for (int tt = 0; tt < 5; ++tt) for (int f = 0; f < 2; ++f) for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) {
        int64_t c = m[orc_metric_bin(tt, f, a, b)];
        if (!c) continue;
        by_type[tt] += c; total += c;
        by_match[a] += c; by_match[b] += c; attempts += 2 * c;   /* both r1 and r2 types are counted, even for fragments */
        ends[(a != ORC_NONE) + (b != ORC_NONE)] += c;
        if (f) fr += c;
    }
    o[0] = total;                                   /* templates */
    o[1] = by_type[0] + by_type[1] + by_type[2];    /* pairs */
    o[2] = by_type[3] + by_type[4];                 /* fragments */
    o[3] = by_type[0]; o[4] = by_type[1]; o[5] = by_type[2]; /* mapped_pairs, unpaired, unmapped_pairs */
    o[6] = fr;                                      /* fr_pairs */
    o[7] = by_type[3]; o[8] = by_type[4];           /* mapped_fragments, unmapped_fragments */
    o[9] = ends[2]; o[10] = ends[1]; o[11] = ends[0]; /* paired_matches, unpaired_matches, no_paired_matches */
    o[12] = attempts;                               /* match_attempts */
    o[13] = by_match[ORC_LOCATION]; o[14] = by_match[ORC_UNGAPPED]; o[15] = by_match[ORC_GAPPED]; o[16] = by_match[ORC_NONE];
}

/* ------------------------------------------------------------------------------------------------
 * PMa:78-117 info(); Java f"%.2f"
 * ---------------------------------------------------------------------------------------------- */
void orc_java_format_2f(double v, char *buf, int cap) {
    /* java.util.Formatter %.2f: take the shortest decimal digits that round-trip (FloatingDecimal), then round HALF_UP */
    if (isnan(v)) { snprintf(buf, cap, "NaN"); return; }
    if (isinf(v)) { snprintf(buf, cap, v > 0 ? "Infinity" : "-Infinity"); return; }
    char digits[64]; int prec;
    for (prec = 1; prec <= 17; ++prec) { snprintf(digits, sizeof digits, "%.*e", prec - 1, fabs(v)); if (strtod(digits, NULL) == fabs(v)) break; }
    /* digits = d.ddddde[+-]xx */
    char mant[32]; int nm = 0; int exp10 = 0;
    for (char *c = digits; *c; ++c) { if (*c == 'e') { exp10 = atoi(c + 1); break; } if (*c >= '0' && *c <= '9') mant[nm++] = *c; }
    /* value = 0.mant * 10^(exp10+1); build fixed-point digit string with enough fraction digits */
    char ip[400], fp[400]; int ni = 0, nf = 0; int point = exp10 + 1; /* number of integer digits */
    for (int i = 0; i < point; ++i) ip[ni++] = i < nm ? mant[i] : '0';
    if (ni == 0) ip[ni++] = '0';
    for (int i = point; i < nm || nf < 3; ++i) { fp[nf++] = i < 0 ? '0' : (i < nm ? mant[i] : '0'); if (nf > 380) break; }
    /* round half up at 2 fraction digits */
    int carry = fp[2] >= '5';
    for (int i = 1; i >= 0 && carry; --i) { if (fp[i] == '9') fp[i] = '0'; else { fp[i]++; carry = 0; } }
    for (int i = ni - 1; i >= 0 && carry; --i) { if (ip[i] == '9') ip[i] = '0'; else { ip[i]++; carry = 0; } }
    int neg = signbit(v) != 0;
    int pos = 0;
    if (neg && pos < cap - 1) buf[pos++] = '-';
    if (carry && pos < cap - 1) buf[pos++] = '1';
    for (int i = 0; i < ni && pos < cap - 1; ++i) buf[pos++] = ip[i];
    if (pos < cap - 1) buf[pos++] = '.';
    for (int i = 0; i < 2 && pos < cap - 1; ++i) buf[pos++] = fp[i];
    buf[pos] = 0;
}

void orc_format_info(const orc_panel *p, const orc_result *r, uint16_t fl, char *buf, int cap) {
    if (r->type == ORC_NONE) { snprintf(buf, cap, "none"); return; }                /* IP:242 */
    const orc_primer *pr = &p->primers[r->primer];
    int read_num = (!(fl & ORC_F_PAIRED) || (fl & ORC_F_FIRST)) ? 1 : 2;             /* PMa:84 */
    const char *name = r->type == ORC_LOCATION ? "Location" : r->type == ORC_UNGAPPED ? "Ungapped" : "Gapped";
    int n = snprintf(buf, cap, "%s,%s,%s:%d-%d,%s,%d,%s", pr->pair_id, pr->primer_id, pr->ref_name, pr->start, pr->end,
                     pr->positive_strand ? "+" : "-", read_num, name);
    if (n >= cap) return;
    if (r->type == ORC_LOCATION) snprintf(buf + n, cap - n, ",%d", r->mm);          /* PMa:97 */
    else if (r->type == ORC_UNGAPPED) {                                             /* PMa:103-106 */
        if (r->next_mm == ORC_NEXT_NA) snprintf(buf + n, cap - n, ",%d,na", r->mm);
        else snprintf(buf + n, cap - n, ",%d,%d", r->mm, r->next_mm);
    } else {                                                                        /* PMa:115 */
        char a[64], b[64];
        orc_java_format_2f(orc_result_score(r), a, sizeof a); orc_java_format_2f(orc_result_second(r), b, sizeof b);
        snprintf(buf + n, cap - n, ",%s,%s,%d,%d", a, b, r->t_start, r->t_end);
    }
}
