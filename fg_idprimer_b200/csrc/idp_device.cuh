// Device-side data model and helpers shared by every kernel of the IdentifyPrimers matching path (sm_100a).
#pragma once
#include <cstddef>

#include <cuda_runtime.h>

#include "idp_internal.h"

namespace idp {

struct DevReads {
    const uint8_t *seq4;
    const uint64_t *seq_off;
    const int32_t *len;
    const int32_t *ref_idx;
    const int32_t *ustart;
    const int32_t *uend;
    const uint16_t *flags;
    int32_t fixed_len;
    int32_t n;
};

// device-side counters of one batch (one per stream slot)
struct Counters {
    unsigned int n_fall;          // reads that fell through to the gapped stage
    unsigned int n_parked;        // reads scan_fx_kernel left to scan_kernel; MUST follow n_fall (one 64-bit atomic bumps both)
    unsigned int n_pairs;         // (read, primer) pairs emitted for sw_pairs_kernel
    unsigned int n_all;           // 3' reads that take every primer
    unsigned int n_defer;         // parked reads whose one accepted primer still needs the exact k-mer walk (member_check_kernel)
    unsigned int n_spill;         // fall-through reads with more -K candidates than the per-thread list holds
    unsigned int overflow;        // pair buffer too small: number of work items not emitted.  FIRST of the sticky fields: everything above is per batch
    unsigned long long n_align5;  // alignments performed by the 5' aligner (IP:424)
    unsigned long long n_align3;
    unsigned long long sw_cells;
    unsigned long long base_compares;
    unsigned long long metrics[160];
    unsigned long long classes[6];  // templates per IDP_CLASS_* (classify_kernel; only templates whose first record is read 1)
};
static_assert(offsetof(Counters, n_parked) == offsetof(Counters, n_fall) + 4 && offsetof(Counters, n_fall) % 8 == 0, "n_fall / n_parked share one 64-bit word");

// the winner (and runner-up) of one read's gapped candidates, before traceback
struct Fin {
    int32_t primer, raw, raw_next;
    int16_t qlen, qlen_next;
    int32_t count;  // number of candidates aligned; 0 = none
};

__device__ __forceinline__ void read_geom(const DevReads &R, int r, const uint8_t *&s, int &len) {
    if (R.fixed_len > 0) { len = R.fixed_len; s = R.seq4 + (size_t)r * (size_t)((R.fixed_len + 1) >> 1); }
    else { len = R.len[r]; s = R.seq4 + R.seq_off[r]; }
}
__device__ __forceinline__ uint32_t stored_nib(const uint8_t *s, int i) {
    uint32_t b = __ldg(s + (i >> 1));
    return (i & 1) ? (b & 15u) : (b >> 4);
}
// htsjdk 2.16 SequenceUtil.complement touches only A<->T, C<->G (nibbles 1<->8, 2<->4); every other code is kept (PM:42)
__device__ __forceinline__ uint32_t comp_nib(uint32_t x) { return (uint32_t)(0xFEDCBA9176523480ull >> (4 * x)) & 15u; }

// 8 BAM bytes-order nibbles (first base in the HIGH nibble of each byte) -> base i in nibble i
__device__ __forceinline__ uint32_t nibswap(uint32_t x) { return ((x & 0x0F0F0F0Fu) << 4) | ((x >> 4) & 0x0F0F0F0Fu); }
// 1 in the low bit of every nibble that is non-zero
__device__ __forceinline__ uint32_t nib_nonzero(uint32_t x) {
    x |= x >> 1;
    x |= x >> 2;
    return x & 0x11111111u;
}
// bits 0,4,...,28 -> bits 0..7
__device__ __forceinline__ uint32_t gather_nibble_bits(uint32_t z) {
    z = (z | (z >> 3)) & 0x03030303u;
    z = (z | (z >> 6)) & 0x000F000Fu;
    return (z | (z >> 12)) & 0xFFu;
}

// a read viewed in a given orientation: forward = as stored, rc = reverse-complemented
struct Oriented {
    const uint8_t *s;
    int len;
    bool rc;
    __device__ __forceinline__ uint32_t nib(int j) const { return rc ? comp_nib(stored_nib(s, len - 1 - j)) : stored_nib(s, j); }
    // same through a generic pointer (the bases may sit in shared memory)
    __device__ __forceinline__ uint32_t nib_generic(int j) const {
        const int i = rc ? len - 1 - j : j;
        const uint32_t b = s[i >> 1];
        const uint32_t v = (i & 1) ? (b & 15u) : (b >> 4);
        return rc ? comp_nib(v) : v;
    }
};
// sequential access to an oriented read through a one-word cache: one 32-bit load per 8 bases instead of a byte load per base
struct TargetStream {
    const uint8_t *s;
    int len;
    bool rc;
    const uint32_t *curp;
    uint32_t word;
    __device__ __forceinline__ explicit TargetStream(const Oriented &o) : s(o.s), len(o.len), rc(o.rc), curp(nullptr), word(0) {}
    __device__ __forceinline__ uint32_t nib(int j) {
        const int i = rc ? len - 1 - j : j;
        const uintptr_t a = reinterpret_cast<uintptr_t>(s) + (uintptr_t)(i >> 1);
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
        if (wp != curp) { curp = wp; word = __ldg(wp); }
        const uint32_t b = (word >> (8 * (int)(a & 3))) & 0xFFu;
        const uint32_t v = (i & 1) ? (b & 15u) : (b >> 4);
        return rc ? comp_nib(v) : v;
    }
};
// sequencing order (PM:39-45): raw when unmapped or positive strand, else reverse complement
__device__ __forceinline__ bool seq_order_is_rc(uint16_t fl) { return !(fl & IDP_F_UNMAPPED) && (fl & IDP_F_REVERSE); }

// where a kernel writes (and cand3_kernel reads) the per-read PrimerMatch records: idp_result (32 B) or idp_result16
struct ResultOut {
    void *p;
    int32_t compact;  // 1: idp_result16
};
__device__ __forceinline__ uint4 pack_result16(const idp_result &v) {
    uint4 w;
    w.x = ((uint32_t)(v.primer + 1) & 0xFFFFFFu) | ((uint32_t)v.type << 24) | ((uint32_t)(v.multi != 0) << 26) | ((uint32_t)v.status << 27);
    const bool gapped = v.type == IDP_GAPPED;
    const int a = gapped ? v.raw_score : v.mm;
    const int b = gapped ? v.raw_next : (v.next_mm == IDP_NEXT_NA ? IDP_NEXT16_NA : v.next_mm);
    w.y = ((uint32_t)a & 0xFFFFu) | ((uint32_t)b << 16);
    w.z = ((uint32_t)(uint16_t)v.q_len) | ((uint32_t)(uint16_t)v.q_len_next << 16);
    w.w = ((uint32_t)(uint16_t)v.t_start) | ((uint32_t)(uint16_t)v.t_end << 16);
    return w;
}
__device__ __forceinline__ void write_result(const ResultOut &o, int r, const idp_result &v) {
    if (o.compact) {
        reinterpret_cast<uint4 *>(o.p)[r] = pack_result16(v);
    } else {
        const uint4 *src = reinterpret_cast<const uint4 *>(&v);
        uint4 *d = reinterpret_cast<uint4 *>(o.p) + 2 * (size_t)r;
        d[0] = src[0];
        d[1] = src[1];
    }
}
// match type and primer of a record already written (what matchThreePrime needs of the 5' results, IP:490-499)
__device__ __forceinline__ void read_type_primer(const ResultOut &o, int r, int &type, int &primer) {
    if (o.compact) {
        const uint32_t h = reinterpret_cast<const uint32_t *>(o.p)[4 * (size_t)r];
        type = (int)((h >> 24) & 3u);
        primer = (int)(h & 0xFFFFFFu) - 1;
    } else {
        const idp_result *q = reinterpret_cast<const idp_result *>(o.p) + r;
        type = q->type;
        primer = q->primer;
    }
}
__device__ __forceinline__ idp_result none_result() {
    idp_result r;
    r.primer = -1; r.type = IDP_NONE; r.status = 0; r.multi = 0; r.reserved = 0;
    r.mm = 0; r.next_mm = 0; r.raw_score = 0; r.raw_next = 0; r.q_len = 0; r.q_len_next = 0; r.t_start = 0; r.t_end = 0;
    return r;
}

// ---- read window loaders: basesInSequencingOrder (PM:39-45), only the prefix the matchers can look at ----
// reverse the 8 nibbles of a word
__device__ __forceinline__ uint32_t rev8nib(uint32_t x) { return nibswap(__byte_perm(x, 0u, 0x0123)); }
// htsjdk 2.16 SequenceUtil.complement on 8 codes at once: A<->T, C<->G (a nibble with exactly one bit set is bit-reversed),
// every other code is kept (PM:42)
__device__ __forceinline__ uint32_t comp8nib(uint32_t x) {
    const uint32_t c = (x & 0x55555555u) + ((x >> 1) & 0x55555555u);
    const uint32_t d = (c & 0x33333333u) + ((c >> 2) & 0x33333333u);   // bits set per nibble
    const uint32_t one = (~nib_nonzero(d ^ 0x11111111u) & 0x11111111u) * 15u;  // nibbles holding exactly one bit
    const uint32_t rev = ((x & 0x11111111u) << 3) | ((x & 0x22222222u) << 1) | ((x >> 1) & 0x22222222u) | ((x >> 3) & 0x11111111u);
    return (x & ~one) | (rev & one);
}
template <int RW>
__device__ __forceinline__ void mask_window(uint32_t (&win)[RW], int len) {
    if (len < RW * 8) {
#pragma unroll
        for (int w = 0; w < RW; ++w) {
            const int rem = len - w * 8;
            if (rem < 8) win[w] = rem <= 0 ? 0u : (win[w] & ((1u << (4 * rem)) - 1u));
        }
    }
}
// wp: 4-byte aligned word pointer at or below the read's first byte (shared or global), bo: the read's byte offset from it
// (0..3).  The caller guarantees that RW + 1 words (forward) / the words around the read's end (reverse, len >= 8 * RW) may be
// read.  Forward: the first 8 * RW bases.  Reverse: the reverse complement of the last 8 * RW stored bases.
template <int RW>
__device__ __forceinline__ void window_from_words(const uint32_t *wp, int bo, int len, bool rc, uint32_t (&win)[RW]) {
    if (!rc) {
        const int sh = bo * 8;
        uint32_t prev = wp[0];
#pragma unroll
        for (int w = 0; w < RW; ++w) {
            const uint32_t next = wp[w + 1];
            win[w] = nibswap(__funnelshift_r(prev, next, sh));
            prev = next;
        }
        mask_window<RW>(win, len);
    } else {
        // stored base k sits at nibble 2 * bo + k of the stream whose word t is nibswap(wp[t]); window word w wants stored
        // bases len - 8 - 8w .. len - 1 - 8w, reversed and complemented
        const int n0 = 2 * bo + len - 8;
        const int t0 = n0 >> 3, sh = 4 * (n0 & 7);
        uint32_t hi = nibswap(wp[t0 + 1]);
#pragma unroll
        for (int w = 0; w < RW; ++w) {
            const uint32_t lo = nibswap(wp[t0 - w]);
            win[w] = comp8nib(rev8nib(__funnelshift_r(lo, hi, sh)));
            hi = lo;
        }
    }
}
// the slow, always safe form: base by base through a generic pointer
template <int RW>
__device__ __forceinline__ void window_by_bases(const uint8_t *sp, int len, bool rc, uint32_t (&win)[RW]) {
    Oriented o{sp, len, rc};
    const int nb = min(len, RW * 8);
#pragma unroll
    for (int w = 0; w < RW; ++w) {
        uint32_t v = 0;
        for (int b = 0; b < 8; ++b) {
            const int i = w * 8 + b;
            if (i < nb) v |= o.nib_generic(i) << (4 * b);
        }
        win[w] = v;
    }
}
// from global memory.  The word form reads up to 4 * RW + 7 bytes from the aligned address below the read (forward) or up
// to 7 bytes past its last byte (reverse): allowed when the batch is fixed-length (rows are contiguous) and the read is not
// among the last few, or when the read itself is long enough (forward only)
template <int RW>
__device__ __forceinline__ void window_from_global(const DevReads &R, int r, const uint8_t *s, int len, bool rc, uint32_t (&win)[RW]) {
    const int bytes = (len + 1) >> 1;
    const bool roomy = R.fixed_len > 0 && (size_t)(R.n - 1 - r) * (size_t)bytes >= (size_t)(4 * RW + 8);
    const bool word_ok = rc ? (roomy && len >= 8 * RW) : (roomy || bytes >= 4 * RW + 4);
    if (word_ok) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(s);
        window_from_words<RW>(reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3)), (int)(a & 3), len, rc, win);
    } else {
        window_by_bases<RW>(s, len, rc, win);
    }
}

__device__ __forceinline__ bool kmer_lookup(const DevKmerIndex &ix, uint64_t key, OffCnt &oc) {
    if (key == 0) return false;
    uint32_t h = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> ix.shift);
    for (;;) {
        const uint64_t kk = __ldg(&ix.keys[h]);
        if (kk == key) { oc = ix.off_cnt[h]; return true; }
        if (kk == 0) return false;
        h = (h + 1) & ix.slot_mask;
    }
}

}  // namespace idp
