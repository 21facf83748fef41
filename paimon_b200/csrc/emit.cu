// emit.cu — the dominant kernel of the merge: per tile and column, stage the k run segments into
// shared memory with 1-D bulk async copies (TMA, cp.async.bulk + mbarrier), resolve every output row
// from its key group's members (select / fold per the plan's op codes), and store the result column
// coalesced.  All random access happens in shared memory; global memory only sees contiguous streams.
//
// Replaces, per output row and column, MergeFunction.add()/getResult() of the reference:
//   DeduplicateMergeFunction.java:47-60, PartialUpdateMergeFunction.java:177-188 (updateNonNullFields),
//   aggregate/AggregateMergeFunction.java:91-101 + FieldAggregator implementations.
//
// Structure per CTA (one tile = two plan tiles, <= 4096 input rows), round-2 design:
//   * a tile walks a list of PASSES: [columns without staged data (sequence number, kind)] [var-len columns: SIZE
//     pass] [fixed-width columns] [var-len columns: COPY pass].  The size pass selects every output row's source
//     member and publishes the tile's byte total per column; the cross-tile prefix (decoupled look-back over tiles
//     taken in ticket order) is only needed by the copy pass, ~40 column passes later, when every resident
//     predecessor has long published — round 1 did both in one pass and every var-len column waited ~12 k cycles
//     for lagging predecessor tiles.
//   * passes are pipelined over two shared-memory stages with full / empty mbarriers instead of block barriers:
//     warp 0 refills a stage (bulk copies + the validity words) as soon as all 16 warps have released it, every warp
//     releases a stage when it is done with its own rows, so warps drift up to one pass apart instead of meeting at
//     a __syncthreads per column.
//   * "newest non-null member wins" (partial-update / deduplicate select) is a bit operation: the tile keeps one
//     bitmap over merged positions of the members that may provide a value (op != NOOP, not older than the newest
//     SET); per column a warp ballots the staged validity of the merged positions its 32 output rows span, ANDs, and
//     every lane takes the highest set bit of its group's window — no per-cell walk, no divergence.
#include <stdio.h>

#include "device_utils.cuh"

namespace pg {

constexpr int kStages = 2;
constexpr int kEmitThreads = 512;
constexpr int kEmitWarps = kEmitThreads / 32;
#define kFlagAgg (1ull << 62)
#define kFlagPrefix (2ull << 62)
#define kValMask ((1ull << 62) - 1)

// per merged position: staged row position (13 bits) | op (2 bits) | first-member-of-group (1 bit)
constexpr uint32_t kPmPosMask = 0x1FFF;
constexpr int kPmOpShift = 13;
constexpr uint32_t kPmHead = 0x8000;

// pass list entries (EmitArgs.col_order): column | phase << 16
enum : int { PH_PLAIN = 0, PH_SIZE = 1, PH_FIXED = 2, PH_COPY = 3 };

struct EmitLayout {
    int rt;              // staged rows capacity per stage (multiple of 32)
    size_t stage_bytes;
    size_t total;
};
__host__ __device__ inline EmitLayout emit_layout(int k, int nv) {
    EmitLayout L;
    L.rt = kTileMax + 64 * k;
    L.stage_bytes = (((size_t)L.rt * 8 + (size_t)L.rt / 32 * 4) + 127) & ~(size_t)127;
    L.total = kStages * L.stage_bytes + (size_t)kTileMax * (2 + 2 + 1 + 1) + kTileMax / 8 + (PG_MAX_RUNS + 1) * 4 * 3 +
              PG_MAX_RUNS * 8 + kStages * PG_MAX_RUNS * 8 + 2 * kStages * 8 + 16 + 34 * 4 + 16 +
              (size_t)(kTileMax + 64 * PG_MAX_RUNS) / 32 * 4 + (size_t)nv * (kEmitWarps * 4 + 4 + 4 + 8) + 128;
    return L;
}

__device__ __forceinline__ bool staged_valid(const uint32_t *vw, int p) { return (vw[p >> 5] >> (p & 31)) & 1; }

template <int W> __device__ __forceinline__ uint64_t lds_fixed(const unsigned char *vals, int p);
template <> __device__ __forceinline__ uint64_t lds_fixed<1>(const unsigned char *v, int p) { return v[p]; }
template <> __device__ __forceinline__ uint64_t lds_fixed<2>(const unsigned char *v, int p) { return ((const uint16_t *)v)[p]; }
template <> __device__ __forceinline__ uint64_t lds_fixed<4>(const unsigned char *v, int p) { return ((const uint32_t *)v)[p]; }
template <> __device__ __forceinline__ uint64_t lds_fixed<8>(const unsigned char *v, int p) { return ((const uint64_t *)v)[p]; }
// output / run column pointers come out of device tables as generic pointers: tell the compiler they are global
// (STG / LDG instead of generic ST / LD).  Only for non-null pointers.
template <typename T> __device__ __forceinline__ T *as_global(T *p) { __builtin_assume(__isGlobal(p)); return p; }
template <int W> __device__ __forceinline__ void stg_fixed(void *d, int64_t row, uint64_t v);
template <> __device__ __forceinline__ void stg_fixed<1>(void *d, int64_t r, uint64_t v) { as_global((uint8_t *)d)[r] = (uint8_t)v; }
template <> __device__ __forceinline__ void stg_fixed<2>(void *d, int64_t r, uint64_t v) { as_global((uint16_t *)d)[r] = (uint16_t)v; }
template <> __device__ __forceinline__ void stg_fixed<4>(void *d, int64_t r, uint64_t v) { as_global((uint32_t *)d)[r] = (uint32_t)v; }
template <> __device__ __forceinline__ void stg_fixed<8>(void *d, int64_t r, uint64_t v) { as_global((uint64_t *)d)[r] = v; }

// partial-update sequence group column: the member the plan kernel marked (merged position), or -1
__device__ __forceinline__ int select_marked_idx(const uint16_t *pm, const uint32_t *gplan, uint32_t bit, int last) {
    int j = last;
    while (true) {
        if (gplan[j] & bit) return j;
        if (pm[j] & kPmHead) return -1;
        --j;
    }
}

__device__ __forceinline__ int group_first(const uint16_t *pm, int last) {
    int j = last;
    while (!(pm[j] & kPmHead)) --j;
    return j;
}

// aggregate engine on a var-len column: fold with the member index as accumulator
__device__ int fold_member_idx(const ColDesc &cd, const uint16_t *pm, const uint8_t *mrun, const uint32_t *vw,
                               const int32_t *offs, const uint8_t *const *cdata, int last, int32_t *err) {
    int acc = -1;
    bool initialized = false;
    for (int j = group_first(pm, last); j <= last; j++) {
        uint32_t e = pm[j];
        int op = (e >> kPmOpShift) & 3;
        if (op == OP_NOOP) continue;
        bool v = staged_valid(vw, e & kPmPosMask);
        int in = v ? j : -1;
        if (op == OP_SET) { acc = in; continue; }
        if (op == OP_RETRACT) {
            if (cd.retract == RT_IGNORE) continue;
            switch (cd.agg) {
                case PG_AGG_LAST_VALUE: acc = -1; break;
                case PG_AGG_LAST_NON_NULL_VALUE: if (v) acc = -1; break;
                case PG_AGG_PRIMARY_KEY: acc = in; break;
                default: atomicCAS(err, KERR_NONE, KERR_AGG_RETRACT); break;
            }
            continue;
        }
        switch (cd.agg) {
            case PG_AGG_LAST_VALUE: case PG_AGG_PRIMARY_KEY: acc = in; break;
            case PG_AGG_LAST_NON_NULL_VALUE: if (v) acc = in; break;
            case PG_AGG_FIRST_VALUE: if (!initialized) { initialized = true; acc = in; } break;
            case PG_AGG_FIRST_NON_NULL_VALUE: if (!initialized && v) { initialized = true; acc = in; } break;
            case PG_AGG_MAX: case PG_AGG_MIN:
                if (acc < 0 || in < 0) { if (acc < 0) acc = in; break; }
                {
                    int pa = pm[acc] & kPmPosMask, pb = e & kPmPosMask;
                    int d = bytes_compare(cdata[mrun[acc]] + offs[pa], offs[pa + 1] - offs[pa],
                                          cdata[mrun[j]] + offs[pb], offs[pb + 1] - offs[pb]);
                    if (cd.agg == PG_AGG_MAX) { if (d < 0) acc = in; }
                    else { if (!(d < 0)) acc = in; }
                }
                break;
            default: break;
        }
    }
    return acc;
}

// strict left fold of a fixed-width column in sequence order (AggregateMergeFunction.java:91-101)
__device__ void fold_fixed(const ColDesc &cd, const uint16_t *pm, const uint32_t *vw, const unsigned char *vals,
                           int last, uint64_t *out_val, bool *out_valid, int32_t *err) {
    const int w = cd.width;
    uint64_t val = 0;
    bool is_valid = false, initialized = false;
    for (int j = group_first(pm, last); j <= last; j++) {
        uint32_t e = pm[j];
        int op = (e >> kPmOpShift) & 3;
        if (op == OP_NOOP) continue;
        int pj = e & kPmPosMask;
        bool v = staged_valid(vw, pj);
        uint64_t in = v ? load_fixed(vals, w, pj) : 0;
        if (op == OP_SET) { val = in; is_valid = v; continue; }
        if (op == OP_RETRACT) {
            if (cd.retract == RT_IGNORE) continue;
            switch (cd.agg) {
                case PG_AGG_SUM:       // FieldSumAgg.retract :87-131, negative :133-163
                    if (!is_valid) { if (v) { val = negate_fixed(cd.type, w, in); is_valid = true; } }
                    else if (v) val = arith(cd.type, w, 1, val, in, err);
                    break;
                case PG_AGG_PRODUCT:
                    if (is_valid && v) val = arith(cd.type, w, 3, val, in, err);
                    break;
                case PG_AGG_LAST_VALUE: is_valid = false; val = 0; break;
                case PG_AGG_LAST_NON_NULL_VALUE: if (v) { is_valid = false; val = 0; } break;
                case PG_AGG_PRIMARY_KEY: val = in; is_valid = v; break;
                default: atomicCAS(err, KERR_NONE, KERR_AGG_RETRACT); break;
            }
            continue;
        }
        switch (cd.agg) {
            case PG_AGG_SUM: case PG_AGG_PRODUCT:
                if (!is_valid || !v) { if (!is_valid) { val = in; is_valid = v; } }
                else val = arith(cd.type, w, cd.agg == PG_AGG_SUM ? 0 : 2, val, in, err);
                break;
            case PG_AGG_MAX: case PG_AGG_MIN:
                if (!is_valid || !v) { if (!is_valid) { val = in; is_valid = v; } }
                else {
                    int d = compare_fixed(cd.type, w, val, in);
                    if (cd.agg == PG_AGG_MAX) { if (d < 0) val = in; }
                    else { if (!(d < 0)) val = in; }
                }
                break;
            case PG_AGG_BOOL_AND: case PG_AGG_BOOL_OR:
                if (!is_valid || !v) { if (!is_valid) { val = in; is_valid = v; } }
                else val = cd.agg == PG_AGG_BOOL_AND ? ((val != 0) && (in != 0)) : ((val != 0) || (in != 0));
                break;
            case PG_AGG_LAST_VALUE: case PG_AGG_PRIMARY_KEY: val = in; is_valid = v; break;
            case PG_AGG_LAST_NON_NULL_VALUE: if (v) { val = in; is_valid = true; } break;
            case PG_AGG_FIRST_VALUE:
                if (!initialized) { initialized = true; val = in; is_valid = v; }
                break;
            case PG_AGG_FIRST_NON_NULL_VALUE:
                if (!initialized && v) { initialized = true; val = in; is_valid = true; }
                break;
            default: break;
        }
    }
    *out_val = val;
    *out_valid = is_valid;
}

// agg(accumulator = a, input = b) of a fixed-width aggregator as a function of both operands, so that
// aggReversed(acc, in) = agg(in, acc) (FieldAggregator.java:40-42) can be evaluated too
__device__ void agg_pair(const ColDesc &cd, uint64_t a, bool av, uint64_t b, bool bv, bool *initialized,
                         uint64_t *out, bool *out_valid, int32_t *err) {
    const int w = cd.width;
    switch (cd.agg) {
        case PG_AGG_SUM: case PG_AGG_PRODUCT:
            if (!av || !bv) { *out = av ? a : b; *out_valid = av || bv; }
            else { *out = arith(cd.type, w, cd.agg == PG_AGG_SUM ? 0 : 2, a, b, err); *out_valid = true; }
            return;
        case PG_AGG_MAX: case PG_AGG_MIN:
            if (!av || !bv) { *out = av ? a : b; *out_valid = av || bv; return; }
            {
                const int d = compare_fixed(cd.type, w, a, b);
                const bool take_b = cd.agg == PG_AGG_MAX ? d < 0 : !(d < 0);
                *out = take_b ? b : a; *out_valid = true;
            }
            return;
        case PG_AGG_BOOL_AND: case PG_AGG_BOOL_OR:
            if (!av || !bv) { *out = av ? a : b; *out_valid = av || bv; return; }
            *out = cd.agg == PG_AGG_BOOL_AND ? ((a != 0) && (b != 0)) : ((a != 0) || (b != 0));
            *out_valid = true;
            return;
        case PG_AGG_LAST_VALUE: case PG_AGG_PRIMARY_KEY: *out = b; *out_valid = bv; return;
        case PG_AGG_LAST_NON_NULL_VALUE: *out = bv ? b : a; *out_valid = bv ? true : av; return;
        case PG_AGG_FIRST_VALUE:
            if (!*initialized) { *initialized = true; *out = b; *out_valid = bv; }
            else { *out = a; *out_valid = av; }
            return;
        case PG_AGG_FIRST_NON_NULL_VALUE:
            if (!*initialized && bv) { *initialized = true; *out = b; *out_valid = true; }
            else { *out = a; *out_valid = av; }
            return;
        default: *out = a; *out_valid = av; return;
    }
}

// Field of a partial-update sequence group that has an aggregate function
// (PartialUpdateMergeFunction.updateWithSequenceGroup :228-244, retractWithSequenceGroup :323-339): every member
// whose group is not empty takes part — in order (agg) when its group sequence is >= the accumulated one, else
// reversed (aggReversed); retract members call retract(); the first record of a key initialises the row.
__device__ void fold_group_agg(const ColDesc &cd, const uint16_t *pm, const uint32_t *gagg, const uint32_t *vw,
                               const unsigned char *vals, int last, uint64_t *out_val, bool *out_valid, int32_t *err) {
    const int w = cd.width, g = cd.group;
    uint64_t val = 0;
    bool is_valid = false, initialized = false;
    for (int j = group_first(pm, last); j <= last; j++) {
        const uint32_t e = pm[j];
        const int op = (e >> kPmOpShift) & 3;
        if (op == OP_NOOP) continue;
        const int pj = e & kPmPosMask;
        const bool v = staged_valid(vw, pj);
        const uint64_t in = v ? load_fixed(vals, w, pj) : 0;
        if (op == OP_SET) { val = in; is_valid = v; continue; }            // initRow / row restart: verbatim
        const uint32_t marks = gagg[j];
        if (op == OP_RETRACT) {
            if (e & kPmHead) { val = in; is_valid = v; }                    // initRow, then the retract itself
            if (!((marks >> g) & 1) || cd.retract == RT_IGNORE) continue;
            switch (cd.agg) {
                case PG_AGG_SUM:
                    if (!is_valid) { if (v) { val = negate_fixed(cd.type, w, in); is_valid = true; } }
                    else if (v) val = arith(cd.type, w, 1, val, in, err);
                    break;
                case PG_AGG_PRODUCT:
                    if (is_valid && v) val = arith(cd.type, w, 3, val, in, err);
                    break;
                case PG_AGG_LAST_VALUE: is_valid = false; val = 0; break;
                case PG_AGG_LAST_NON_NULL_VALUE: if (v) { is_valid = false; val = 0; } break;
                case PG_AGG_PRIMARY_KEY: val = in; is_valid = v; break;
                default: atomicCAS(err, KERR_NONE, KERR_AGG_RETRACT); break;
            }
            continue;
        }
        uint64_t r;
        bool rv;
        if ((marks >> g) & 1) agg_pair(cd, val, is_valid, in, v, &initialized, &r, &rv, err);
        else if ((marks >> (16 + g)) & 1) agg_pair(cd, in, v, val, is_valid, &initialized, &r, &rv, err);
        else continue;                                                       // empty group in this record
        val = rv ? r : 0;
        is_valid = rv;
    }
    *out_val = val;
    *out_valid = is_valid;
}

struct TileView {
    const uint16_t *pm;
    const uint16_t *glast;
    const uint8_t *gcnt;
    const uint32_t *aebits;
    int n;                       // merged positions (input rows) of the tile
    int single;                  // deduplicate / first-row: glast[] = the group's ONE eligible member
    int n_out;
    int o_shift;
    int64_t out_base;
    int64_t in_base;
    int n_out_a;                 // output rows of the first plan tile of the pair
    int seq_shift;               // tmp_seq / tmp_kind are laid out per plan tile: rows of the second one sit
                                 // (n_a - n_out_a) entries further
};
__device__ __forceinline__ int64_t result_index(const TileView &tv, int ob) {
    return tv.in_base + ob + (ob >= tv.n_out_a ? tv.seq_shift : 0);
}

// Select for 32 consecutive output rows at once (one per lane; `active` lanes hold a row): the newest eligible member
// with a non-null cell (PartialUpdateMergeFunction.updateNonNullFields :177-188; deduplicate / first-row have ONE
// eligible member per group, found in the prologue: tv.single).  Returns the winner's merged position or -1 (NULL).
// Must be called by all 32 lanes.
//   default: newest-first walk over the group's members — UPD members with a NULL cell are skipped, a SET member ends
//     the walk.  Short (the newest member usually has the value) and measured faster than the variant below.
//   -DPG_EMIT_BITMAP_SELECT: the warp ballots the staged validity of every merged position its 32 rows span, ANDs it
//     with the tile's eligibility bitmap and every lane takes the highest set bit of its group's window.  No
//     divergence, but ~3 ballot rounds per 32 rows cost more than the walks they replace (profiles/README.md).
template <bool GAGG>
__device__ __forceinline__ int select_idx_warp(const TileView &tv, const uint32_t *vw, bool active, int last, int cnt) {
    if (tv.single) {
        // glast[] holds the group's only eligible member (or a member without an op: no value)
        if (!active) return -1;
        const uint32_t e = tv.pm[last];
        return (((e >> kPmOpShift) & 3) != OP_NOOP && staged_valid(vw, e & kPmPosMask)) ? last : -1;
    }
#ifdef PG_EMIT_BITMAP_SELECT
    const int lane = threadIdx.x & 31;
    const unsigned act = __ballot_sync(0xffffffffu, active);
    if (act == 0) return -1;
    const int first = last - cnt + 1;
    const int a = __shfl_sync(0xffffffffu, first, __ffs(act) - 1);
    const int b = __shfl_sync(0xffffffffu, last, 31 - __clz(act));
    const int mylo = first >> 5;
    uint32_t lo = 0, hi = 0;
    for (int w = a >> 5; w <= (b >> 5); w++) {
        const int i = 32 * w + lane;
        const uint32_t p = (i < tv.n ? tv.pm[i] : 0) & kPmPosMask;
        const unsigned V = __ballot_sync(0xffffffffu, (vw[p >> 5] >> (p & 31)) & 1);
        const uint32_t C = V & tv.aebits[w];
        if (w == mylo) lo = C;
        if (w == mylo + 1) hi = C;
    }
    uint32_t win = __funnelshift_r(lo, hi, first & 31);
    if (cnt < 32) win &= (1u << cnt) - 1;
    if (!active || win == 0) return -1;
    return first + 31 - __clz(win);
#else
    if (!active) return -1;
    int j = last;
    while (true) {
        const uint32_t e = tv.pm[j];
        uint32_t op = (e >> kPmOpShift) & 3;
        // (only with aggregates inside sequence groups does the plan mark retracts on a partial-update merge) a
        // RETRACT member leaves select columns alone, unless it is the group's first record (initRow: verbatim)
        if (GAGG && op == OP_RETRACT) op = (e & kPmHead) ? OP_SET : OP_NOOP;
        if (op != OP_NOOP) {
            if (staged_valid(vw, e & kPmPosMask)) return j;
            if (op == OP_SET) return -1;
        }
        if (e & kPmHead) return -1;
        --j;
    }
#endif
}

// one output validity word per warp iteration: interior words are plain stores, tile-boundary words OR
__device__ __forceinline__ void put_validity_word(uint8_t *validity, int64_t out_base, int wb, int n_out,
                                                  bool bit) {
    unsigned mask = __ballot_sync(0xffffffffu, bit);
    if ((threadIdx.x & 31) == 0) {
        int64_t word = (out_base + wb) >> 5;
        bool full = wb >= 0 && wb + 32 <= n_out;
        uint32_t *bm = as_global((uint32_t *)validity);
        if (full) bm[word] = mask;
        else if (mask) atomicOr(&bm[word], mask);
    }
}

template <int W, bool GAGG>
__device__ __forceinline__ void emit_fixed_column(const EmitArgs &ea, const ColDesc &cd, const pg_out_column &oc,
                                                  const TileView &tv, const unsigned char *vals, const uint32_t *vw) {
    const int tid = threadIdx.x, lane = tid & 31;
    for (int wb = (tid & ~31) - tv.o_shift; wb < tv.n_out; wb += kEmitThreads) {   // warp-uniform trip count
        const int ob = wb + lane;
        const bool active = ob >= 0 && ob < tv.n_out;
        bool is_valid = false;
        uint64_t val = 0;
        const int last = active ? tv.glast[ob] : 0;
        if (cd.mode == CM_SELECT) {
            const int j = select_idx_warp<GAGG>(tv, vw, active, last, active ? tv.gcnt[ob] : 1);
            if (j >= 0) { val = lds_fixed<W>(vals, tv.pm[j] & kPmPosMask); is_valid = true; }
        } else if (active) {
            if (cd.mode == CM_GVAL || cd.mode == CM_GSEQ) {
                const uint32_t gbit = 1u << (cd.agg + (cd.mode == CM_GSEQ ? 16 : 0));
                const int j = select_marked_idx(tv.pm, ea.gplan + tv.in_base, gbit, last);
                if (j >= 0) {
                    const int pj = tv.pm[j] & kPmPosMask;
                    if (staged_valid(vw, pj)) { val = lds_fixed<W>(vals, pj); is_valid = true; }
                }
            } else if (GAGG && cd.mode == CM_GAGG) {
                fold_group_agg(cd, tv.pm, ea.gagg + tv.in_base, vw, vals, last, &val, &is_valid, ea.err);
                if (!is_valid) val = 0;
            } else if (cd.mode == CM_KEY) {
                val = lds_fixed<W>(vals, tv.pm[last] & kPmPosMask); is_valid = true;
            } else if (cd.mode == CM_SEQ) {
                val = (uint64_t)ea.tmp_seq[result_index(tv, ob)]; is_valid = true;
            } else if (cd.mode == CM_KIND) {
                val = (uint8_t)ea.tmp_kind[result_index(tv, ob)]; is_valid = true;
            } else {
                fold_fixed(cd, tv.pm, vw, vals, last, &val, &is_valid, ea.err);
                if (!is_valid) val = 0;
            }
        }
        if (active) stg_fixed<W>(oc.data, tv.out_base + ob, val);
        if (oc.validity != nullptr) put_validity_word(oc.validity, tv.out_base, wb, tv.n_out, is_valid);
    }
}

template <bool GAGG>
__device__ __forceinline__ void emit_fixed_dispatch(const EmitArgs &ea, const ColDesc &cd, const pg_out_column &oc,
                                                    const TileView &tv, const unsigned char *vals, const uint32_t *vw) {
    if (cd.width == 8) emit_fixed_column<8, GAGG>(ea, cd, oc, tv, vals, vw);
    else if (cd.width == 4) emit_fixed_column<4, GAGG>(ea, cd, oc, tv, vals, vw);
    else if (cd.width == 1) emit_fixed_column<1, GAGG>(ea, cd, oc, tv, vals, vw);
    else emit_fixed_column<2, GAGG>(ea, cd, oc, tv, vals, vw);
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

#ifdef PG_EMIT_TIMING
// per-pass cycle stamps of a few tiles in the middle of the grid (experiments; profiles/README.md):
// slot 0 = kernel entry, 1 = prologue done, then per pipelined pass u: 2+3u = top, 3+3u = stage full, 4+3u = pass done
constexpr int kTsTiles = 64, kTsSlots = 256;
__device__ long long g_emit_ts[2 * kTsTiles * kTsSlots];
#define TS(slot) do { if (ts_on && lane == 0 && (slot) < kTsSlots) g_emit_ts[(ts_w * kTsTiles + ts_tile) * kTsSlots + (slot)] = clock64(); } while (0)
#else
#define TS(slot) do { } while (0)
#endif

// GAGG: the merge has aggregate functions inside sequence groups (the fold for them is compiled into its own
// kernel variant: inlined into the common one it costs every workload registers and spills)
template <bool GAGG>
__global__ void __launch_bounds__(kEmitThreads, 2)
k_emit(EmitArgs ea) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int k = ea.k, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nv = ea.n_varlen;
    const EmitLayout L = emit_layout(k, nv);
    // (computed from `smem` at every use: pointers picked out of a local array lose their address space, and every
    // read of staged data becomes a generic LD through L1TEX instead of an LDS — ncu: long-scoreboard stalls)
    auto stage_vals = [&](int s) -> unsigned char * { return smem + (size_t)s * L.stage_bytes; };
    auto stage_vw = [&](int s) -> uint32_t * { return (uint32_t *)(smem + (size_t)s * L.stage_bytes + (size_t)L.rt * 8); };
    unsigned char *p = smem + kStages * L.stage_bytes;
    int64_t *rstart = (int64_t *)p;            p += PG_MAX_RUNS * 8;
    const uint8_t **cdata = (const uint8_t **)p; p += kStages * PG_MAX_RUNS * 8;   // payload base per run (var-len)
    uint64_t *mbar_full = (uint64_t *)p;       p += kStages * 8;
    uint64_t *mbar_empty = (uint64_t *)p;      p += kStages * 8;
    int64_t *s_base = (int64_t *)p;            p += (size_t)nv * 8;            // per var-len column: byte base of the tile
    p += 16;
    uint16_t *pm = (uint16_t *)p;              p += kTileMax * 2;
    uint16_t *glast = (uint16_t *)p;           p += kTileMax * 2;   // per output row: last member position
    uint8_t *mrun = (uint8_t *)p;              p += kTileMax;
    uint8_t *gcnt = (uint8_t *)p;              p += kTileMax;       // per output row: members of its key group
    uint32_t *aebits = (uint32_t *)p;          p += kTileMax / 8;   // per merged position: may provide a value
    int *seg = (int *)p;                       p += (PG_MAX_RUNS + 1) * 4;
    int *seg_a = (int *)p;                     p += (PG_MAX_RUNS + 1) * 4;        // first plan tile: slot base per run
    int *rr = (int *)p;                        p += (PG_MAX_RUNS + 1) * 4;         // staged row base per run
    int *ws = (int *)p;                        p += 34 * 4;
    int *s_i32 = (int *)p;                     p += 16;
    int *vwg = (int *)p;                       p += (size_t)(kTileMax + 64 * PG_MAX_RUNS) / 32 * 4;   // staged validity word -> global word
    int *wtot = (int *)p;                      p += (size_t)nv * kEmitWarps * 4;   // per var-len column, per warp: bytes
    int *s_tot = (int *)p;                     p += (size_t)nv * 4;
    int *s_cnt = (int *)p;                     p += (size_t)nv * 4;

    // ---- tile ticket (tiles are started in order => look-back never waits on an unscheduled tile)
    if (tid == 0) {
        s_i32[0] = atomicAdd(ea.tile_counter, 1);
        for (int s = 0; s < kStages; s++) { mbar_init(&mbar_full[s], 1 + kEmitWarps); mbar_init(&mbar_empty[s], kEmitWarps); }
        mbar_fence_init();
    }
    for (int i = tid; i < kTileMax / 32; i += kEmitThreads) aebits[i] = 0;
    for (int i = tid; i < nv; i += kEmitThreads) { s_tot[i] = 0; s_cnt[i] = 0; }
    __syncthreads();
    const int tile = s_i32[0];
#ifdef PG_EMIT_TIMING
    const int ts_tile = tile - ea.n_tiles / 2;
    const bool ts_on = ts_tile >= 0 && ts_tile < kTsTiles && (warp == 0 || warp == kEmitWarps - 1);
    const int ts_w = warp == 0 ? 0 : 1;
#endif
    TS(0);
    // an emit tile is two consecutive plan tiles (the last one may be single)
    const int plan_a = 2 * tile, plan_end = min(plan_a + 2, ea.n_plan_tiles);
    if (warp == 0) {
        // lane r = run r (its three bounds in flight together); slot / staged-row bases by warp scans
        int64_t b0 = 0, bm = 0, b1 = 0;
        if (lane < k) {
            b0 = ea.bounds[(int64_t)plan_a * k + lane];
            bm = ea.bounds[(int64_t)(plan_a + 1) * k + lane];
            b1 = ea.bounds[(int64_t)plan_end * k + lane];
        }
        const int len = (int)(b1 - b0), len_a = (int)(bm - b0);
        const int rlen = lane < k ? (((int)(b0 & 31) + len + 1) + 31) & ~31 : 0;
        const int i_len = warp_scan_incl(len), i_a = warp_scan_incl(len_a), i_r = warp_scan_incl(rlen);
        if (lane < k) {
            rstart[lane] = b0;
            seg[lane] = i_len - len;
            seg_a[lane] = i_a - len_a;
            rr[lane] = i_r - rlen;
        }
        if (lane == 31) { seg[k] = i_len; seg_a[k] = i_a; rr[k] = i_r; }
    }
    __syncthreads();
    const int n = seg[k];
    const int n_vw = rr[k] >> 5;                       // staged validity words
    int64_t in_base = 0;
    for (int r = 0; r < k; r++) in_base += rstart[r];

    // ---- plan -> (staged position | op | head), output rows
    constexpr int VT = kTileMax / kEmitThreads;
    const int p0 = tid * VT, p1 = min(p0 + VT, n);
    int my = 0;
    uint32_t emit_bits = 0;
    const bool single = ea.single_winner != 0;
    // slot -> (staged position, run) tables, built per run segment (no searches); they live in memory that
    // is not in use yet (glast, stage 1)
    uint16_t *spos = glast;
    uint8_t *srun = stage_vals(1);
    // plan slots are run-major inside their plan tile: table index = slot (first tile) or n_a + slot (second)
    const int n_a = seg_a[k];
    for (int r = 0; r < k; r++) {
        const int len_a = seg_a[r + 1] - seg_a[r], len = seg[r + 1] - seg[r];
        const int p0r = rr[r] + (int)(rstart[r] & 31);                  // staged position of the run's first row
        const int sb0 = n_a + (seg[r] - seg_a[r]);                      // second tile: slot base of run r
        for (int j = tid; j < len; j += kEmitThreads) {
            const int idx = j < len_a ? seg_a[r] + j : sb0 + (j - len_a);
            spos[idx] = (uint16_t)(p0r + j);
            srun[idx] = (uint8_t)r;
        }
    }
    // staged validity word j -> word of its run's bitmap (or -1: beyond the run's bitmap, never used)
    if (tid < n_vw) {
        int r = 0;
        while (r + 1 < k && (rr[r + 1] >> 5) <= tid) r++;
        const int64_t gw = (rstart[r] >> 5) + (tid - (rr[r] >> 5));
        vwg[tid] = gw > ((ea.run_rows[r] - 1) >> 5) ? -1 : (int)((gw << 5) | r);     // word index (26 bits) | run
    }
    __syncthreads();
    for (int i = p0; i < p1; i++) {
        uint16_t e = ea.plan[in_base + i];
        int slot = (e & kPlanSlotMask) + (i < n_a ? 0 : n_a);
        mrun[i] = srun[slot];
        uint32_t pos = spos[slot];
        uint32_t op = (e >> kPlanOpShift) & 3;
        pm[i] = (uint16_t)(pos | (op << kPmOpShift) | ((e & kPlanHead) ? kPmHead : 0));
        if ((e & kPlanHead) && (e & kPlanEmit)) { my++; emit_bits |= 1u << (i - p0); }
    }
    int n_out = 0;
    int o = block_scan_excl(my, ws, &n_out);          // (syncs: pm is complete afterwards)
    for (int i = p0; i < p1; i++) {
        if (emit_bits & (1u << (i - p0))) {
            int e = i + 1;
            while (e < n && !(pm[e] & kPmHead)) e++;
            int gl = e - 1;
            if (single) {
                // deduplicate / first-row: exactly one member of an emitted group has an op (its cells are the row)
                for (int j = e - 1; j >= i; j--)
                    if (((pm[j] >> kPmOpShift) & 3) != OP_NOOP) { gl = j; break; }
            }
            glast[o] = (uint16_t)gl;
            gcnt[o] = (uint8_t)(e - i);
            o++;
#ifdef PG_EMIT_BITMAP_SELECT
            // eligibility: newest first, every member with an op until (and including) the newest SET
            for (int j = e - 1; j >= i; j--) {
                uint32_t ej = pm[j];
                uint32_t op = (ej >> kPmOpShift) & 3;
                if (GAGG && op == OP_RETRACT) op = (ej & kPmHead) ? OP_SET : OP_NOOP;
                if (op != OP_NOOP) atomicOr(&aebits[j >> 5], 1u << (j & 31));
                if (op == OP_SET) break;
            }
#endif
        }
    }
    __syncthreads();

    TileView tv;
    tv.pm = pm;
    tv.glast = glast;
    tv.gcnt = gcnt;
    tv.aebits = aebits;
    tv.n = n;
    tv.single = single;
    tv.n_out = n_out;
    tv.out_base = ea.row_base[plan_a];
    tv.o_shift = (int)(tv.out_base & 31);
    tv.in_base = in_base;
    tv.n_out_a = ea.tile_rows[plan_a];
    tv.seq_shift = n_a - tv.n_out_a;

    // issue the bulk copies of column c into stage s (warp 0; lane r = run r)
    auto issue = [&](int c, int s) {
        const ColDesc cd = ea.cols[c];
        // the stage was last touched through the generic proxy (scratch + validity words)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        uint32_t bytes = 0;
        const unsigned char *src = nullptr;
        unsigned char *dst = nullptr;
        if (lane < k) {
            const int r = lane;
            const int len = seg[r + 1] - seg[r];
            const int64_t row0 = rstart[r] & ~(int64_t)31;
            const int head = (int)(rstart[r] & 31);
            if (cd.width > 0) {
                if (len > 0) {
                    bytes = (uint32_t)(((head + len) * cd.width + 15) & ~15);
                    src = (const unsigned char *)ea.ptrs.data[(int64_t)c * k + r] + row0 * cd.width;
                    dst = stage_vals(s) + (size_t)rr[r] * cd.width;
                }
            } else {
                if (len > 0) {
                    bytes = (uint32_t)(((head + len + 1) * 4 + 15) & ~15);
                    src = (const unsigned char *)(ea.ptrs.offsets[(int64_t)c * k + r] + row0);
                    dst = stage_vals(s) + (size_t)rr[r] * 4;
                }
                cdata[s * PG_MAX_RUNS + r] = (const uint8_t *)ea.ptrs.data[(int64_t)c * k + r];
            }
        }
        uint32_t total = bytes;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
        if (lane == 0) mbar_arrive_expect_tx(&mbar_full[s], total);
        __syncwarp();
        if (bytes) bulk_g2s(dst, src, bytes, &mbar_full[s]);
    };

    const int n_pass = ea.n_passes;
    // passes without staged data (sequence number, kind) come first in the list and do not take part in the pipeline
    int pp = 0;
    while (pp < n_pass && (ea.col_order[pp] >> 16) == PH_PLAIN) pp++;
    const int pbase = pp;                              // first pipelined pass
    // validity words of a column: every thread brings (at most) one staged word, global -> register -> shared.  The
    // load for pass pp + 1 is issued at the top of pass pp and stored at its end, so its latency hides behind the pass.
    const int my_vwg = tid < n_vw ? vwg[tid] : -1;
    auto load_vw = [&](int c) -> uint32_t {
        uint32_t x = 0xffffffffu;
        if (my_vwg >= 0) {
            const uint32_t *vp = ea.ptrs.validity[(int64_t)c * k + (my_vwg & 31)];
            if (vp) x = __ldg(vp + (my_vwg >> 5));
        }
        return x;
    };
    const bool vw_warp = warp == 0 || warp * 32 < n_vw;   // this warp stores validity words (warp 0: always, it issues)
    if (pbase < n_pass) {
        // the first staged column is in flight while the unstaged passes run
        const int c0 = ea.col_order[pbase] & 0xffff;
        if (warp == 0) issue(c0, 0);
        const uint32_t x = load_vw(c0);
        if (tid < n_vw) stage_vw(0)[tid] = x;
        __syncwarp();
        if (lane == 0) mbar_arrive(&mbar_full[0]);
    }
    for (int q = 0; q < pbase; q++) {
        const int c = ea.col_order[q] & 0xffff;
        emit_fixed_dispatch<GAGG>(ea, ea.cols[c], ea.out_cols[c], tv, nullptr, nullptr);
    }

    // rows of a var-len column are dealt to warps in contiguous chunks of RW rows (RW a multiple of 32, aligned to
    // the output's 32-row validity words): offsets come from warp-level scans
    const int span = tv.o_shift + n_out;
    const int RW = (((span + kEmitWarps - 1) / kEmitWarps) + 31) & ~31;
    const int wbeg = -tv.o_shift + warp * RW;

    TS(1);
    for (pp = pbase; pp < n_pass; pp++) {
        const int u = pp - pbase;                      // pipelined pass number: stage u & 1, use u >> 1 of that stage
        TS(2 + 3 * u);
        const int s = u & 1;
        const int ent = ea.col_order[pp];
        const int c = ent & 0xffff, phase = ent >> 16;
        const ColDesc cd = ea.cols[c];
        const pg_out_column oc = ea.out_cols[c];
        const bool more = pp + 1 < n_pass;
        uint32_t next_vw = 0;
        if (more) {
            const int cn = ea.col_order[pp + 1] & 0xffff;
            next_vw = load_vw(cn);
            if (warp == 0) {
                // refill the other stage for the next pass, once every warp has released its previous use
                if (u >= 1) mbar_wait(&mbar_empty[s ^ 1], (uint32_t)(((u - 1) >> 1) & 1));
                issue(cn, s ^ 1);
            } else if (!vw_warp && lane == 0) {
                mbar_arrive(&mbar_full[s ^ 1]);        // nothing to store: arrive right away
            }
        }
        if (phase == PH_COPY && (ea.col_order[pp - 1] >> 16) != PH_COPY) {
            // ---- every var-len column's byte base: decoupled look-back over earlier tiles, one warp per column.
            // The aggregates were published by the size passes ~40 passes ago, so no predecessor is waited for.
            __syncthreads();
            for (int vi = warp; vi < nv; vi += kEmitWarps) {
                const int tile_bytes = s_tot[vi];
                uint64_t *state = ea.vl_state + (int64_t)vi * ea.n_tiles;
                uint64_t excl = 0;
                // every lane looks at kLook consecutive predecessors per step (all loads of a step are independent)
                constexpr int kLook = 8;
                int t = tile - 1;
                while (t >= 0) {
                    uint64_t sv[kLook];
#pragma unroll
                    for (int q = 0; q < kLook; q++) {
                        const int idx = t - (lane * kLook + q);
                        sv[q] = idx >= 0 ? *((volatile uint64_t *)&state[idx]) : kFlagPrefix;
                    }
                    // position p = lane * kLook + q (0 = nearest predecessor): first prefix, nothing unpublished in front
                    int lp = 32 * kLook;
#pragma unroll
                    for (int q = kLook - 1; q >= 0; q--) if ((unsigned)(sv[q] >> 62) == 2) lp = lane * kLook + q;
                    int first = lp;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, d));
                    bool inval = false;
                    uint64_t contrib = 0;
#pragma unroll
                    for (int q = 0; q < kLook; q++) {
                        if (lane * kLook + q <= first) {
                            if ((unsigned)(sv[q] >> 62) == 0) inval = true;
                            contrib += sv[q] & kValMask;
                        }
                    }
                    if (__any_sync(0xffffffffu, inval)) continue;      // a needed predecessor has not published yet
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
                    excl += contrib;
                    if (first < 32 * kLook) break;
                    t -= 32 * kLook;
                }
                if (lane == 0) {
                    __threadfence();
                    atomicExch((unsigned long long *)&state[tile], kFlagPrefix | (excl + (uint64_t)tile_bytes));
                    s_base[vi] = (int64_t)excl;
                    if (tile == ea.n_tiles - 1) {
                        const int vc = ea.varlen_cols[vi];
                        uint64_t tot = excl + (uint64_t)tile_bytes;
                        ea.totals[1 + vi] = (int64_t)tot;
                        if (tot > 0x7fffffffull) atomicCAS(ea.err, KERR_NONE, KERR_OFFSET_OVERFLOW);
                        ea.out_cols[vc].offsets[ea.totals[0]] = (int32_t)tot;
                    }
                }
            }
            __syncthreads();
        }
        mbar_wait(&mbar_full[s], (uint32_t)((u >> 1) & 1));
        TS(3 + 3 * u);
        const unsigned char *vals = stage_vals(s);
        const uint32_t *vw = stage_vw(s);

        if (phase == PH_FIXED) {
            emit_fixed_dispatch<GAGG>(ea, cd, oc, tv, vals, vw);
        } else if (phase == PH_SIZE) {
            // ---- var-len column, size pass: source member per output row (kept in global scratch for the copy
            // pass), byte total of this warp's rows; the last warp to finish publishes the tile's aggregate
            const int vi = cd.varlen_index;
            const int32_t *offs = (const int32_t *)vals;
            const uint8_t *const *cd_data = cdata + s * PG_MAX_RUNS;
            uint16_t *vsrc = ea.vsrc + (int64_t)vi * ea.vsrc_stride + tv.out_base;
            int my_bytes = 0;
            for (int ob0 = wbeg; ob0 < wbeg + RW && ob0 < n_out; ob0 += 32) {
                const int ob = ob0 + lane;
                const bool active = ob >= 0 && ob < n_out;
                const int last = active ? glast[ob] : 0;
                int src = -1;
                if (cd.mode == CM_SELECT) src = select_idx_warp<GAGG>(tv, vw, active, last, active ? gcnt[ob] : 1);
                else if (active) {
                    if (cd.mode == CM_KEY) src = last;
                    else if (cd.mode == CM_FOLD) src = fold_member_idx(cd, pm, mrun, vw, offs, cd_data, last, ea.err);
                    else {                                              // CM_GVAL / CM_GSEQ
                        src = select_marked_idx(pm, ea.gplan + in_base, 1u << (cd.agg + (cd.mode == CM_GSEQ ? 16 : 0)), last);
                        if (src >= 0 && !staged_valid(vw, pm[src] & kPmPosMask)) src = -1;
                    }
                }
                if (active) {
                    vsrc[ob] = src < 0 ? (uint16_t)0xFFFF : (uint16_t)src;
                    if (src >= 0) { const int ps = pm[src] & kPmPosMask; my_bytes += offs[ps + 1] - offs[ps]; }
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) my_bytes += __shfl_xor_sync(0xffffffffu, my_bytes, d);
            if (lane == 0) {
                wtot[vi * kEmitWarps + warp] = my_bytes;
                atomicAdd(&s_tot[vi], my_bytes);
                __threadfence_block();
                if (atomicAdd(&s_cnt[vi], 1) == kEmitWarps - 1) {
                    const uint64_t tb = (uint64_t)(uint32_t)atomicAdd(&s_tot[vi], 0);
                    __threadfence();
                    // (tile 0 knows its prefix; everyone else publishes the aggregate now, the prefix in the look-back)
                    atomicExch((unsigned long long *)&ea.vl_state[(int64_t)vi * ea.n_tiles + tile],
                               (tile == 0 ? kFlagPrefix : kFlagAgg) | tb);
                }
            }
        } else {
            // ---- var-len column, copy pass: offsets, validity, payload — warp-local.  The upper half of the stage is
            // free (offsets are 4 bytes per row) and holds per-warp scratch.
            const int vi = cd.varlen_index;
            const int32_t *offs = (const int32_t *)vals;
            int *wpre = (int *)(vals + (size_t)L.rt * 4);                            // per warp: 33 ints
            const uint8_t **wsrc = (const uint8_t **)(wpre + ((kEmitWarps * 33 + 1) & ~1));   // 8-byte aligned
            int *wend = (int *)(wsrc + kEmitWarps * 32);                             // per warp: 32 ints
            const uint8_t *const *cd_data = cdata + s * PG_MAX_RUNS;
            const uint16_t *vsrc = ea.vsrc + (int64_t)vi * ea.vsrc_stride + tv.out_base;
            const int64_t byte_base = s_base[vi];
            uint8_t *dbase = as_global((uint8_t *)oc.data) + byte_base;
            int32_t *out_offs = as_global(oc.offsets);
            int carry = 0;
            for (int w2 = 0; w2 < warp; w2++) carry += wtot[vi * kEmitWarps + w2];
            int *my_pre = wpre + warp * 33;
            int *my_end = wend + warp * 32;
            const uint8_t **my_src = wsrc + warp * 32;
            for (int ob0 = wbeg; ob0 < wbeg + RW && ob0 < n_out; ob0 += 32) {
                const int ob = ob0 + lane;
                const bool active = ob >= 0 && ob < n_out;
                int len = 0;
                const uint8_t *sp = nullptr;
                bool has = false;
                if (active) {
                    int src = vsrc[ob];
                    if (src != 0xFFFF) {
                        int ps = pm[src] & kPmPosMask;
                        int st = offs[ps];
                        len = offs[ps + 1] - st;
                        sp = as_global(cd_data[mrun[src]]) + st;
                        has = true;
                    }
                }
                const int incl = warp_scan_incl(len);
                const int off = carry + incl - len;
                carry += __shfl_sync(0xffffffffu, incl, 31);
                if (active) out_offs[tv.out_base + ob] = (int32_t)(byte_base + off);
                if (oc.validity != nullptr) put_validity_word(oc.validity, tv.out_base, ob0, n_out, has);
                // warp-cooperative payload copy: the warp's 32 rows form one contiguous destination range;
                // rows without payload (NULL / empty) are squeezed out first, then 8 lanes serve one row (so
                // stores coalesce and short strings do not idle a whole warp), two row groups are in flight at
                // a time and all loads are issued before the stores
                const unsigned pay = __ballot_sync(0xffffffffu, len > 0);
                const int n_pay = __popc(pay);
                if (len > 0) {
                    const int ci = __popc(pay & ((1u << lane) - 1));
                    my_pre[ci] = off;
                    my_end[ci] = off + len;
                    my_src[ci] = sp;
                }
                __syncwarp();
                for (int rg = 0; rg < n_pay; rg += 8) {
                    const int ra = rg + (lane >> 3), rb = ra + 4, l8 = lane & 7;
                    int a0 = 0, la = 0, b0 = 0, lb = 0;                   // destination offset, bytes
                    const uint8_t *pa = nullptr, *pb = nullptr;
                    if (ra < n_pay) { a0 = my_pre[ra]; la = my_end[ra] - a0; pa = my_src[ra]; }
                    if (rb < n_pay) { b0 = my_pre[rb]; lb = my_end[rb] - b0; pb = my_src[rb]; }
                    uint8_t *da = dbase + a0, *db = dbase + b0;
                    uint8_t xa0 = 0, xa1 = 0, xa2 = 0, xb0 = 0, xb1 = 0, xb2 = 0;
                    if (l8 < la) xa0 = __ldg(&pa[l8]);
                    if (l8 + 8 < la) xa1 = __ldg(&pa[l8 + 8]);
                    if (l8 + 16 < la) xa2 = __ldg(&pa[l8 + 16]);
                    if (l8 < lb) xb0 = __ldg(&pb[l8]);
                    if (l8 + 8 < lb) xb1 = __ldg(&pb[l8 + 8]);
                    if (l8 + 16 < lb) xb2 = __ldg(&pb[l8 + 16]);
                    if (l8 < la) da[l8] = xa0;
                    if (l8 + 8 < la) da[l8 + 8] = xa1;
                    if (l8 + 16 < la) da[l8 + 16] = xa2;
                    if (l8 < lb) db[l8] = xb0;
                    if (l8 + 8 < lb) db[l8 + 8] = xb1;
                    if (l8 + 16 < lb) db[l8 + 16] = xb2;
                    for (int b = l8 + 24; b < la; b += 8) da[b] = __ldg(&pa[b]);
                    for (int b = l8 + 24; b < lb; b += 8) db[b] = __ldg(&pb[b]);
                }
                __syncwarp();
            }
        }
        // release the stage: this warp is done with its rows of the pass
        TS(4 + 3 * u);
        __syncwarp();
        if (lane == 0) mbar_arrive(&mbar_empty[s]);
        if (more && vw_warp) {
            // the next pass's validity words go into the other stage once its previous use (pass pp - 1) is released
            if (u >= 1 && warp != 0) mbar_wait(&mbar_empty[s ^ 1], (uint32_t)(((u - 1) >> 1) & 1));
            if (tid < n_vw) stage_vw(s ^ 1)[tid] = next_vw;
            __syncwarp();
            if (lane == 0) mbar_arrive(&mbar_full[s ^ 1]);
        }
    }
}

#ifdef PG_EMIT_TIMING
static void emit_timing_dump(const EmitArgs &ea) {
    static long long h[2 * kTsTiles * kTsSlots];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h, g_emit_ts, sizeof(h));
    std::vector<int32_t> order(ea.n_passes);
    cudaMemcpy(order.data(), ea.col_order, 4 * (size_t)ea.n_passes, cudaMemcpyDeviceToHost);
    int pbase = 0;
    while (pbase < ea.n_passes && (order[pbase] >> 16) == PH_PLAIN) pbase++;
    for (int w = 0; w < 2; w++) {
        // per phase kind: wait for the stage, compute; plus the prologue and the whole tile
        double wait[4] = {0}, comp[4] = {0}, gap[4] = {0}, pro = 0, whole = 0;
        int cnt[4] = {0}, tiles = 0;
        for (int t = 0; t < kTsTiles; t++) {
            const long long *x = h + (size_t)(w * kTsTiles + t) * kTsSlots;
            if (!x[0] || !x[1]) continue;
            tiles++;
            pro += (double)(x[1] - x[0]);
            long long last = x[1];
            for (int u = 0; pbase + u < ea.n_passes && 4 + 3 * u < kTsSlots; u++) {
                const int ph = order[pbase + u] >> 16;
                const long long a = x[2 + 3 * u], b = x[3 + 3 * u], c = x[4 + 3 * u];
                if (!a || !b || !c) break;
                wait[ph] += (double)(b - a); comp[ph] += (double)(c - b); gap[ph] += (double)(a - last); cnt[ph]++;
                last = c;
            }
            whole += (double)(last - x[0]);
        }
        if (!tiles) continue;
        fprintf(stderr, "[emit timing] warp %s, %d tiles: prologue %.0f, tile %.0f cycles;", w ? "15" : "0", tiles, pro / tiles, whole / tiles);
        const char *nm[4] = {"plain", "size", "fixed", "copy"};
        for (int ph = 1; ph < 4; ph++)
            if (cnt[ph]) fprintf(stderr, " %s x%.1f: wait %.0f compute %.0f tail %.0f;", nm[ph], (double)cnt[ph] / tiles,
                                 wait[ph] / cnt[ph], comp[ph] / cnt[ph], gap[ph] / cnt[ph]);
        fprintf(stderr, "\n");
    }
    static long long z[2 * kTsTiles * kTsSlots];
    cudaMemcpyToSymbol(g_emit_ts, z, sizeof(z));
}
#endif

static bool g_emit_attr = false;
void launch_emit(const EmitArgs &ea) {
    EmitLayout L = emit_layout(ea.k, ea.n_varlen);
    if (!g_emit_attr) {
        cudaFuncSetAttribute(k_emit<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(k_emit<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        g_emit_attr = true;
    }
    if (ea.gagg) k_emit<true><<<ea.n_tiles, kEmitThreads, L.total, ea.stream>>>(ea);
    else k_emit<false><<<ea.n_tiles, kEmitThreads, L.total, ea.stream>>>(ea);
#ifdef PG_EMIT_TIMING
    emit_timing_dump(ea);
#endif
}

}  // namespace pg
