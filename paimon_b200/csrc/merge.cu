// merge.cu — sm_100a kernels of the k-way merge + per-key-group reduce.
//
// Replaces the per-record loop of SortMergeReaderWithLoserTree.SortMergeIterator.next()
// (reference: paimon-core/.../mergetree/compact/SortMergeReaderWithLoserTree.java:87-112,
// LoserTree.java:95-170) with a data-parallel pipeline over columnar runs resident in HBM:
//
//   1. sampled partition   every S-th key of each run forms the next level; levels are merged
//                          top-down so that level 0 is cut into key-range tiles of <= kTileMax rows
//                          (all rows of one key fall into one tile => the reduce is tile-local)
//   2. plan kernel         one CTA per tile: k sorted segments -> shared memory -> log2(k) rounds of
//                          merge-path pair merges -> key groups -> members ordered by sequence number
//                          -> per-member op (the MergeFunction's row-level semantics) -> uint16 plan
//   3. scan                tile row counts / var-len byte counts -> output offsets
//   4. emit kernel         per tile and column: select / fold the group's members, coalesced stores
//
// No tensor cores: the path is integer compare + gather (HBM bound).
#include "pg_internal.h"

namespace pg {

// ------------------------------------------------------------------ helpers

__device__ __forceinline__ uint64_t norm_field(const void *base, int type, int64_t row) {
    // order-preserving unsigned image of a signed integer key field (GenerateUtils.scala:122-123:
    // integers compare with signed < / >)
    switch (type) {
        case PG_INT8:  return (uint64_t)(uint8_t)(((const int8_t *)base)[row] ^ (int8_t)0x80);
        case PG_BOOL:  return ((const uint8_t *)base)[row] ? 1u : 0u;
        case PG_INT16: return (uint64_t)(uint16_t)(((const int16_t *)base)[row] ^ (int16_t)0x8000);
        case PG_INT32: return (uint64_t)(uint32_t)(((const int32_t *)base)[row] ^ (int32_t)0x80000000);
        default:       return (uint64_t)(((const int64_t *)base)[row]) ^ 0x8000000000000000ull;
    }
}

__device__ __forceinline__ uint64_t load_key(const void *const *key_ptrs, const KeyDesc &kd, int run,
                                             int64_t row) {
    if (kd.n_fields == 1) return norm_field(key_ptrs[run], kd.type[0], row);
    uint64_t k = 0;
    for (int f = 0; f < kd.n_fields; f++)
        k |= norm_field(key_ptrs[run * kd.n_fields + f], kd.type[f], row) << kd.shift[f];
    return k;
}

__device__ __forceinline__ int warp_scan_incl(int v) {
    int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}

// exclusive scan over the block; *total receives the block sum.  `ws` = 33 ints of shared memory.
__device__ __forceinline__ int block_scan_excl(int v, int *ws, int *total) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = warp_scan_incl(v);
    if (lane == 31) ws[w] = incl;
    __syncthreads();
    if (w == 0) {
        int x = lane < nw ? ws[lane] : 0;
        int xi = warp_scan_incl(x);
        ws[lane] = xi - x;
        if (lane == 31) ws[32] = xi;
    }
    __syncthreads();
    int res = ws[w] + incl - v;
    *total = ws[32];
    __syncthreads();
    return res;
}

__device__ __forceinline__ bool valid_bit(const uint8_t *bm, int64_t row) {
    return bm == nullptr || ((bm[row >> 3] >> (row & 7)) & 1);
}

__device__ __forceinline__ bool kind_is_retract(int kind) {   // RowKind.java:101-103
    return kind == PG_UPDATE_BEFORE || kind == PG_DELETE;
}

// ------------------------------------------------------------------ partition

__global__ void k_partition(int k, KeyDesc kd, const void *const *key_ptrs, LevelView lv,
                            const uint64_t *__restrict__ sk, int q, int n_tiles, int64_t *bounds) {
    int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= (int64_t)(n_tiles + 1) * k) return;
    int t = (int)(idx / k), r = (int)(idx % k);
    int64_t res;
    if (t == 0) res = 0;
    else if (t == n_tiles) res = lv.count[r];
    else {
        uint64_t x = sk[(int64_t)t * q];
        int64_t lo = 0, hi = lv.count[r];
        while (lo < hi) {                      // lower_bound: first j with key(j) >= x
            int64_t mid = (lo + hi) >> 1;
            uint64_t km = load_key(key_ptrs, kd, r, (mid + 1) * lv.stride - 1);
            if (km < x) lo = mid + 1; else hi = mid;
        }
        res = lo;
    }
    bounds[idx] = res;
}

// ------------------------------------------------------------------ in-tile merge

struct TileCtx {
    uint64_t *key[2];
    uint16_t *idx[2];
    int *lb[2];          // list boundaries, k+1 entries each
    int *seg;            // slot base per run, k+1 entries
    int64_t *rstart;     // first row (at this level) of every run's segment
    int n;               // rows in the tile
    int fin;             // which buffer holds the merged result
};

__device__ __forceinline__ int run_of_slot(const int *seg, int k, int slot) {
    int r = 0;
    while (r + 1 < k && seg[r + 1] <= slot) r++;
    return r;
}

// Loads the tile's k segments and merges them.  Returns false when the tile overflows.
__device__ bool merge_tile(TileCtx &tc, int k, const KeyDesc &kd, const void *const *key_ptrs,
                           const int64_t *bounds, int tile, int64_t stride, int32_t *err) {
    const int tid = threadIdx.x;
    if (tid == 0) {
        int acc = 0;
        for (int r = 0; r < k; r++) {
            int64_t b0 = bounds[(int64_t)tile * k + r], b1 = bounds[(int64_t)(tile + 1) * k + r];
            tc.rstart[r] = b0;
            tc.seg[r] = acc;
            tc.lb[0][r] = acc;
            int64_t len = b1 - b0;
            if (len < 0 || acc + len > kTileMax) { len = 0; acc = kTileMax + 1; }
            else acc += (int)len;
        }
        tc.seg[k] = acc;
        tc.lb[0][k] = acc;
    }
    __syncthreads();
    tc.n = tc.seg[k];
    if (tc.n > kTileMax) {
        if (tid == 0) atomicCAS(err, KERR_NONE, KERR_TILE_OVERFLOW);
        return false;
    }
    const int n = tc.n;
    for (int i = tid; i < n; i += blockDim.x) {
        int r = run_of_slot(tc.seg, k, i);
        int64_t j = tc.rstart[r] + (i - tc.seg[r]);
        tc.key[0][i] = load_key(key_ptrs, kd, r, (j + 1) * stride - 1);
        tc.idx[0][i] = (uint16_t)i;
    }
    __syncthreads();

    int L = k, cur = 0;
    constexpr int VT = kTileMax / kThreads;
    while (L > 1) {
        const uint64_t *sk = tc.key[cur];
        const uint16_t *si = tc.idx[cur];
        uint64_t *dk = tc.key[cur ^ 1];
        uint16_t *di = tc.idx[cur ^ 1];
        const int *lb = tc.lb[cur];
        int pos = tid * VT, end = min(pos + VT, n);
        int p = 0;
        while (pos < end) {
            // pair p merges lists 2p and 2p+1
            while (lb[min(2 * p + 2, L)] <= pos) p++;
            int a0 = lb[2 * p], a1 = lb[min(2 * p + 1, L)], b1 = lb[min(2 * p + 2, L)];
            int na = a1 - a0, nb = b1 - a1;
            int d = pos - a0;
            int cnt = min(end, b1) - pos;
            int lo = max(0, d - nb), hi = min(d, na);
            while (lo < hi) {                    // merge path: #elements taken from A among the first d
                int mid = (lo + hi) >> 1;
                if (sk[a0 + mid] <= sk[a1 + d - 1 - mid]) lo = mid + 1; else hi = mid;
            }
            int ai = lo, bi = d - lo;
            uint64_t ka = ai < na ? sk[a0 + ai] : 0, kb = bi < nb ? sk[a1 + bi] : 0;
            for (int s = 0; s < cnt; s++) {
                bool take_a = (bi >= nb) || (ai < na && ka <= kb);   // stable: lower run first on ties
                if (take_a) {
                    dk[pos + s] = ka; di[pos + s] = si[a0 + ai];
                    ai++; ka = ai < na ? sk[a0 + ai] : 0;
                } else {
                    dk[pos + s] = kb; di[pos + s] = si[a1 + bi];
                    bi++; kb = bi < nb ? sk[a1 + bi] : 0;
                }
            }
            pos += cnt;
        }
        int P = (L + 1) >> 1;
        if (tid <= P) tc.lb[cur ^ 1][tid] = lb[min(2 * tid, L)];
        __syncthreads();
        L = P;
        cur ^= 1;
    }
    tc.fin = cur;
    return true;
}

__device__ __forceinline__ void carve_tile(TileCtx &tc, unsigned char *smem, int k) {
    tc.key[0] = (uint64_t *)smem;
    tc.key[1] = tc.key[0] + kTileMax;
    tc.idx[0] = (uint16_t *)(tc.key[1] + kTileMax);
    tc.idx[1] = tc.idx[0] + kTileMax;
    tc.rstart = (int64_t *)(tc.idx[1] + kTileMax);
    tc.lb[0] = (int *)(tc.rstart + PG_MAX_RUNS);
    tc.lb[1] = tc.lb[0] + PG_MAX_RUNS + 1;
    tc.seg = tc.lb[1] + PG_MAX_RUNS + 1;
}
constexpr size_t kTileSmem = (size_t)kTileMax * (8 + 8 + 2 + 2) + PG_MAX_RUNS * 8 + 3 * (PG_MAX_RUNS + 1) * 4;

__global__ void __launch_bounds__(kThreads)
k_merge_keys(int k, KeyDesc kd, const void *const *key_ptrs, LevelView lv, const int64_t *bounds,
             uint64_t *sorted_keys, int32_t *err) {
    extern __shared__ __align__(16) unsigned char smem[];
    TileCtx tc;
    carve_tile(tc, smem, k);
    int tile = blockIdx.x;
    if (!merge_tile(tc, k, kd, key_ptrs, bounds, tile, lv.stride, err)) return;
    int64_t base = 0;
    for (int r = 0; r < k; r++) base += tc.rstart[r];
    const uint64_t *fk = tc.key[tc.fin];
    for (int i = threadIdx.x; i < tc.n; i += blockDim.x) sorted_keys[base + i] = fk[i];
}

// ------------------------------------------------------------------ var-len source selection

// Which member supplies a var-len cell of the result?  Returns the member position (index into the
// tile's merged order) or -1 for NULL.  `members` = [g0, g0+g) positions; plan entries give slot + op.
struct MemberRef { int run; int64_t row; };

__device__ __forceinline__ int bytes_compare(const uint8_t *a, int la, const uint8_t *b, int lb) {
    int n = min(la, lb);
    for (int i = 0; i < n; i++) {
        int d = (int)a[i] - (int)b[i];
        if (d) return d;
    }
    return la - lb;
}

template <typename GetRef>
__device__ int select_varlen_member(const ColDesc &cd, const DevColumn *run_cols, int n_cols, int col,
                                    const uint16_t *plan, int g0, int g, GetRef get_ref, int32_t *err) {
    if (cd.mode != CM_FOLD) {
        // ops-based select: newest UPD with a non-null cell, or the newest SET (null allowed)
        for (int j = g - 1; j >= 0; j--) {
            int op = (plan[g0 + j] >> kPlanOpShift) & 3;
            if (op == OP_NOOP) continue;
            MemberRef m = get_ref(g0 + j);
            bool v = valid_bit(run_cols[m.run * n_cols + col].validity, m.row);
            if (op == OP_SET) return v ? g0 + j : -1;
            if (v) return g0 + j;
        }
        return -1;
    }
    // aggregate engine on a var-len column: fold with the member index as accumulator
    int acc = -1;
    bool initialized = false;
    for (int j = 0; j < g; j++) {
        int op = (plan[g0 + j] >> kPlanOpShift) & 3;
        if (op == OP_NOOP) continue;
        MemberRef m = get_ref(g0 + j);
        const DevColumn &dc = run_cols[m.run * n_cols + col];
        bool v = valid_bit(dc.validity, m.row);
        int in = v ? g0 + j : -1;
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
                    MemberRef a = get_ref(acc);
                    const DevColumn &ac = run_cols[a.run * n_cols + col];
                    int oa = ac.offsets[a.row], la = ac.offsets[a.row + 1] - oa;
                    int ob = dc.offsets[m.row], lbn = dc.offsets[m.row + 1] - ob;
                    int d = bytes_compare((const uint8_t *)ac.data + oa, la, (const uint8_t *)dc.data + ob, lbn);
                    if (cd.agg == PG_AGG_MAX) { if (d < 0) acc = in; }
                    else { if (!(d < 0)) acc = in; }
                }
                break;
            default: break;
        }
    }
    return acc;
}

// ------------------------------------------------------------------ plan kernel

struct PlanSmemExtra {
    uint16_t *res_slot;   // per head position: slot whose sequence number is the result's (0xFFFF = 0)
    uint8_t *res_kind;    // per head position: result RowKind
    int *ws;              // 33 ints scan scratch
};
constexpr size_t kPlanSmem = kTileSmem + (size_t)kTileMax * 3 + 34 * 4 + 16;

__global__ void __launch_bounds__(kThreads, 2)
k_plan(int k, KeyDesc kd, const void *const *key_ptrs, PlanArgs pa, int32_t *err) {
    extern __shared__ __align__(16) unsigned char smem[];
    TileCtx tc;
    carve_tile(tc, smem, k);
    PlanSmemExtra px;
    px.res_slot = (uint16_t *)(smem + ((kTileSmem + 15) & ~(size_t)15));
    px.res_kind = (uint8_t *)(px.res_slot + kTileMax);
    px.ws = (int *)(px.res_kind + kTileMax);

    const int tile = blockIdx.x, tid = threadIdx.x;
    if (!merge_tile(tc, k, kd, key_ptrs, pa.bounds, tile, 1, err)) {
        if (tid == 0) {
            pa.tile_rows[tile] = 0;
            for (int v = 0; v < pa.n_varlen; v++) pa.tile_bytes[(int64_t)v * pa.n_tiles + tile] = 0;
        }
        return;
    }
    const int n = tc.n;
    const uint64_t *fk = tc.key[tc.fin];
    uint16_t *fi = tc.idx[tc.fin];
    int64_t *seq_s = (int64_t *)tc.key[tc.fin ^ 1];          // staged by slot
    uint8_t *kind_s = (uint8_t *)tc.idx[tc.fin ^ 1];          // first 4 KiB: kinds by slot
    uint8_t *ops = kind_s + kTileMax;                         // second 4 KiB: op per merged position

    int64_t in_base = 0;
    for (int r = 0; r < k; r++) in_base += tc.rstart[r];

    // stage sequence numbers and kinds (coalesced per run segment)
    for (int s = tid; s < n; s += blockDim.x) {
        int r = run_of_slot(tc.seg, k, s);
        int64_t row = tc.rstart[r] + (s - tc.seg[r]);
        seq_s[s] = pa.seq_ptrs[r][row];
        kind_s[s] = (uint8_t)pa.kind_ptrs[r][row];
    }
    __syncthreads();

    constexpr int VT = kTileMax / kThreads;
    const int p0 = tid * VT, p1 = min(p0 + VT, n);
    const MergeFlags fl = pa.flags;
    int my_emit = 0;

    for (int i = p0; i < p1; i++) {
        bool head = (i == 0) || (fk[i] != fk[i - 1]);
        if (!head) continue;
        int e = i + 1;
        while (e < n && fk[e] == fk[i]) e++;
        const int g = e - i;
        // members in ascending sequence order (SortMergeReaderWithLoserTree.java:52-65); ties (which the
        // reference leaves unspecified) resolve by run order = slot order
        for (int a = i + 1; a < e; a++) {
            uint16_t sa = fi[a];
            int64_t qa = seq_s[sa];
            int b = a - 1;
            while (b >= i) {
                uint16_t sb = fi[b];
                int64_t qb = seq_s[sb];
                if (qb < qa || (qb == qa && sb < sa)) break;
                fi[b + 1] = sb;
                b--;
            }
            fi[b + 1] = sa;
        }
        // row-level semantics of the merge function
        bool emit = true;
        int res_kind = PG_INSERT;
        uint16_t res_slot = 0xFFFF;
        if (g == 1) {
            // ReducerMergeFunctionWrapper.java:53-73: a lone record is returned untouched
            ops[i] = OP_SET;
            res_slot = fi[i];
            res_kind = kind_s[fi[i]];
        } else if (fl.engine == PG_ENGINE_DEDUPLICATE) {
            // DeduplicateMergeFunction.java:47-60
            int win = -1;
            for (int j = e - 1; j >= i; j--) {
                if (fl.ignore_delete && kind_is_retract(kind_s[fi[j]])) continue;
                win = j;
                break;
            }
            for (int j = i; j < e; j++) ops[j] = (j == win) ? OP_SET : OP_NOOP;
            if (win < 0) emit = false;
            else { res_slot = fi[win]; res_kind = kind_s[fi[win]]; }
        } else if (fl.engine == PG_ENGINE_FIRST_ROW) {
            // FirstRowMergeFunction.java:50-73
            int win = -1;
            for (int j = i; j < e; j++) {
                ops[j] = OP_NOOP;
                if (kind_is_retract(kind_s[fi[j]])) {
                    if (!fl.ignore_delete) atomicCAS(err, KERR_NONE, KERR_FIRST_ROW_RETRACT);
                    continue;
                }
                if (win < 0) win = j;
            }
            if (win < 0) emit = false;
            else { ops[win] = OP_SET; res_slot = fi[win]; res_kind = kind_s[fi[win]]; }
        } else if (fl.engine == PG_ENGINE_PARTIAL_UPDATE) {
            // PartialUpdateMergeFunction.java:121-175 (no sequence groups), getResult :354-362
            bool filled = false, meet = false, cur_del = false;
            for (int j = i; j < e; j++) {
                int kind = kind_s[fi[j]];
                int op = OP_NOOP;
                cur_del = false;
                if (kind_is_retract(kind)) {
                    if (!filled) { op = OP_SET; filled = true; }          // initRow
                    if (!fl.ignore_delete) {
                        res_slot = fi[j];                                  // latestSequenceNumber
                        if (fl.remove_record_on_delete) {
                            if (kind == PG_DELETE) { cur_del = true; op = OP_SET; }
                        } else {
                            atomicCAS(err, KERR_NONE, KERR_PU_DELETE);
                        }
                    }
                } else {
                    res_slot = fi[j];
                    op = OP_UPD;
                    meet = true;
                    filled = true;
                }
                ops[j] = (uint8_t)op;
            }
            res_kind = (cur_del || !meet) ? PG_DELETE : PG_INSERT;
        } else {
            // AggregateMergeFunction.java:80-125
            bool cur_del = false;
            for (int j = i; j < e; j++) {
                int kind = kind_s[fi[j]];
                cur_del = fl.remove_record_on_delete && kind == PG_DELETE;
                ops[j] = cur_del ? OP_SET : (kind_is_retract(kind) ? OP_RETRACT : OP_UPD);
            }
            res_slot = fi[e - 1];
            res_kind = cur_del ? PG_DELETE : PG_INSERT;
        }
        // DropDeleteReader.java:58: only kv.isAdd() survives
        if (emit && fl.drop_delete && kind_is_retract(res_kind)) emit = false;
        px.res_slot[i] = res_slot;
        px.res_kind[i] = (uint8_t)(res_kind | (emit ? 0x80 : 0));
        if (emit) my_emit++;
    }
    __syncthreads();

    // plan entries + per-output sequence number / kind
    int total = 0;
    int o = block_scan_excl(my_emit, px.ws, &total);
    for (int i = p0; i < p1; i++) {
        bool head = (i == 0) || (fk[i] != fk[i - 1]);
        uint16_t entry = (uint16_t)(fi[i] | (ops[i] << kPlanOpShift));
        if (head) {
            entry |= kPlanHead;
            if (px.res_kind[i] & 0x80) {
                entry |= kPlanEmit;
                uint16_t rs = px.res_slot[i];
                pa.tmp_seq[in_base + o] = rs == 0xFFFF ? 0 : seq_s[rs];
                pa.tmp_kind[in_base + o] = (int8_t)(px.res_kind[i] & 0x7f);
                o++;
            }
        }
        pa.plan[in_base + i] = entry;
        fi[i] = entry;                       // keep the finished plan in shared memory for the byte counts
    }
    if (tid == 0) pa.tile_rows[tile] = total;
    __syncthreads();

    // var-len byte counts of the emitted rows
    for (int v = 0; v < pa.n_varlen; v++) {
        const int col = pa.varlen_cols[v];
        const ColDesc cd = pa.cols[col];
        int bytes = 0;
        for (int i = p0; i < p1; i++) {
            uint16_t entry = fi[i];
            if (!(entry & kPlanHead) || !(entry & kPlanEmit)) continue;
            int e = i + 1;
            while (e < n && !(fi[e] & kPlanHead)) e++;
            auto get_ref = [&](int pos) {
                int slot = fi[pos] & kPlanSlotMask;
                int r = run_of_slot(tc.seg, k, slot);
                return MemberRef{r, tc.rstart[r] + (slot - tc.seg[r])};
            };
            int src;
            if (cd.mode == CM_KEY) src = e - 1;
            else src = select_varlen_member(cd, pa.run_cols, pa.n_cols, col, fi, i, e - i, get_ref, err);
            if (src >= 0) {
                MemberRef m = get_ref(src);
                const int32_t *off = pa.run_cols[m.run * pa.n_cols + col].offsets;
                bytes += off[m.row + 1] - off[m.row];
            }
        }
        int tot = 0;
        block_scan_excl(bytes, px.ws, &tot);
        if (tid == 0) pa.tile_bytes[(int64_t)v * pa.n_tiles + tile] = tot;
    }
}

// ------------------------------------------------------------------ scan of tile counts

__global__ void k_scan(const int32_t *tile_rows, const int32_t *tile_bytes, int n_tiles, int n_varlen,
                       int64_t *row_base, int64_t *byte_base, int64_t *totals, int32_t *err) {
    // one block per array (rows, then each var-len column); sequential chunks per thread
    __shared__ int64_t part[1024];
    const int a = blockIdx.x;
    const int32_t *src = a == 0 ? tile_rows : tile_bytes + (int64_t)(a - 1) * n_tiles;
    int64_t *dst = a == 0 ? row_base : byte_base + (int64_t)(a - 1) * n_tiles;
    int per = (n_tiles + blockDim.x - 1) / blockDim.x;
    int b = threadIdx.x * per, e = min(b + per, n_tiles);
    int64_t s = 0;
    for (int i = b; i < e; i++) s += src[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t acc = 0;
        for (int i = 0; i < (int)blockDim.x; i++) { int64_t t = part[i]; part[i] = acc; acc += t; }
        totals[a] = acc;
        if (a > 0 && acc > 0x7fffffffLL) atomicCAS(err, KERR_NONE, KERR_OFFSET_OVERFLOW);
    }
    __syncthreads();
    int64_t acc = part[threadIdx.x];
    for (int i = b; i < e; i++) { dst[i] = acc; acc += src[i]; }
}

// ------------------------------------------------------------------ emit kernel

__device__ __forceinline__ uint64_t load_fixed(const void *data, int width, int64_t row) {
    switch (width) {
        case 1: return ((const uint8_t *)data)[row];
        case 2: return ((const uint16_t *)data)[row];
        case 4: return ((const uint32_t *)data)[row];
        default: return ((const uint64_t *)data)[row];
    }
}
__device__ __forceinline__ void store_fixed(void *data, int width, int64_t row, uint64_t v) {
    switch (width) {
        case 1: ((uint8_t *)data)[row] = (uint8_t)v; break;
        case 2: ((uint16_t *)data)[row] = (uint16_t)v; break;
        case 4: ((uint32_t *)data)[row] = (uint32_t)v; break;
        default: ((uint64_t *)data)[row] = v; break;
    }
}

// Float.compare / Double.compare total order (InternalRowUtils.java:409-414)
__device__ __forceinline__ int java_double_compare(double a, double b) {
    if (a < b) return -1;
    if (a > b) return 1;
    long long x = __double_as_longlong(a), y = __double_as_longlong(b);
    if (a != a) x = 0x7ff8000000000000LL;
    if (b != b) y = 0x7ff8000000000000LL;
    return x == y ? 0 : (x < y ? -1 : 1);
}
__device__ __forceinline__ int java_float_compare(float a, float b) {
    if (a < b) return -1;
    if (a > b) return 1;
    int x = __float_as_int(a), y = __float_as_int(b);
    if (a != a) x = 0x7fc00000;
    if (b != b) y = 0x7fc00000;
    return x == y ? 0 : (x < y ? -1 : 1);
}

__device__ __forceinline__ int64_t sext(uint64_t v, int width) {
    switch (width) {
        case 1: return (int8_t)v;
        case 2: return (int16_t)v;
        case 4: return (int32_t)v;
        default: return (int64_t)v;
    }
}

// acc (op) in for SUM / PRODUCT; integer results wrap exactly like the Java casts
// (FieldSumAgg.java:57-74, FieldProductAgg.java)
__device__ __forceinline__ uint64_t arith(int type, int width, int fn, uint64_t a, uint64_t b, int32_t *err) {
    // fn: 0 add, 1 sub, 2 mul, 3 div
    if (type == PG_DOUBLE) {
        double x = __longlong_as_double((long long)a), y = __longlong_as_double((long long)b), r;
        r = fn == 0 ? __dadd_rn(x, y) : fn == 1 ? __dsub_rn(x, y) : fn == 2 ? __dmul_rn(x, y) : __ddiv_rn(x, y);
        return (uint64_t)__double_as_longlong(r);
    }
    if (type == PG_FLOAT) {
        float x = __int_as_float((int)a), y = __int_as_float((int)b), r;
        r = fn == 0 ? __fadd_rn(x, y) : fn == 1 ? __fsub_rn(x, y) : fn == 2 ? __fmul_rn(x, y) : __fdiv_rn(x, y);
        return (uint32_t)__float_as_int(r);
    }
    int64_t x = sext(a, width), y = sext(b, width);
    uint64_t r;
    if (fn == 0) r = (uint64_t)x + (uint64_t)y;
    else if (fn == 1) r = (uint64_t)x - (uint64_t)y;
    else if (fn == 2) r = (uint64_t)x * (uint64_t)y;
    else {
        if (y == 0) { atomicCAS(err, KERR_NONE, KERR_DIV_ZERO); r = 0; }
        else if (y == -1) r = 0ull - (uint64_t)x;
        else r = (uint64_t)(x / y);
    }
    return r;        // store_fixed truncates to the column width == Java's narrowing cast
}

__device__ __forceinline__ uint64_t negate_fixed(int type, int width, uint64_t a) {
    if (type == PG_DOUBLE) return a ^ 0x8000000000000000ull;
    if (type == PG_FLOAT) return (uint32_t)a ^ 0x80000000u;
    return 0ull - (uint64_t)sext(a, width);
}

__device__ __forceinline__ int compare_fixed(int type, int width, uint64_t a, uint64_t b) {
    if (type == PG_DOUBLE)
        return java_double_compare(__longlong_as_double((long long)a), __longlong_as_double((long long)b));
    if (type == PG_FLOAT) return java_float_compare(__int_as_float((int)a), __int_as_float((int)b));
    int64_t x = sext(a, width), y = sext(b, width);
    return x < y ? -1 : x > y ? 1 : 0;
}

constexpr size_t kEmitSmem = (size_t)kTileMax * (2 + 1 + 4 + 2) + PG_MAX_RUNS * 8 + (PG_MAX_RUNS + 1) * 4 +
                             34 * 4 + 64;

__global__ void __launch_bounds__(kThreads, 4)
k_emit(EmitArgs ea) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t *mrow = (uint32_t *)smem;                        // row inside the run, per merged position
    uint16_t *plan = (uint16_t *)(mrow + kTileMax);
    uint16_t *gstart = plan + kTileMax;                       // per output row: first member position
    uint8_t *mrun = (uint8_t *)(gstart + kTileMax);
    int64_t *rstart = (int64_t *)(mrun + kTileMax);
    int *seg = (int *)(rstart + PG_MAX_RUNS);
    int *ws = seg + PG_MAX_RUNS + 1;

    const int tile = blockIdx.x, tid = threadIdx.x, k = ea.k;
    if (tile == 0) {
        // terminating offset of every var-len output column
        for (int c = tid; c < ea.n_cols; c += blockDim.x)
            if (ea.cols[c].width == 0)
                ea.out_cols[c].offsets[ea.totals[0]] = (int32_t)ea.totals[1 + ea.cols[c].varlen_index];
    }
    if (tid == 0) {
        int acc = 0;
        for (int r = 0; r < k; r++) {
            int64_t b0 = ea.bounds[(int64_t)tile * k + r], b1 = ea.bounds[(int64_t)(tile + 1) * k + r];
            rstart[r] = b0;
            seg[r] = acc;
            acc += (int)(b1 - b0);
        }
        seg[k] = acc;
    }
    __syncthreads();
    const int n = seg[k];
    if (n > kTileMax || n <= 0) return;
    int64_t in_base = 0;
    for (int r = 0; r < k; r++) in_base += rstart[r];
    const int64_t out_base = ea.row_base[tile];

    constexpr int VT = kTileMax / kThreads;
    const int p0 = tid * VT, p1 = min(p0 + VT, n);
    int my = 0;
    for (int i = p0; i < p1; i++) {
        uint16_t e = ea.plan[in_base + i];
        plan[i] = e;
        int slot = e & kPlanSlotMask;
        int r = run_of_slot(seg, k, slot);
        mrun[i] = (uint8_t)r;
        mrow[i] = (uint32_t)(rstart[r] + (slot - seg[r]));
        if ((e & kPlanHead) && (e & kPlanEmit)) my++;
    }
    int n_out = 0;
    int o = block_scan_excl(my, ws, &n_out);
    for (int i = p0; i < p1; i++) {
        uint16_t e = plan[i];
        if ((e & kPlanHead) && (e & kPlanEmit)) gstart[o++] = (uint16_t)i;
    }
    __syncthreads();
    if (n_out == 0) return;

    auto group_end = [&](int g0) {
        int e = g0 + 1;
        while (e < n && !(plan[e] & kPlanHead)) e++;
        return e;
    };

    const int lane = tid & 31;
    // rows are walked in chunks aligned to 32 *global* output rows so that one warp iteration owns
    // exactly one validity word: interior words are plain stores, tile-boundary words use atomicOr
    const int o_shift = (int)(out_base & 31);

    for (int col = 0; col < ea.n_cols; col++) {
        const ColDesc cd = ea.cols[col];
        const pg_out_column oc = ea.out_cols[col];
        const DevColumn *rc = ea.run_cols + col;              // index with run * n_cols
        const int ncs = ea.n_cols;

        if (cd.width > 0) {
            for (int wb = (tid & ~31) - o_shift; wb < n_out; wb += blockDim.x) {   // warp-uniform trip count
                const int ob = wb + lane;
                const bool active = ob >= 0 && ob < n_out;
                uint64_t val = 0;
                bool is_valid = false;
                if (active) {
                    const int g0 = gstart[ob];
                    if (cd.mode == CM_SEQ) {
                        val = (uint64_t)ea.tmp_seq[in_base + ob]; is_valid = true;
                    } else if (cd.mode == CM_KIND) {
                        val = (uint8_t)ea.tmp_kind[in_base + ob]; is_valid = true;
                    } else if (cd.mode == CM_KEY) {
                        val = load_fixed(rc[mrun[g0] * ncs].data, cd.width, mrow[g0]); is_valid = true;
                    } else if (cd.mode == CM_SELECT) {
                        const int ge = group_end(g0);
                        for (int j = ge - 1; j >= g0; j--) {
                            int op = (plan[j] >> kPlanOpShift) & 3;
                            if (op == OP_NOOP) continue;
                            const DevColumn &dc = rc[mrun[j] * ncs];
                            bool v = valid_bit(dc.validity, mrow[j]);
                            if (v) { val = load_fixed(dc.data, cd.width, mrow[j]); is_valid = true; break; }
                            if (op == OP_SET) break;
                        }
                    } else {
                        // CM_FOLD: strict left fold in sequence order (AggregateMergeFunction.java:91-101)
                        const int ge = group_end(g0);
                        bool initialized = false;
                        for (int j = g0; j < ge; j++) {
                            int op = (plan[j] >> kPlanOpShift) & 3;
                            if (op == OP_NOOP) continue;
                            const DevColumn &dc = rc[mrun[j] * ncs];
                            bool v = valid_bit(dc.validity, mrow[j]);
                            uint64_t in = v ? load_fixed(dc.data, cd.width, mrow[j]) : 0;
                            if (op == OP_SET) { val = in; is_valid = v; continue; }
                            if (op == OP_RETRACT) {
                                if (cd.retract == RT_IGNORE) continue;
                                switch (cd.agg) {
                                    case PG_AGG_SUM:       // FieldSumAgg.retract :87-131, negative :133-163
                                        if (!is_valid) { if (v) { val = negate_fixed(cd.type, cd.width, in); is_valid = true; } }
                                        else if (v) val = arith(cd.type, cd.width, 1, val, in, ea.err);
                                        break;
                                    case PG_AGG_PRODUCT:
                                        if (is_valid && v) val = arith(cd.type, cd.width, 3, val, in, ea.err);
                                        break;
                                    case PG_AGG_LAST_VALUE: is_valid = false; val = 0; break;
                                    case PG_AGG_LAST_NON_NULL_VALUE: if (v) { is_valid = false; val = 0; } break;
                                    case PG_AGG_PRIMARY_KEY: val = in; is_valid = v; break;
                                    default: atomicCAS(ea.err, KERR_NONE, KERR_AGG_RETRACT); break;
                                }
                                continue;
                            }
                            switch (cd.agg) {
                                case PG_AGG_SUM: case PG_AGG_PRODUCT:
                                    if (!is_valid || !v) { if (!is_valid) { val = in; is_valid = v; } }
                                    else val = arith(cd.type, cd.width, cd.agg == PG_AGG_SUM ? 0 : 2, val, in, ea.err);
                                    break;
                                case PG_AGG_MAX: case PG_AGG_MIN:
                                    if (!is_valid || !v) { if (!is_valid) { val = in; is_valid = v; } }
                                    else {
                                        int d = compare_fixed(cd.type, cd.width, val, in);
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
                    }
                    store_fixed(oc.data, cd.width, out_base + ob, is_valid ? val : 0);
                }
                if (oc.validity != nullptr) {
                    unsigned mask = __ballot_sync(0xffffffffu, active && is_valid);
                    if (lane == 0) {
                        int64_t word = (out_base + ob) >> 5;      // ob - lane offset: lane 0 is 32-aligned
                        bool full = ob >= 0 && ob + 32 <= n_out;
                        uint32_t *bm = (uint32_t *)oc.validity;
                        if (full) bm[word] = mask;
                        else if (mask) atomicOr(&bm[word], mask);
                    }
                }
            }
        } else {
            // var-len column: choose the source member, scan lengths, copy bytes
            const int64_t byte_base = ea.byte_base[(int64_t)cd.varlen_index * ea.n_tiles + tile];
            int carry = 0;
            auto get_ref = [&](int pos) { return MemberRef{mrun[pos], (int64_t)mrow[pos]}; };
            for (int ob0 = -o_shift; ob0 < n_out; ob0 += blockDim.x) {
                const int ob = ob0 + tid;
                const bool active = ob >= 0 && ob < n_out;
                int src = -1, len = 0;
                const uint8_t *sp = nullptr;
                if (active) {
                    const int g0 = gstart[ob];
                    const int ge = group_end(g0);
                    if (cd.mode == CM_KEY) src = ge - 1;
                    else src = select_varlen_member(cd, ea.run_cols, ncs, col, plan, g0, ge - g0, get_ref, ea.err);
                    if (src >= 0) {
                        const DevColumn &dc = rc[mrun[src] * ncs];
                        int o0 = dc.offsets[mrow[src]];
                        len = dc.offsets[mrow[src] + 1] - o0;
                        sp = (const uint8_t *)dc.data + o0;
                    }
                }
                int tot = 0;
                int off = block_scan_excl(len, ws, &tot) + carry;
                carry += tot;
                if (active) {
                    oc.offsets[out_base + ob] = (int32_t)(byte_base + off);
                    uint8_t *dp = (uint8_t *)oc.data + byte_base + off;
                    for (int b = 0; b < len; b++) dp[b] = sp[b];
                }
                if (oc.validity != nullptr) {
                    unsigned mask = __ballot_sync(0xffffffffu, active && src >= 0);
                    if (lane == 0) {
                        int64_t word = (out_base + ob) >> 5;
                        bool full = ob >= 0 && ob + 32 <= n_out;
                        uint32_t *bm = (uint32_t *)oc.validity;
                        if (full) bm[word] = mask;
                        else if (mask) atomicOr(&bm[word], mask);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------ launchers

static bool g_attr_done = false;
static void set_attrs() {
    if (g_attr_done) return;
    cudaFuncSetAttribute(k_merge_keys, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmem);
    cudaFuncSetAttribute(k_plan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPlanSmem);
    cudaFuncSetAttribute(k_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEmitSmem);
    g_attr_done = true;
}

void launch_partition(const MergeLaunch &ml, const LevelView &lv, const uint64_t *splitter_keys,
                      int64_t n_splitter_keys, int q, int n_tiles, int64_t *bounds) {
    (void)n_splitter_keys;
    int64_t total = (int64_t)(n_tiles + 1) * ml.k;
    int blocks = (int)((total + 127) / 128);
    k_partition<<<blocks, 128, 0, ml.stream>>>(ml.k, ml.key, ml.key_ptrs, lv, splitter_keys, q, n_tiles, bounds);
}

void launch_merge_keys(const MergeLaunch &ml, const LevelView &lv, const int64_t *bounds, int n_tiles,
                       uint64_t *sorted_keys) {
    set_attrs();
    k_merge_keys<<<n_tiles, kThreads, kTileSmem, ml.stream>>>(ml.k, ml.key, ml.key_ptrs, lv, bounds, sorted_keys,
                                                                ml.err);
}

void launch_plan(const MergeLaunch &ml, const PlanArgs &pa) {
    set_attrs();
    k_plan<<<pa.n_tiles, kThreads, kPlanSmem, ml.stream>>>(ml.k, ml.key, ml.key_ptrs, pa, ml.err);
}

void launch_scan(cudaStream_t stream, const int32_t *tile_rows, const int32_t *tile_bytes, int n_tiles,
                 int n_varlen, int64_t *row_base, int64_t *byte_base, int64_t *totals, int32_t *err) {
    k_scan<<<1 + n_varlen, 1024, 0, stream>>>(tile_rows, tile_bytes, n_tiles, n_varlen, row_base, byte_base,
                                               totals, err);
}

void launch_emit(const EmitArgs &ea) {
    set_attrs();
    k_emit<<<ea.n_tiles, kThreads, kEmitSmem, ea.stream>>>(ea);
}

}  // namespace pg
