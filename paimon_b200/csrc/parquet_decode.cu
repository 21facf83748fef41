// parquet_decode.cu — Parquet column-chunk decode on the device, feeding the merge without leaving HBM.
//
// Reference being replaced (paths under /root/reference/paimon-format/src/main/java/org/apache/paimon/format/):
//   parquet/ParquetReaderFactory.java:113-148          createReader (footer, schema clip, vectors)
//   parquet/reader/VectorizedParquetRecordReader.java:178-241   nextBatch / row-group loop
//   parquet/reader/VectorizedColumnReader.java:143-383 page loop: V1 = [def RLE][values], V2 = separate levels
//   parquet/reader/VectorizedRleValuesReader.java:928-1019     RLE / bit-packed hybrid
//   parquet/reader/VectorizedPlainValuesReader.java:117-189    PLAIN fixed width and BYTE_ARRAY
//   parquet/ParquetSchemaConverter.java:76-160         TINYINT/SMALLINT/INT/DATE -> INT32, BIGINT -> INT64, ...
// The reference fills 1024-row ColumnVectors on one CPU thread; here the whole file is staged to HBM once,
// the host walks the (tiny) Thrift page headers, and every page is decoded by its own warp / CTA into the
// Arrow-layout columns the merge kernels consume.  ABI v1 scope: flat schema, INT32 / INT64 / FLOAT / DOUBLE /
// BYTE_ARRAY, PLAIN and RLE/PLAIN_DICTIONARY encodings, data pages V1 and V2, max definition level 1,
// UNCOMPRESSED pages.  Anything else is refused with PG_ERR_UNSUPPORTED (no CPU fallback).
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <unordered_map>

#include "device_utils.cuh"
#include "parquet_meta.h"

namespace pg {

struct PqPageJob {
    const uint8_t *body;      // device: page body (after the Thrift header)
    int32_t body_len;
    int32_t num_values;       // rows in the page (flat schema: values incl. nulls)
    int32_t col;
    int32_t page_type;        // pq::P_DATA / pq::P_DATA_V2
    int32_t is_dict;          // values are dictionary ids
    int32_t max_def;          // 0 = REQUIRED, 1 = OPTIONAL
    int32_t v2_def_len;
    int32_t dict;             // index into PqDictJob, -1 = none
    int64_t row0;             // absolute row of the page's first value
};

struct PqDictJob {
    const uint8_t *body;
    int32_t body_len;
    int32_t num_values;
    int32_t col;
    int32_t pad;
    int64_t entry_base;       // first entry in the dict_ptr / dict_len tables (BYTE_ARRAY only)
};

struct PqCol {
    int32_t phys;             // pq::PhysType
    int32_t phys_width;       // 4 / 8, 0 for BYTE_ARRAY
    int32_t out_width;        // bytes of the output type, 0 for var-len
    int32_t nullable;
    void *out_data;
    int32_t *out_offsets;
    uint32_t *out_validity;   // zeroed; bits are OR-ed in
    uint8_t *defs;            // scratch [n_rows], optional columns
    int32_t *ids;             // scratch [n_rows], dictionary ids by (page row0 + value ordinal)
    const uint8_t **vptr;     // scratch [n_rows], BYTE_ARRAY PLAIN: value payload pointer by (row0 + ordinal)
    int32_t *vlen;            // scratch [n_rows]
    const uint8_t **rowsrc;   // scratch [n_rows], var-len: payload pointer per row
};

// page-derived values written by k_pq_hybrid
struct PqPageState {
    int32_t values_off;       // offset of the values section inside the body
    int32_t n_nonnull;
};

__device__ __forceinline__ uint32_t pq_varint(const uint8_t *&p, const uint8_t *end) {
    uint32_t v = 0;
    int shift = 0;
    while (p < end) {
        uint8_t b = *p++;
        v |= (uint32_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) break;
        shift += 7;
    }
    return v;
}

// Warp-cooperative RLE / bit-packed hybrid decode (VectorizedRleValuesReader.java:928-1019).
// Calls out(i, value) for i in [0, count).  Returns the number of values equal to `count_eq` (for def levels).
template <typename Out>
__device__ int pq_hybrid_decode(const uint8_t *p, const uint8_t *end, int bw, int count, uint32_t count_eq, Out out) {
    const int lane = threadIdx.x & 31;
    const uint32_t mask = bw >= 32 ? 0xffffffffu : ((1u << bw) - 1);
    int pos = 0, matches = 0;
    while (pos < count) {
        if (bw == 0) {                        // a zero-width stream encodes only zeros
            for (int i = pos + lane; i < count; i += 32) { out(i, 0u); matches += (count_eq == 0); }
            break;
        }
        if (p >= end) break;
        uint32_t h = pq_varint(p, end);
        if (h & 1) {
            int groups = (int)(h >> 1);
            int nvals = groups * 8;
            for (int i = lane; i < nvals && pos + i < count; i += 32) {
                int64_t bit = (int64_t)i * bw;
                const uint8_t *q = p + (bit >> 3);
                uint64_t w = 0;
#pragma unroll
                for (int b = 0; b < 5; b++)
                    if (q + b < end) w |= (uint64_t)q[b] << (8 * b);
                uint32_t v = (uint32_t)(w >> (bit & 7)) & mask;
                out(pos + i, v);
                matches += (v == count_eq);
            }
            p += (int64_t)groups * bw;
            pos += nvals;
        } else {
            int run = (int)(h >> 1);
            uint32_t v = 0;
            int nb = (bw + 7) / 8;
            for (int b = 0; b < nb; b++)
                if (p + b < end) v |= (uint32_t)p[b] << (8 * b);
            p += nb;
            for (int i = lane; i < run && pos + i < count; i += 32) { out(pos + i, v); matches += (v == count_eq); }
            pos += run;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) matches += __shfl_xor_sync(0xffffffffu, matches, d);
    return matches;
}

// one warp per data page: definition levels and dictionary ids -> scratch
__global__ void k_pq_hybrid(const PqPageJob *jobs, int n_jobs, const PqCol *cols, PqPageState *state) {
    int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= n_jobs) return;
    const PqPageJob job = jobs[j];
    const PqCol col = cols[job.col];
    const uint8_t *body = job.body, *end = job.body + job.body_len;
    int values_off = 0, n_nonnull = job.num_values;
    if (job.max_def > 0) {
        const uint8_t *dp;
        int dlen;
        if (job.page_type == pq::P_DATA_V2) { dp = body; dlen = job.v2_def_len; values_off = dlen; }
        else {
            dlen = (int)((uint32_t)body[0] | ((uint32_t)body[1] << 8) | ((uint32_t)body[2] << 16) | ((uint32_t)body[3] << 24));
            dp = body + 4;
            values_off = 4 + dlen;
        }
        uint8_t *defs = col.defs + job.row0;
        const uint8_t *dend = dp + dlen < end ? dp + dlen : end;
        n_nonnull = pq_hybrid_decode(dp, dend, 1, job.num_values, 1u,
                                     [&](int i, uint32_t v) { defs[i] = (uint8_t)v; });
    }
    if (job.is_dict) {
        const uint8_t *vp = body + values_off;
        int bw = vp < end ? vp[0] : 0;
        int32_t *ids = col.ids + job.row0;
        pq_hybrid_decode(vp + 1, end, bw, n_nonnull, 0xffffffffu, [&](int i, uint32_t v) { ids[i] = (int32_t)v; });
    }
    if ((threadIdx.x & 31) == 0) state[j] = PqPageState{values_off, n_nonnull};
}

// One warp per PLAIN BYTE_ARRAY page (data or dictionary): walk the [len:4][bytes] stream
// (VectorizedPlainValuesReader.java:275-288).  The lengths are embedded in the stream, so the walk itself is
// sequential; the warp stages the page through shared memory in 2 KiB windows (coalesced 16-byte loads), lane 0
// walks the window at shared-memory latency, and the (pointer, length) pairs go out coalesced.
constexpr int kWalkWindow = 2048;
constexpr int kWalkWarps = 8;

__device__ void pq_walk_stream(const uint8_t *p, const uint8_t *end, int count, const uint8_t **out_ptr,
                               int32_t *out_len, uint8_t *win, uint32_t *w_off, int32_t *w_len) {
    const int lane = threadIdx.x & 31;
    int done = 0;
    while (done < count && p + 4 <= end) {
        // window = [base, base + kWalkWindow + 16) clipped to the page, base 16-byte aligned
        const uint8_t *base = (const uint8_t *)((uintptr_t)p & ~(uintptr_t)15);
        const int skip = (int)(p - base);
        const int avail = (int)min((int64_t)(end - base), (int64_t)(kWalkWindow + 16));
        for (int o = lane * 16; o < avail; o += 32 * 16) *(uint4 *)(win + o) = *(const uint4 *)(base + o);
        __syncwarp();
        int nfound = 0, q = skip;
        if (lane == 0) {
            while (done + nfound < count && q + 4 <= avail && nfound < 256) {
                int32_t len = (int32_t)((uint32_t)win[q] | ((uint32_t)win[q + 1] << 8) | ((uint32_t)win[q + 2] << 16) |
                                        ((uint32_t)win[q + 3] << 24));
                w_off[nfound] = (uint32_t)(q + 4);
                w_len[nfound] = len;
                nfound++;
                q += 4 + len;
                if (len < 0) { q = avail; break; }            // corrupt length: stop
            }
        }
        nfound = __shfl_sync(0xffffffffu, nfound, 0);
        q = __shfl_sync(0xffffffffu, q, 0);
        __syncwarp();
        for (int i = lane; i < nfound; i += 32) {
            out_ptr[done + i] = base + w_off[i];
            out_len[done + i] = w_len[i];
        }
        __syncwarp();
        if (nfound == 0) break;                               // a length word straddles the page end: malformed
        done += nfound;
        p = base + q;
    }
}

__global__ void __launch_bounds__(kWalkWarps * 32)
k_pq_walk_bytes(const PqPageJob *jobs, int n_jobs, const PqDictJob *dicts, int n_dicts, const PqCol *cols,
                const PqPageState *state, const uint8_t **dict_ptr, int32_t *dict_len) {
    __shared__ __align__(16) uint8_t s_win[kWalkWarps][kWalkWindow + 32];
    __shared__ uint32_t s_off[kWalkWarps][256];
    __shared__ int32_t s_len[kWalkWarps][256];
    const int w = threadIdx.x >> 5;
    const int t = blockIdx.x * kWalkWarps + w;
    if (t < n_jobs) {
        const PqPageJob job = jobs[t];
        const PqCol col = cols[job.col];
        if (col.phys != pq::T_BYTE_ARRAY || job.is_dict) return;
        pq_walk_stream(job.body + state[t].values_off, job.body + job.body_len, state[t].n_nonnull,
                       col.vptr + job.row0, col.vlen + job.row0, s_win[w], s_off[w], s_len[w]);
    } else if (t < n_jobs + n_dicts) {
        const PqDictJob dj = dicts[t - n_jobs];
        if (cols[dj.col].phys != pq::T_BYTE_ARRAY) return;
        pq_walk_stream(dj.body, dj.body + dj.body_len, dj.num_values, dict_ptr + dj.entry_base,
                       dict_len + dj.entry_base, s_win[w], s_off[w], s_len[w]);
    }
}

__device__ __forceinline__ uint64_t pq_load_unaligned(const uint8_t *p, int w) {
    uint64_t v = 0;
    for (int b = 0; b < w; b++) v |= (uint64_t)p[b] << (8 * b);
    return v;
}

// one CTA per data page: rows -> output values / lengths + validity
__global__ void __launch_bounds__(256)
k_pq_assemble(const PqPageJob *jobs, const PqDictJob *dicts, const PqCol *cols, const PqPageState *state,
              const uint8_t *const *dict_ptr, const int32_t *dict_len) {
    __shared__ int ws[34];
    const PqPageJob job = jobs[blockIdx.x];
    const PqCol col = cols[job.col];
    const PqPageState st = state[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31;
    const uint8_t *values = job.body + st.values_off;
    const PqDictJob *dj = job.dict >= 0 ? &dicts[job.dict] : nullptr;
    int carry = 0;
    for (int base = 0; base < job.num_values; base += blockDim.x) {
        const int i = base + tid;
        const bool in = i < job.num_values;
        const int64_t row = job.row0 + i;
        int valid = in ? (job.max_def ? col.defs[row] : 1) : 0;
        int tot = 0;
        int ord = carry + block_scan_excl(valid, ws, &tot);
        carry += tot;
        if (in) {
            if (col.phys != pq::T_BYTE_ARRAY) {
                uint64_t v = 0;
                if (valid) {
                    const uint8_t *src = job.is_dict ? dj->body + (int64_t)col.ids[job.row0 + ord] * col.phys_width
                                                     : values + (int64_t)ord * col.phys_width;
                    v = pq_load_unaligned(src, col.phys_width);
                }
                store_fixed(col.out_data, col.out_width, row, v);     // narrowing keeps the low bytes (INT32 -> TINYINT)
            } else {
                const uint8_t *src = nullptr;
                int len = 0;
                if (valid) {
                    if (job.is_dict) {
                        int64_t e = dj->entry_base + col.ids[job.row0 + ord];
                        src = dict_ptr[e];
                        len = dict_len[e];
                    } else {
                        src = col.vptr[job.row0 + ord];
                        len = col.vlen[job.row0 + ord];
                    }
                }
                col.rowsrc[row] = src;
                col.out_offsets[row + 1] = len;                      // lengths now, prefix-summed afterwards
            }
        }
        if (col.out_validity != nullptr) {
            unsigned m = __ballot_sync(0xffffffffu, valid != 0);
            if (lane == 0 && m) {
                int64_t r0 = job.row0 + base + (tid & ~31);
                int sh = (int)(r0 & 31);
                atomicOr(&col.out_validity[r0 >> 5], m << sh);
                if (sh) atomicOr(&col.out_validity[(r0 >> 5) + 1], m >> (32 - sh));
            }
        }
    }
}

// ---- Snappy page decompression (parquet-mr hands compressed pages to snappy-java 1.1.10.8; the format restated
// here is the public Snappy format description: a varint uncompressed length, then literal and copy elements).
// One warp per page: the element stream is parsed by all lanes in lock step, the bytes of a literal / copy are
// moved lane-parallel.  A copy may overlap its own output (offset < length): byte i comes from
// out - offset + (i mod offset), which always lies in front of the copy.
struct SnappyJob {
    const uint8_t *src;       // device: [prefix bytes copied verbatim (data page V2 levels)][snappy stream]
    uint8_t *dst;
    int32_t src_len, dst_len, prefix, pad;
};
__global__ void k_pq_snappy(const SnappyJob *jobs, int n_jobs, int32_t *err) {
    const int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= n_jobs) return;
    const SnappyJob j = jobs[w];
    for (int i = lane; i < j.prefix; i += 32) j.dst[i] = j.src[i];
    const uint8_t *src = j.src + j.prefix;
    uint8_t *dst = j.dst + j.prefix;
    const int n_src = j.src_len - j.prefix, n_dst = j.dst_len - j.prefix;
    int pos = 0, out = 0;
    // preamble: uncompressed length
    uint32_t ulen = 0;
    for (int sh = 0; pos < n_src && sh < 35; sh += 7) {
        const uint8_t b = src[pos++];
        ulen |= (uint32_t)(b & 0x7f) << sh;
        if (!(b & 0x80)) break;
    }
    bool bad = (int)ulen != n_dst;
    while (!bad && pos < n_src) {
        const uint32_t tag = src[pos];
        int len, offset = 0;
        if ((tag & 3) == 0) {
            len = (int)(tag >> 2) + 1;
            pos += 1;
            if (len > 60) {
                const int extra = len - 60;
                if (pos + extra > n_src) { bad = true; break; }
                len = 0;
                for (int b = 0; b < extra; b++) len |= (int)src[pos + b] << (8 * b);
                len += 1;
                pos += extra;
            }
            if (len < 0 || pos + len > n_src || out + len > n_dst) { bad = true; break; }
            for (int i = lane; i < len; i += 32) dst[out + i] = src[pos + i];
            pos += len;
        } else {
            if ((tag & 3) == 1) {
                if (pos + 2 > n_src) { bad = true; break; }
                len = 4 + (int)((tag >> 2) & 7);
                offset = (int)((tag >> 5) << 8) | src[pos + 1];
                pos += 2;
            } else if ((tag & 3) == 2) {
                if (pos + 3 > n_src) { bad = true; break; }
                len = 1 + (int)(tag >> 2);
                offset = src[pos + 1] | (src[pos + 2] << 8);
                pos += 3;
            } else {
                if (pos + 5 > n_src) { bad = true; break; }
                len = 1 + (int)(tag >> 2);
                offset = (int)(src[pos + 1] | (src[pos + 2] << 8) | (src[pos + 3] << 16) | ((uint32_t)src[pos + 4] << 24));
                pos += 5;
            }
            if (offset <= 0 || offset > out || out + len > n_dst) { bad = true; break; }
            const uint8_t *from = dst + out - offset;
            for (int i = lane; i < len; i += 32) dst[out + i] = from[i % offset];
        }
        out += len;
        __syncwarp();                                    // later copies may read what other lanes just wrote
    }
    if ((bad || out != n_dst) && lane == 0) atomicCAS(err, KERR_NONE, KERR_BAD_PAGE);
}

// ---- DELTA_BINARY_PACKED (VectorizedDeltaBinaryPackedReader.java; the layout restated here is the public Parquet
// encoding specification): <block size> <miniblocks per block> <total count> <first value> then per block
// <min delta> <bit width per miniblock> <bit-packed miniblocks>.  One warp per page expands the values into a
// PLAIN image (the level bytes in front are copied), so the PLAIN assemble path reads the page afterwards.
struct DeltaJob {
    const uint8_t *src;       // device: page body
    uint8_t *dst;
    int32_t src_len, prefix;  // prefix = level bytes in front of the values
    int32_t width, max_values;
};
__device__ __forceinline__ uint64_t dl_varint(const uint8_t *p, int n, int &pos) {
    uint64_t v = 0;
    for (int sh = 0; pos < n && sh < 70; sh += 7) {
        const uint8_t b = p[pos++];
        v |= (uint64_t)(b & 0x7f) << sh;
        if (!(b & 0x80)) break;
    }
    return v;
}
__global__ void k_pq_delta(const DeltaJob *jobs, int n_jobs, int32_t *err) {
    const int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= n_jobs) return;
    const DeltaJob j = jobs[w];
    for (int i = lane; i < j.prefix; i += 32) j.dst[i] = j.src[i];
    const uint8_t *p = j.src + j.prefix;
    const int n = j.src_len - j.prefix;
    uint8_t *out = j.dst + j.prefix;
    int pos = 0;
    const int block_size = (int)dl_varint(p, n, pos);
    const int n_mini = (int)dl_varint(p, n, pos);
    const int64_t total = (int64_t)dl_varint(p, n, pos);
    const uint64_t zz = dl_varint(p, n, pos);
    uint64_t last = (zz >> 1) ^ (0 - (zz & 1));               // first value
    bool bad = n_mini <= 0 || block_size <= 0 || block_size % n_mini != 0 || total > j.max_values || total < 0;
    const int mini = bad ? 1 : block_size / n_mini;
    if (!bad && total > 0 && lane == 0) {
        if (j.width == 8) memcpy(out, &last, 8); else { uint32_t x = (uint32_t)last; memcpy(out, &x, 4); }
    }
    int64_t done = 1;
    while (!bad && done < total) {
        const uint64_t mz = dl_varint(p, n, pos);
        const uint64_t min_delta = (mz >> 1) ^ (0 - (mz & 1));
        const int bw_pos = pos;
        pos += n_mini;
        if (pos > n) { bad = true; break; }
        for (int m = 0; m < n_mini && done < total; m++) {
            const int bw = p[bw_pos + m];
            if (bw > 64 || pos + (int64_t)mini * bw / 8 > n) { bad = true; break; }
            for (int v0 = 0; v0 < mini && done < total; v0 += 32) {
                const int v = v0 + lane;
                uint64_t d = 0;
                if (v < mini && bw > 0) {
                    const int64_t bit = (int64_t)v * bw;
                    const uint8_t *q = p + pos + (bit >> 3);
                    const int sh = (int)(bit & 7);
                    // up to 9 bytes hold the value
                    uint64_t lo = 0;
                    const int nb = (sh + bw + 7) >> 3;
                    for (int b = 0; b < nb && b < 8; b++) lo |= (uint64_t)q[b] << (8 * b);
                    d = lo >> sh;
                    if (nb > 8) d |= (uint64_t)q[8] << (64 - sh);
                    if (bw < 64) d &= ((uint64_t)1 << bw) - 1;
                }
                uint64_t x = v < mini ? d + min_delta : 0;
                // inclusive scan of the deltas over the warp, then the running value
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
                    if (lane >= o) x += y;
                }
                const uint64_t val = last + x;
                const int64_t idx = done + lane;
                if (v < mini && idx < total) {
                    if (j.width == 8) memcpy(out + idx * 8, &val, 8);
                    else { uint32_t t = (uint32_t)val; memcpy(out + idx * 4, &t, 4); }
                }
                const int cnt = min(32, mini - v0);
                last = __shfl_sync(0xffffffffu, val, cnt - 1);
                done += cnt;
            }
            pos += mini * bw / 8;
        }
    }
    if (bad && lane == 0) atomicCAS(err, KERR_NONE, KERR_BAD_PAGE);
}

// ---- device-wide inclusive scan of int32 (three small kernels; offsets of one var-len column)
__global__ void k_scan_block_sums(const int32_t *data, int64_t n, int64_t *block_sums) {
    __shared__ int64_t sh[256];
    int64_t b0 = (int64_t)blockIdx.x * 4096;
    int64_t s = 0;
    for (int i = threadIdx.x; i < 4096 && b0 + i < n; i += blockDim.x) s += data[b0 + i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) { if (threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d]; __syncthreads(); }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = sh[0];
}
__global__ void k_scan_block_prefix(int64_t *block_sums, int64_t n_blocks, int32_t *err) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int64_t acc = 0;
        for (int64_t i = 0; i < n_blocks; i++) { int64_t t = block_sums[i]; block_sums[i] = acc; acc += t; }
        if (acc > 0x7fffffffLL) atomicCAS(err, KERR_NONE, KERR_OFFSET_OVERFLOW);
    }
}
__global__ void __launch_bounds__(256) k_scan_apply(int32_t *data, int64_t n, const int64_t *block_sums) {
    __shared__ int ws[34];
    int64_t b0 = (int64_t)blockIdx.x * 4096;
    int carry = (int)block_sums[blockIdx.x];
    for (int base = 0; base < 4096; base += 256) {
        int64_t i = b0 + base + threadIdx.x;
        int v = i < n ? data[i] : 0;
        int tot = 0;
        int ex = block_scan_excl(v, ws, &tot);
        if (i < n) data[i] = carry + ex + v;
        carry += tot;
    }
}

// payload copy: 8 lanes per row
__global__ void k_pq_copy_bytes(const uint8_t *const *rowsrc, const int32_t *offsets, uint8_t *out, int64_t n_rows) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (row >= n_rows) return;
    const uint8_t *src = rowsrc[row];
    int o0 = offsets[row], o1 = offsets[row + 1];
    for (int b = o0 + (threadIdx.x & 7); b < o1; b += 8) out[b] = src[b - o0];
}

// ------------------------------------------------------------------ host side

struct PqReader {
    const Schema *schema = nullptr;
    uint64_t schema_h = 0;
    pq::FileMetaData meta;
    std::vector<uint8_t> file;             // host copy (page headers are parsed from it)
    std::vector<PqPageJob> jobs;           // body pointers hold FILE OFFSETS until decode time
    std::vector<PqDictJob> dicts;
    std::vector<int64_t> dict_entries_per_col;
    int64_t n_rows = 0;
    int64_t dict_entries = 0;
    // compressed page bodies (Snappy), decompressed on the device before the decode kernels run
    struct Unc { int64_t src_off; int32_t src_len, dst_len, prefix; int64_t dst_off; };
    std::vector<Unc> unc;
    std::vector<int32_t> job_unc, dict_unc;    // per page / dictionary job: index into unc, -1 = stored uncompressed
    int64_t unc_bytes = 0;
    // DELTA_BINARY_PACKED pages, expanded to PLAIN images on the device
    struct Delta { int64_t src_off; int32_t src_len, prefix, width, max_values; int64_t dst_off; };
    std::vector<Delta> delta;
    std::vector<int32_t> job_delta;            // per page job: index into delta, -1 = none
    int64_t delta_bytes = 0;
    float ms_decode = 0;
    int launches = 0;
};

static std::mutex g_pq_mu;
static std::unordered_map<uint64_t, std::unique_ptr<PqReader>> g_pq;
static uint64_t g_pq_next = 1;

static int out_width_of(int t) {
    switch (t) {
        case PG_INT8: case PG_BOOL: return 1;
        case PG_INT16: return 2;
        case PG_INT32: case PG_FLOAT: return 4;
        case PG_INT64: case PG_DOUBLE: return 8;
        default: return 0;
    }
}

// ParquetSchemaConverter.java:76-160 — which physical type a Paimon column must have in the file
static bool phys_compatible(int pg_t, int phys) {
    switch (pg_t) {
        case PG_INT8: case PG_INT16: case PG_INT32: return phys == pq::T_INT32;
        case PG_INT64: return phys == pq::T_INT64;
        case PG_FLOAT: return phys == pq::T_FLOAT;
        case PG_DOUBLE: return phys == pq::T_DOUBLE;
        case PG_STRING: case PG_BINARY: return phys == pq::T_BYTE_ARRAY;
        default: return false;
    }
}

Schema *schema_from_handle(uint64_t h);                 // api.cu
void *device_buffer_take(size_t bytes, size_t *got);    // api.cu: recycled device buffers
void device_buffer_give(void *p, size_t bytes);
cudaStream_t thread_stream();                           // api.cu: the calling thread's non-blocking stream
uint64_t register_run(std::unique_ptr<Run> run);        // api.cu
pg_status require_device();                             // api.cu

static pg_status pq_open(uint64_t schema_h, const uint8_t *bytes, int64_t size, uint64_t *out) {
    Schema *s = schema_from_handle(schema_h);
    if (!s || !bytes || !out) return fail(PG_ERR_INVALID, "bad schema handle or null argument");
    auto rd = std::make_unique<PqReader>();
    rd->schema = s;
    rd->schema_h = schema_h;
    try {
        rd->meta = pq::parse_footer(bytes, size);
    } catch (const std::exception &e) {
        return fail(PG_ERR_FORMAT, e.what());
    }
    const pq::FileMetaData &m = rd->meta;
    const int nc = s->n_cols();
    if (m.schema.empty() || m.schema[0].num_children != nc || (int)m.schema.size() != nc + 1)
        return fail(PG_ERR_UNSUPPORTED, "parquet: only flat schemas whose columns match the KeyValue file schema "
                                        "[_KEY_*, _SEQUENCE_NUMBER, _VALUE_KIND, value...] are decoded on device");
    for (int c = 0; c < nc; c++) {
        const pq::SchemaElement &e = m.schema[c + 1];
        if (e.num_children != 0 || e.repetition == pq::R_REPEATED)
            return fail(PG_ERR_UNSUPPORTED, "parquet: nested / repeated column " + e.name);
        if (!phys_compatible(s->field(c).type, e.type))
            return fail(PG_ERR_UNSUPPORTED, "parquet: column " + e.name + " has a physical type the device decoder "
                                            "does not map to the table type");
    }
    rd->file.assign(bytes, bytes + size);
    rd->n_rows = m.num_rows;
    rd->dict_entries_per_col.assign(nc, 0);
    int64_t row0 = 0;
    for (const pq::RowGroup &g : m.row_groups) {
        if ((int)g.columns.size() != nc) return fail(PG_ERR_FORMAT, "parquet: row group with a different column count");
        for (int c = 0; c < nc; c++) {
            const pq::ColumnChunk &cc = g.columns[c];
            if (cc.codec != pq::C_UNCOMPRESSED && cc.codec != pq::C_SNAPPY)
                return fail(PG_ERR_UNSUPPORTED, "parquet: compression codec " + std::to_string(cc.codec) +
                                                " is not decoded on device (UNCOMPRESSED and SNAPPY are); write with "
                                                "'file.compression'='none' / 'snappy' or let the Java side decompress");
            const bool snappy = cc.codec == pq::C_SNAPPY;
            // returns the index of the decompression item of a page body, or -1 when it is stored as is
            auto add_unc = [&](int64_t body, const pq::PageHeader &h, int32_t prefix, bool compressed) -> int32_t {
                if (!snappy || !compressed) return -1;
                PqReader::Unc u{body, h.compressed_size, h.uncompressed_size, prefix, rd->unc_bytes};
                rd->unc_bytes += ((int64_t)h.uncompressed_size + 63) & ~(int64_t)63;
                rd->unc.push_back(u);
                return (int32_t)rd->unc.size() - 1;
            };
            const int max_def = m.schema[c + 1].repetition == pq::R_OPTIONAL ? 1 : 0;
            int64_t pos = cc.start(), vals = 0, page_row = row0;
            int dict_index = -1;
            while (vals < cc.num_values) {
                if (pos < 4 || pos >= size) return fail(PG_ERR_FORMAT, "parquet: page offset out of range");
                pq::PageHeader h;
                try {
                    h = pq::parse_page_header(bytes + pos, size - pos);
                } catch (const std::exception &e) {
                    return fail(PG_ERR_FORMAT, e.what());
                }
                int64_t body = pos + h.header_size;
                if (body + h.compressed_size > size) return fail(PG_ERR_FORMAT, "parquet: truncated page");
                if (h.type == pq::P_DICTIONARY) {
                    if (h.encoding != pq::E_PLAIN && h.encoding != pq::E_PLAIN_DICTIONARY)
                        return fail(PG_ERR_UNSUPPORTED, "parquet: dictionary page encoding");
                    PqDictJob dj{};
                    dj.body = (const uint8_t *)(uintptr_t)body;
                    dj.body_len = h.compressed_size;
                    dj.num_values = h.num_values;
                    dj.col = c;
                    dj.entry_base = rd->dict_entries;
                    if (cc.type == pq::T_BYTE_ARRAY) rd->dict_entries += h.num_values;
                    dict_index = (int)rd->dicts.size();
                    rd->dict_unc.push_back(add_unc(body, h, 0, true));
                    if (rd->dict_unc.back() >= 0) dj.body_len = h.uncompressed_size;
                    rd->dicts.push_back(dj);
                } else if (h.type == pq::P_DATA || h.type == pq::P_DATA_V2) {
                    bool is_dict = h.encoding == pq::E_PLAIN_DICTIONARY || h.encoding == pq::E_RLE_DICTIONARY;
                    const bool is_delta = h.encoding == pq::E_DELTA_BINARY_PACKED &&
                                          (cc.type == pq::T_INT32 || cc.type == pq::T_INT64) && !snappy;
                    if (!is_dict && !is_delta && h.encoding != pq::E_PLAIN)
                        return fail(PG_ERR_UNSUPPORTED, "parquet: value encoding " + std::to_string(h.encoding) +
                                                        " (PLAIN, dictionary and DELTA_BINARY_PACKED on uncompressed "
                                                        "integer pages are decoded on device)");
                    if (is_dict && dict_index < 0) return fail(PG_ERR_FORMAT, "parquet: dictionary-encoded page without dictionary");
                    if (h.type == pq::P_DATA_V2 && h.rep_levels_byte_length != 0)
                        return fail(PG_ERR_UNSUPPORTED, "parquet: repetition levels");
                    PqPageJob pj{};
                    pj.body = (const uint8_t *)(uintptr_t)body;
                    pj.body_len = h.compressed_size;
                    pj.num_values = h.num_values;
                    pj.col = c;
                    pj.page_type = h.type;
                    pj.is_dict = is_dict;
                    pj.max_def = max_def;
                    pj.v2_def_len = h.def_levels_byte_length;
                    pj.dict = is_dict ? dict_index : -1;
                    pj.row0 = page_row;
                    // V1: the whole body (levels + values) is one compressed block; V2: the levels stay as they
                    // are in front of the (optionally) compressed values
                    const int32_t prefix = h.type == pq::P_DATA_V2 ? h.def_levels_byte_length + h.rep_levels_byte_length : 0;
                    rd->job_unc.push_back(add_unc(body, h, prefix, h.type == pq::P_DATA_V2 ? h.is_compressed : true));
                    if (rd->job_unc.back() >= 0) pj.body_len = h.uncompressed_size;
                    int32_t di = -1;
                    if (is_delta) {
                        // level bytes in front of the values: V1 = 4-byte length + RLE levels (OPTIONAL only)
                        int32_t lv = prefix;
                        if (h.type == pq::P_DATA && max_def > 0) {
                            if (h.compressed_size < 4) return fail(PG_ERR_FORMAT, "parquet: truncated page");
                            uint32_t l4;
                            memcpy(&l4, bytes + body, 4);
                            lv = 4 + (int32_t)l4;
                        }
                        if (lv < 0 || lv > h.compressed_size) return fail(PG_ERR_FORMAT, "parquet: bad level length");
                        const int32_t width = cc.type == pq::T_INT32 ? 4 : 8;
                        PqReader::Delta d{body, h.compressed_size, lv, width, h.num_values, rd->delta_bytes};
                        rd->delta_bytes += ((int64_t)lv + (int64_t)h.num_values * width + 63) & ~(int64_t)63;
                        rd->delta.push_back(d);
                        di = (int32_t)rd->delta.size() - 1;
                        pj.body_len = lv + h.num_values * width;
                    }
                    rd->job_delta.push_back(di);
                    rd->jobs.push_back(pj);
                    page_row += h.num_values;
                    vals += h.num_values;
                }                                        // index pages are skipped
                pos = body + h.compressed_size;
            }
            if (page_row - row0 != g.num_rows) return fail(PG_ERR_FORMAT, "parquet: page row counts do not add up");
        }
        row0 += g.num_rows;
    }
    if (row0 != m.num_rows) return fail(PG_ERR_FORMAT, "parquet: row group row counts do not add up");
    std::lock_guard<std::mutex> lk(g_pq_mu);
    uint64_t h = (5ull << 56) | g_pq_next++;
    g_pq[h] = std::move(rd);
    *out = h;
    return PG_OK;
}

static pg_status pq_read_run(PqReader *rd, uint64_t *out_run) {
    const Schema *s = rd->schema;
    const int nc = s->n_cols();
    const int64_t n = rd->n_rows;
    auto run = std::make_unique<Run>();
    run->own_schema = *s;
    run->schema = &run->own_schema;
    run->n_rows = n;
    run->cols.resize(nc);
    run->varlen_bytes.assign(nc, 0);
    // every reader thread decodes on its own stream: the files of a section decode concurrently
    cudaStream_t sm = thread_stream();
    auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };

    // ---- device memory: file bytes + scratch (freed at the end) and the output columns (owned by the run)
    size_t scratch = pad(rd->file.size() + 64) + pad(sizeof(PqPageJob) * rd->jobs.size() + 64) +
                     pad(sizeof(PqDictJob) * rd->dicts.size() + 64) + pad(sizeof(PqCol) * nc) +
                     pad(sizeof(PqPageState) * rd->jobs.size() + 64) + pad(sizeof(void *) * (rd->dict_entries + 1)) +
                     pad(4 * (rd->dict_entries + 1)) + 4096 + pad((size_t)rd->unc_bytes + 64) +
                     pad(sizeof(SnappyJob) * rd->unc.size() + 64) + 256 + pad((size_t)rd->delta_bytes + 64) +
                     pad(sizeof(DeltaJob) * rd->delta.size() + 64) + 256;
    size_t outb = 4096;
    std::vector<PqCol> cols(nc);
    for (int c = 0; c < nc; c++) {
        pg_field f = s->field(c);
        const pq::SchemaElement &e = rd->meta.schema[c + 1];
        PqCol &pc = cols[c];
        memset(&pc, 0, sizeof(pc));
        pc.phys = e.type;
        pc.phys_width = e.type == pq::T_INT32 || e.type == pq::T_FLOAT ? 4 : (e.type == pq::T_BYTE_ARRAY ? 0 : 8);
        pc.out_width = out_width_of(f.type);
        pc.nullable = e.repetition == pq::R_OPTIONAL;
        if (pc.nullable) { scratch += pad((size_t)n + 64); outb += pad((size_t)((n + 31) / 32) * 4 + 64); }
        scratch += pad(4 * (size_t)n + 64);                                        // ids
        if (pc.phys == pq::T_BYTE_ARRAY) {
            scratch += 2 * pad(8 * (size_t)n + 64) + pad(4 * (size_t)n + 64);       // vptr, rowsrc, vlen
            scratch += pad(8 * ((size_t)n / 4096 + 2));                            // scan block sums
            outb += pad(4 * (size_t)(n + 1) + 64);
        } else {
            outb += pad((size_t)n * pc.out_width + 64);
        }
    }
    unsigned char *d_scratch = nullptr, *d_out = nullptr;
    // recycled buffers (no cudaMalloc / cudaFree, which would serialise concurrent decodes)
    size_t scratch_got = 0, out_got = 0;
    d_scratch = (unsigned char *)device_buffer_take(scratch, &scratch_got);
    if (!d_scratch) return fail(PG_ERR_CUDA, "parquet: out of device memory");
    struct ScratchGuard { unsigned char *p; size_t n; ~ScratchGuard() { device_buffer_give(p, n); } } guard{d_scratch, scratch_got};
    d_out = (unsigned char *)device_buffer_take(outb, &out_got);
    if (!d_out) return fail(PG_ERR_CUDA, "parquet: out of device memory");
    run->owned.push_back(d_out);
    run->owned_bytes.push_back(out_got);
    size_t st = 0, ot = 0;
    auto stake = [&](size_t b) { unsigned char *p = d_scratch + st; st += pad(b); return p; };
    auto otake = [&](size_t b) { unsigned char *p = d_out + ot; ot += pad(b); return p; };

    uint8_t *d_file = stake(rd->file.size() + 64);
    PG_CUDA(cudaMemcpyAsync(d_file, rd->file.data(), rd->file.size(), cudaMemcpyHostToDevice, sm));
    run->bytes_h2d = (int64_t)rd->file.size();
    std::vector<PqPageJob> jobs = rd->jobs;
    std::vector<PqDictJob> dicts = rd->dicts;
    for (auto &j : jobs) j.body = d_file + (uintptr_t)j.body;
    for (auto &d : dicts) d.body = d_file + (uintptr_t)d.body;
    cudaEvent_t e0, e1;
    PG_CUDA(cudaEventCreate(&e0));
    PG_CUDA(cudaEventCreate(&e1));
    PG_CUDA(cudaEventRecord(e0, sm));
    int32_t *d_err_early = nullptr;
    if (!rd->unc.empty()) {
        // Snappy pages: one warp per page decompresses into a scratch image the decode kernels then read
        uint8_t *d_unc = stake((size_t)rd->unc_bytes + 64);
        SnappyJob *d_sj = (SnappyJob *)stake(sizeof(SnappyJob) * rd->unc.size());
        d_err_early = (int32_t *)stake(16);
        std::vector<SnappyJob> sj(rd->unc.size());
        for (size_t i = 0; i < sj.size(); i++) {
            const PqReader::Unc &u = rd->unc[i];
            sj[i] = SnappyJob{d_file + u.src_off, d_unc + u.dst_off, u.src_len, u.dst_len, u.prefix, 0};
        }
        PG_CUDA(cudaMemsetAsync(d_err_early, 0, 4, sm));
        PG_CUDA(cudaMemcpyAsync(d_sj, sj.data(), sizeof(SnappyJob) * sj.size(), cudaMemcpyHostToDevice, sm));
        PG_CUDA(cudaStreamSynchronize(sm));           // sj is a local vector
        const int64_t threads = (int64_t)sj.size() * 32;
        k_pq_snappy<<<(int)((threads + 127) / 128), 128, 0, sm>>>(d_sj, (int)sj.size(), d_err_early);
        for (size_t i = 0; i < jobs.size(); i++)
            if (rd->job_unc[i] >= 0) jobs[i].body = d_unc + rd->unc[rd->job_unc[i]].dst_off;
        for (size_t i = 0; i < dicts.size(); i++)
            if (rd->dict_unc[i] >= 0) dicts[i].body = d_unc + rd->unc[rd->dict_unc[i]].dst_off;
    }
    if (!rd->delta.empty()) {
        uint8_t *d_delta = stake((size_t)rd->delta_bytes + 64);
        DeltaJob *d_dj = (DeltaJob *)stake(sizeof(DeltaJob) * rd->delta.size());
        if (!d_err_early) {
            d_err_early = (int32_t *)stake(16);
            PG_CUDA(cudaMemsetAsync(d_err_early, 0, 4, sm));
        }
        std::vector<DeltaJob> dj(rd->delta.size());
        for (size_t i = 0; i < dj.size(); i++) {
            const PqReader::Delta &d = rd->delta[i];
            dj[i] = DeltaJob{d_file + d.src_off, d_delta + d.dst_off, d.src_len, d.prefix, d.width, d.max_values};
        }
        PG_CUDA(cudaMemcpyAsync(d_dj, dj.data(), sizeof(DeltaJob) * dj.size(), cudaMemcpyHostToDevice, sm));
        PG_CUDA(cudaStreamSynchronize(sm));           // dj is a local vector
        const int64_t threads = (int64_t)dj.size() * 32;
        k_pq_delta<<<(int)((threads + 127) / 128), 128, 0, sm>>>(d_dj, (int)dj.size(), d_err_early);
        for (size_t i = 0; i < jobs.size(); i++)
            if (rd->job_delta[i] >= 0) jobs[i].body = d_delta + rd->delta[rd->job_delta[i]].dst_off;
    }
    PqPageJob *d_jobs = (PqPageJob *)stake(sizeof(PqPageJob) * jobs.size() + 64);
    PqDictJob *d_dicts = (PqDictJob *)stake(sizeof(PqDictJob) * dicts.size() + 64);
    PqCol *d_cols = (PqCol *)stake(sizeof(PqCol) * nc);
    PqPageState *d_state = (PqPageState *)stake(sizeof(PqPageState) * jobs.size() + 64);
    const uint8_t **d_dict_ptr = (const uint8_t **)stake(sizeof(void *) * (rd->dict_entries + 1));
    int32_t *d_dict_len = (int32_t *)stake(4 * (rd->dict_entries + 1));
    int32_t *d_err = (int32_t *)stake(64);
    std::vector<int64_t *> block_sums(nc, nullptr);
    for (int c = 0; c < nc; c++) {
        PqCol &pc = cols[c];
        if (pc.nullable) {
            pc.defs = stake((size_t)n + 64);
            size_t vb = (size_t)((n + 31) / 32) * 4 + 64;
            pc.out_validity = (uint32_t *)otake(vb);
            PG_CUDA(cudaMemsetAsync(pc.out_validity, 0, vb, sm));
        }
        pc.ids = (int32_t *)stake(4 * (size_t)n + 64);
        if (pc.phys == pq::T_BYTE_ARRAY) {
            pc.vptr = (const uint8_t **)stake(8 * (size_t)n + 64);
            pc.rowsrc = (const uint8_t **)stake(8 * (size_t)n + 64);
            pc.vlen = (int32_t *)stake(4 * (size_t)n + 64);
            block_sums[c] = (int64_t *)stake(8 * ((size_t)n / 4096 + 2));
            pc.out_offsets = (int32_t *)otake(4 * (size_t)(n + 1) + 64);
            PG_CUDA(cudaMemsetAsync(pc.out_offsets, 0, 4, sm));
        } else {
            pc.out_data = otake((size_t)n * pc.out_width + 64);
        }
    }
    PG_CUDA(cudaMemsetAsync(d_err, 0, 4, sm));
    PG_CUDA(cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(PqPageJob) * jobs.size(), cudaMemcpyHostToDevice, sm));
    if (!dicts.empty())
        PG_CUDA(cudaMemcpyAsync(d_dicts, dicts.data(), sizeof(PqDictJob) * dicts.size(), cudaMemcpyHostToDevice, sm));
    PG_CUDA(cudaMemcpyAsync(d_cols, cols.data(), sizeof(PqCol) * nc, cudaMemcpyHostToDevice, sm));

    const int nj = (int)jobs.size(), nd = (int)dicts.size();
    int launches = 0;
    if (nj > 0) {
        k_pq_hybrid<<<(nj * 32 + 127) / 128, 128, 0, sm>>>(d_jobs, nj, d_cols, d_state);
        k_pq_walk_bytes<<<(nj + nd + kWalkWarps - 1) / kWalkWarps, kWalkWarps * 32, 0, sm>>>(
            d_jobs, nj, d_dicts, nd, d_cols, d_state, d_dict_ptr, d_dict_len);
        k_pq_assemble<<<nj, 256, 0, sm>>>(d_jobs, d_dicts, d_cols, d_state, d_dict_ptr, d_dict_len);
        launches += 3;
    }
    // var-len columns: lengths -> offsets, then the payload (its size needs one read-back per column)
    std::vector<unsigned char *> payload(nc, nullptr);
    for (int c = 0; c < nc && n > 0; c++) {
        if (cols[c].phys != pq::T_BYTE_ARRAY) continue;
        int64_t nb = (n + 4095) / 4096;
        k_scan_block_sums<<<(int)nb, 256, 0, sm>>>(cols[c].out_offsets + 1, n, block_sums[c]);
        k_scan_block_prefix<<<1, 32, 0, sm>>>(block_sums[c], nb, d_err);
        k_scan_apply<<<(int)nb, 256, 0, sm>>>(cols[c].out_offsets + 1, n, block_sums[c]);
        launches += 3;
    }
    {
        // one read-back for all var-len columns (their exact payload sizes), one allocation, then the copies
        std::vector<int32_t> totals(nc, 0);
        for (int c = 0; c < nc && n > 0; c++)
            if (cols[c].phys == pq::T_BYTE_ARRAY)
                PG_CUDA(cudaMemcpyAsync(&totals[c], cols[c].out_offsets + n, 4, cudaMemcpyDeviceToHost, sm));
        PG_CUDA(cudaStreamSynchronize(sm));
        size_t sum = 256;
        for (int c = 0; c < nc; c++) if (cols[c].phys == pq::T_BYTE_ARRAY) sum += pad((size_t)totals[c] + 64);
        unsigned char *pl = nullptr;
        if (n > 0 && sum > 256) {
            size_t pl_got = 0;
            pl = (unsigned char *)device_buffer_take(sum, &pl_got);
            if (!pl) return fail(PG_ERR_CUDA, "parquet: out of device memory");
            run->owned_bytes.push_back(pl_got);
            run->owned.push_back(pl);
        }
        size_t pt = 0;
        for (int c = 0; c < nc && n > 0; c++) {
            if (cols[c].phys != pq::T_BYTE_ARRAY) continue;
            payload[c] = pl + pt;
            pt += pad((size_t)totals[c] + 64);
            run->varlen_bytes[c] = totals[c];
            int64_t threads = n * 8;
            k_pq_copy_bytes<<<(int)((threads + 255) / 256), 256, 0, sm>>>(cols[c].rowsrc, cols[c].out_offsets,
                                                                        payload[c], n);
            launches++;
        }
    }
    PG_CUDA(cudaEventRecord(e1, sm));
    int32_t herr = 0;
    PG_CUDA(cudaMemcpyAsync(&herr, d_err, 4, cudaMemcpyDeviceToHost, sm));
    int32_t herr_pages = 0;
    if (d_err_early) PG_CUDA(cudaMemcpyAsync(&herr_pages, d_err_early, 4, cudaMemcpyDeviceToHost, sm));
    PG_CUDA(cudaStreamSynchronize(sm));
    PG_CUDA(cudaGetLastError());
    cudaEventElapsedTime(&rd->ms_decode, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    rd->launches = launches;
    if (herr_pages != KERR_NONE) {
        for (size_t q = 0; q < run->owned.size(); q++) device_buffer_give(run->owned[q], run->owned_bytes[q]);
        return fail(PG_ERR_FORMAT, "parquet: a Snappy / DELTA_BINARY_PACKED page does not expand to its declared size");
    }
    if (herr != KERR_NONE) {
        for (size_t q = 0; q < run->owned.size(); q++) device_buffer_give(run->owned[q], run->owned_bytes[q]);
        return fail(PG_ERR_INTERNAL, "parquet: a var-len column exceeds 2 GiB of payload");
    }
    for (int c = 0; c < nc; c++) {
        DevColumn dc;
        dc.data = cols[c].phys == pq::T_BYTE_ARRAY ? (const void *)payload[c] : cols[c].out_data;
        if (cols[c].phys == pq::T_BYTE_ARRAY && n == 0) dc.data = d_out;
        dc.offsets = cols[c].out_offsets;
        dc.validity = (const uint8_t *)cols[c].out_validity;
        run->cols[c] = dc;
    }
    *out_run = register_run(std::move(run));
    return PG_OK;
}

// ------------------------------------------------------------------ deletion vectors
//
// ApplyDeletionVectorReader (paimon-core/.../deletionvectors/ApplyDeletionVectorReader.java:31-54) skips the rows of
// a data file whose position is marked in the file's deletion vector.  Here: a device run minus the marked rows.

__global__ void k_dv_keep(const uint8_t *deleted, int64_t n_bits, int64_t n, int32_t *keep) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool del = i < n_bits && ((deleted[i >> 3] >> (i & 7)) & 1);
    keep[i] = del ? 0 : 1;
}
// src[j] = input row of output row j (incl = inclusive scan of keep)
__global__ void k_dv_sources(const int32_t *incl, int64_t n, int32_t *src) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t cur = incl[i], prev = i ? incl[i - 1] : 0;
    if (cur != prev) src[cur - 1] = (int32_t)i;
}
__global__ void k_dv_gather_fixed(const void *in, int width, const int32_t *src, int64_t m, void *out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    store_fixed(out, width, j, load_fixed(in, width, src[j]));
}
__global__ void k_dv_gather_bits(const uint8_t *in, const int32_t *src, int64_t m, uint32_t *out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool v = j < m && valid_bit(in, src[j]);
    const unsigned w = __ballot_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && j < m) out[j >> 5] = w;
}
__global__ void k_dv_lengths(const int32_t *offs, const int32_t *src, int64_t m, int32_t *out_offsets) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) out_offsets[0] = 0;
    if (j >= m) return;
    out_offsets[1 + j] = offs[src[j] + 1] - offs[src[j]];
}
__global__ void k_dv_copy_bytes(const uint8_t *data, const int32_t *offs, const int32_t *src, const int32_t *out_offsets,
                                uint8_t *out, int64_t m) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (row >= m) return;
    const uint8_t *s = data + offs[src[row]];
    const int o0 = out_offsets[row], o1 = out_offsets[row + 1];
    for (int b = o0 + (threadIdx.x & 7); b < o1; b += 8) out[b] = s[b - o0];
}

Run *run_from_handle(uint64_t h);                          // api.cu

static pg_status apply_deletion_vector(uint64_t run_h, const uint8_t *deleted, int64_t n_bits, uint64_t *out_run) {
    Run *in = run_from_handle(run_h);
    if (!in || !out_run || (n_bits > 0 && !deleted)) return fail(PG_ERR_INVALID, "unknown run handle or null argument");
    pg_status st = require_device();
    if (st) return st;
    const Schema *s = in->schema;
    const int nc = s->n_cols();
    const int64_t n = in->n_rows;
    if (n_bits < 0) return fail(PG_ERR_INVALID, "negative deletion vector size");
    auto run = std::make_unique<Run>();
    run->own_schema = *s;
    run->schema = &run->own_schema;
    run->cols.resize(nc);
    run->varlen_bytes.assign(nc, 0);
    run->varlen_base.assign(nc, 0);
    auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
    cudaStream_t sm = 0;
    // ---- kept rows
    uint8_t *d_del = nullptr;
    int32_t *d_incl = nullptr, *d_src = nullptr, *d_err = nullptr;
    int64_t *d_sums = nullptr;
    const int64_t nb = std::max<int64_t>((n + 4095) / 4096, 1);
    std::vector<void *> temps;
    auto tmp = [&](size_t bytes, void **p) -> cudaError_t {
        cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
        if (e == cudaSuccess) temps.push_back(*p);
        return e;
    };
    auto free_temps = [&]() { for (void *p : temps) cudaFree(p); };
#define DV_CUDA(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { free_temps(); for (void *q : run->owned) cudaFree(q); \
        return fail(PG_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(_e)); } } while (0)
    DV_CUDA(tmp((size_t)(n_bits + 7) / 8 + 16, (void **)&d_del));
    DV_CUDA(tmp(sizeof(int32_t) * (size_t)(n + 1), (void **)&d_incl));
    DV_CUDA(tmp(sizeof(int64_t) * (size_t)nb, (void **)&d_sums));
    DV_CUDA(tmp(16, (void **)&d_err));
    DV_CUDA(cudaMemsetAsync(d_err, 0, 4, sm));
    if (n_bits > 0) DV_CUDA(cudaMemcpyAsync(d_del, deleted, (size_t)(n_bits + 7) / 8, cudaMemcpyHostToDevice, sm));
    int64_t m = 0;
    if (n > 0) {
        k_dv_keep<<<(int)((n + 255) / 256), 256, 0, sm>>>(d_del, n_bits, n, d_incl);
        k_scan_block_sums<<<(int)nb, 256, 0, sm>>>(d_incl, n, d_sums);
        k_scan_block_prefix<<<1, 32, 0, sm>>>(d_sums, nb, d_err);
        k_scan_apply<<<(int)nb, 256, 0, sm>>>(d_incl, n, d_sums);
        int32_t last = 0;
        DV_CUDA(cudaMemcpyAsync(&last, d_incl + n - 1, 4, cudaMemcpyDeviceToHost, sm));
        DV_CUDA(cudaStreamSynchronize(sm));
        m = last;
    }
    run->n_rows = m;
    DV_CUDA(tmp(sizeof(int32_t) * (size_t)std::max<int64_t>(m, 1), (void **)&d_src));
    if (n > 0) k_dv_sources<<<(int)((n + 255) / 256), 256, 0, sm>>>(d_incl, n, d_src);
    // ---- one allocation for fixed-width data, offsets and validity; payloads follow once their sizes are known
    std::vector<size_t> o_data(nc), o_off(nc), o_val(nc);
    size_t total = 0;
    for (int c = 0; c < nc; c++) {
        pg_field f = s->field(c);
        o_data[c] = total; total += is_varlen(f.type) ? 0 : pad((size_t)m * type_width(f.type) + 16);
        o_off[c] = total; total += is_varlen(f.type) ? pad(sizeof(int32_t) * (size_t)(m + 1) + 16) : 0;
        o_val[c] = total; total += in->cols[c].validity ? pad((size_t)((m + 31) / 32) * 4 + 16) : 0;
    }
    unsigned char *base = nullptr;
    DV_CUDA(cudaMalloc(&base, total + 256));
    run->owned.push_back(base);
    const int gm = (int)((std::max<int64_t>(m, 1) + 255) / 256);
    std::vector<int64_t *> sums(nc, nullptr);
    for (int c = 0; c < nc; c++) {
        pg_field f = s->field(c);
        const DevColumn &ic = in->cols[c];
        DevColumn oc;
        if (ic.validity) {
            oc.validity = base + o_val[c];
            if (m > 0) k_dv_gather_bits<<<gm, 256, 0, sm>>>(ic.validity, d_src, m, (uint32_t *)(base + o_val[c]));
        }
        if (!is_varlen(f.type)) {
            oc.data = base + o_data[c];
            if (m > 0) k_dv_gather_fixed<<<gm, 256, 0, sm>>>(ic.data, type_width(f.type), d_src, m, base + o_data[c]);
        } else {
            int32_t *oo = (int32_t *)(base + o_off[c]);
            oc.offsets = oo;
            k_dv_lengths<<<gm, 256, 0, sm>>>(ic.offsets, d_src, m, oo);
            if (m > 0) {
                const int64_t mb = (m + 4095) / 4096;
                DV_CUDA(tmp(sizeof(int64_t) * (size_t)mb, (void **)&sums[c]));
                k_scan_block_sums<<<(int)mb, 256, 0, sm>>>(oo + 1, m, sums[c]);
                k_scan_block_prefix<<<1, 32, 0, sm>>>(sums[c], mb, d_err);
                k_scan_apply<<<(int)mb, 256, 0, sm>>>(oo + 1, m, sums[c]);
            }
        }
        run->cols[c] = oc;
    }
    // payload sizes: one read-back for all var-len columns
    std::vector<int32_t> totals(nc, 0);
    for (int c = 0; c < nc; c++)
        if (is_varlen(s->field(c).type) && m > 0)
            DV_CUDA(cudaMemcpyAsync(&totals[c], run->cols[c].offsets + m, 4, cudaMemcpyDeviceToHost, sm));
    int32_t herr = 0;
    DV_CUDA(cudaMemcpyAsync(&herr, d_err, 4, cudaMemcpyDeviceToHost, sm));
    DV_CUDA(cudaStreamSynchronize(sm));
    size_t ptotal = 0;
    for (int c = 0; c < nc; c++) if (is_varlen(s->field(c).type)) ptotal += pad((size_t)totals[c] + 64);
    unsigned char *pl = nullptr;
    if (ptotal) { DV_CUDA(cudaMalloc(&pl, ptotal)); run->owned.push_back(pl); }
    size_t pt = 0;
    for (int c = 0; c < nc; c++) {
        if (!is_varlen(s->field(c).type)) continue;
        run->cols[c].data = pl ? pl + pt : base;
        run->varlen_bytes[c] = totals[c];
        if (m > 0)
            k_dv_copy_bytes<<<(int)((m * 8 + 255) / 256), 256, 0, sm>>>((const uint8_t *)in->cols[c].data, in->cols[c].offsets,
                                                                        d_src, run->cols[c].offsets, pl + pt, m);
        pt += pad((size_t)totals[c] + 64);
    }
    DV_CUDA(cudaStreamSynchronize(sm));
    DV_CUDA(cudaGetLastError());
    free_temps();
#undef DV_CUDA
    if (herr != KERR_NONE) {
        for (void *q : run->owned) cudaFree(q);
        return fail(PG_ERR_INTERNAL, "deletion vector: a var-len column exceeds 2 GiB of payload");
    }
    *out_run = register_run(std::move(run));
    return PG_OK;
}

}  // namespace pg

using namespace pg;

extern "C" {

pg_status pg_parquet_open(uint64_t schema, const uint8_t *file_bytes, int64_t size, uint64_t *out_reader) {
    return pq_open(schema, file_bytes, size, out_reader);
}

pg_status pg_parquet_describe(uint64_t reader, pg_parquet_info *out) {
    std::lock_guard<std::mutex> lk(g_pq_mu);
    auto it = g_pq.find(reader);
    if (it == g_pq.end() || !out) return fail(PG_ERR_INVALID, "unknown parquet reader handle");
    PqReader *rd = it->second.get();
    out->n_rows = rd->n_rows;
    out->n_row_groups = (int32_t)rd->meta.row_groups.size();
    out->n_columns = rd->schema->n_cols();
    out->n_data_pages = (int32_t)rd->jobs.size();
    out->n_dictionary_pages = (int32_t)rd->dicts.size();
    out->ms_decode = rd->ms_decode;
    out->launches = rd->launches;
    return PG_OK;
}

pg_status pg_parquet_read_run(uint64_t reader, uint64_t *out_run) {
    PqReader *rd;
    {
        std::lock_guard<std::mutex> lk(g_pq_mu);
        auto it = g_pq.find(reader);
        if (it == g_pq.end() || !out_run) return fail(PG_ERR_INVALID, "unknown parquet reader handle");
        rd = it->second.get();
    }
    pg_status st = require_device();          // fails loudly without pg_init / a CUDA device: no CPU fallback
    if (st) return st;
    return pq_read_run(rd, out_run);
}

pg_status pg_run_apply_deletion_vector(uint64_t run, const uint8_t *deleted_bitmap, int64_t n_bits, uint64_t *out_run) {
    return apply_deletion_vector(run, deleted_bitmap, n_bits, out_run);
}

pg_status pg_parquet_free(uint64_t reader) {
    std::lock_guard<std::mutex> lk(g_pq_mu);
    return g_pq.erase(reader) ? PG_OK : fail(PG_ERR_INVALID, "unknown parquet reader handle");
}

}  // extern "C"
