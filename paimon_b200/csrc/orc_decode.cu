// orc_decode.cu — ORC stripe decode on the device behind the same format seam as the Parquet decoder
// (SURVEY.md §8 rows a25 / f3).
//
// Reference being replaced: paimon-format/src/main/java/org/apache/paimon/format/orc/OrcReaderFactory.java:98-163
// (createReader: orc-core RecordReader over the projected TypeDescription, batches of VectorizedRowBatch) and the
// vector adapters in orc/reader/*; the decode arithmetic itself is orc-core 1.9.2 (not under /root/reference; restated
// from the public ORC specification in orc_meta.cc / orc_device.cuh).
//
// One call decodes a whole SECTION like pg_parquet_read_section: the files of a sorted run are concatenated into one
// device run.  Host: file tails (protobuf footers, inflated on the host) -> a plan of streams and (stripe, column)
// tasks.  Device: k_orc_inflate (one warp per stream: compression chunks -> contiguous bytes with the shared
// DEFLATE / zstd decoders), k_orc_task<0> (one thread per task: PRESENT -> validity, values / lengths), an offsets
// scan per var-len column, k_orc_task<1> (payload bytes).  The stream decoders are the host-pinned orc_device.cuh.
#include <algorithm>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <unordered_map>

#include "inflate_device.cuh"
#include "orc_device.cuh"
#include "orc_meta.h"
#include "scan_kernels.cuh"
#include "zstd_device.cuh"

namespace pg {

Schema *schema_from_handle(uint64_t h);                 // api.cu
void *device_buffer_take(size_t bytes, size_t *got);
void device_buffer_give(void *p, size_t bytes);
cudaStream_t thread_stream();
uint64_t register_run(std::unique_ptr<Run> run);
pg_status require_device();

struct OrcStream {
    const uint8_t *src;        // device: the stream as stored in the file
    uint8_t *dst;              // scratch image (compressed files)
    int64_t length, bound;
    int32_t codec;             // orc::Compression of the stream's file (files of a section may differ)
    int32_t block_size;        // compression block size of that file
    const uint8_t *bytes;      // result: contiguous decoded bytes
    int64_t n;
};

struct OrcTaskRef {            // stream table indexes of a task (-1 = absent)
    int32_t s_present, s_data, s_length, s_dict, s_secondary;
};

constexpr int kOrcWarps = 4;
__global__ void __launch_bounds__(kOrcWarps * 32)
k_orc_inflate(OrcStream *streams, int n_streams, uint8_t *lit_scratch, int32_t *counter, int32_t *err) {
    __shared__ zs::Tables ZT[kOrcWarps];               // (the DEFLATE tables are smaller and overlay them)
    static_assert(sizeof(inflate::Tables) <= sizeof(zs::Tables), "tables overlay");
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *lit = lit_scratch + ((size_t)blockIdx.x * kOrcWarps + w) * (size_t)(zs::kMaxBlock + 64);
    while (true) {
        int j = 0;
        if (lane == 0) j = atomicAdd(counter, 1);
        j = __shfl_sync(0xffffffffu, j, 0);
        if (j >= n_streams) return;
        OrcStream st = streams[j];
        const int codec = st.codec;
        const int64_t block_size = st.block_size;
        if (codec == orc::C_NONE) {
            if (lane == 0) { streams[j].bytes = st.src; streams[j].n = st.length; }
            continue;
        }
        int64_t pos = 0, out = 0;
        bool bad = false;
        while (pos < st.length && !bad) {
            if (st.length - pos < 3) { bad = true; break; }
            const uint32_t h = st.src[pos] | (st.src[pos + 1] << 8) | (st.src[pos + 2] << 16);
            const int64_t len = h >> 1;
            pos += 3;
            if (st.length - pos < len) { bad = true; break; }
            if (h & 1) {
                if (out + len > st.bound) { bad = true; break; }
                for (int64_t i = lane; i < len; i += 32) st.dst[out + i] = st.src[pos + i];
                __syncwarp();
                out += len;
            } else {
                const int64_t cap = min(block_size, st.bound - out);
                int64_t got;
                if (codec == orc::C_ZLIB) got = inflate::inflate_raw(st.src + pos, len, st.dst + out, cap, *(inflate::Tables *)&ZT[w], nullptr);
                else got = zs::decode(st.src + pos, len, st.dst + out, cap, lit, ZT[w]);
                __syncwarp();
                if (got < 0) { bad = true; break; }
                out += got;
            }
            pos += len;
        }
        if (lane == 0) {
            if (bad) atomicCAS(err, KERR_NONE, KERR_BAD_PAGE);
            streams[j].bytes = st.dst;
            streams[j].n = bad ? 0 : out;
        }
    }
}

// one thread per (stripe, column); PHASE 0 = validity / values / lengths, PHASE 1 = var-len payload
template <int PHASE>
__global__ void k_orc_task(orcdev::Task *tasks, const OrcTaskRef *refs, const OrcStream *streams, int n_tasks, int32_t *err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tasks) return;
    orcdev::Task t = tasks[i];
    if (PHASE == 0) {
        const OrcTaskRef r = refs[i];
        auto bind = [&](int idx, const uint8_t *&p, int64_t &n) {
            if (idx >= 0) { p = streams[idx].bytes; n = streams[idx].n; } else { p = nullptr; n = 0; }
        };
        bind(r.s_present, t.present, t.present_n);
        bind(r.s_data, t.data, t.data_n);
        bind(r.s_length, t.length, t.length_n);
        bind(r.s_dict, t.dict_data, t.dict_data_n);
        bind(r.s_secondary, t.secondary, t.secondary_n);
        orcdev::decode_task_a(t);
    } else {
        orcdev::decode_task_b(t);
    }
    if (t.bad) atomicCAS(err, KERR_NONE, KERR_BAD_PAGE);
    tasks[i] = t;
}

// var-len output column: the payload pointer reaches the tasks after the size read-back
__global__ void k_orc_set_payload(orcdev::Task *tasks, int n_tasks, const int32_t *task_out, uint8_t *const *payload_of_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tasks && tasks[i].out_width == 0) tasks[i].out_payload = payload_of_out[task_out[i]];
}

static int orc_out_width(int t) {
    switch (t) {
        case PG_INT8: case PG_BOOL: return 1;
        case PG_INT16: return 2;
        case PG_INT32: case PG_FLOAT: return 4;
        case PG_INT64: case PG_DOUBLE: return 8;
        default: return 0;
    }
}

// OrcTypeUtil (paimon-format/.../orc/OrcTypeUtil.java convertToOrcType): which ORC type a Paimon column has in the file;
// the integer / float widenings schema evolution allows are accepted (orc-core's SchemaEvolution does the same)
static bool orc_type_ok(int pg_t, const orc::Type &ty) {
    const int k = ty.kind;
    switch (pg_t) {
        case PG_BOOL: return k == orc::K_BOOLEAN;
        case PG_INT8: return k == orc::K_BYTE;
        case PG_INT16: return k == orc::K_SHORT || k == orc::K_BYTE;
        case PG_INT32: return k == orc::K_INT || k == orc::K_SHORT || k == orc::K_BYTE || k == orc::K_DATE;
        case PG_INT64: return k == orc::K_LONG || k == orc::K_INT || k == orc::K_SHORT || k == orc::K_BYTE ||
                              (k == orc::K_DECIMAL && ty.precision <= 18);
        case PG_FLOAT: return k == orc::K_FLOAT;
        case PG_DOUBLE: return k == orc::K_DOUBLE || k == orc::K_FLOAT;
        case PG_STRING: return k == orc::K_STRING || k == orc::K_VARCHAR || k == orc::K_CHAR;
        case PG_BINARY: return k == orc::K_BINARY;
        default: return false;
    }
}

struct OrcBufs {
    cudaStream_t stream = nullptr;
    std::vector<std::pair<void *, size_t>> bufs;
    void *take(size_t bytes) {
        size_t got = 0;
        void *p = device_buffer_take(bytes ? bytes : 256, &got);
        if (p) bufs.push_back({p, got});
        return p;
    }
    ~OrcBufs() {
        if (stream && !bufs.empty()) cudaStreamSynchronize(stream);
        for (auto &b : bufs) device_buffer_give(b.first, b.second);
    }
};

static pg_status orc_decode_section(const Schema *s, const pg_file_desc *files, int nf, int n_runs, const char *const *names,
                                    const uint8_t *read_cols, uint64_t *out_runs, pg_section_info *info) {
    const int nc = s->n_cols();
    cudaStream_t sm = thread_stream();
    auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
    OrcBufs scratch;
    scratch.stream = sm;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    struct EvGuard { cudaEvent_t &a, &b; ~EvGuard() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); } } evg{e0, e1};
    PG_CUDA(cudaEventCreate(&e0));
    PG_CUDA(cudaEventCreate(&e1));
    PG_CUDA(cudaEventRecord(e0, sm));
    if (read_cols)
        for (int c = 0; c < s->n_key + 2; c++)
            if (!read_cols[c]) return fail(PG_ERR_INVALID, "orc: key, sequence number and kind columns are always read");
    std::vector<uint8_t> wanted(nc, 1);
    for (int c = 0; c < nc; c++) wanted[c] = !read_cols || read_cols[c];

    // ---- file tails, schema mapping by name, plans
    std::vector<orc::FileTail> tails(nf);
    std::vector<orc::Plan> plans(nf);
    std::vector<const uint8_t *> d_file(nf, nullptr);
    std::vector<int64_t> run_rows(n_runs, 0), file_row0(nf, 0);
    int64_t file_bytes = 0, page_bytes = 0;
    bool any_compressed = false, any_zstd = false;
    std::vector<uint8_t> col_missing((size_t)n_runs * nc, 0);
    std::vector<int> files_of_run(n_runs, 0);
    for (int f = 0; f < nf; f++) {
        if (files[f].mem != PG_MEM_HOST)
            return fail(PG_ERR_UNSUPPORTED, "orc: the file bytes must be host memory (the footers and chunk headers are walked on the host)");
        if (files[f].run < 0 || files[f].run >= n_runs) return fail(PG_ERR_INVALID, "orc section: run index out of range");
        std::vector<int> file_col(nc, -1);
        try {
            tails[f] = orc::parse_file(files[f].bytes, files[f].size);
            const orc::FileTail &t = tails[f];
            if (t.types.empty() || t.types[0].kind != orc::K_STRUCT) return fail(PG_ERR_UNSUPPORTED, "orc: the root type is not a struct");
            if (t.compression != orc::C_NONE && t.compression != orc::C_ZLIB && t.compression != orc::C_ZSTD)
                return fail(PG_ERR_UNSUPPORTED, "orc: compression kind " + std::to_string(t.compression) +
                                                " is not decoded on device (NONE, ZLIB and ZSTD are)");
            if (t.compression != orc::C_NONE) any_compressed = true;
            if (t.compression == orc::C_ZSTD) any_zstd = true;
            if (t.block_size > (1u << 30)) return fail(PG_ERR_UNSUPPORTED, "orc: compression block size above 1 GiB");
            const orc::Type &root = t.types[0];
            std::unordered_map<std::string, int> by_name;
            for (size_t i = 0; i < root.field_names.size() && i < root.subtypes.size(); i++) by_name.emplace(root.field_names[i], (int)i);
            for (int c = 0; c < nc; c++) {
                if (!wanted[c]) continue;
                int fc = -1;
                if (names) {
                    auto it = by_name.find(names[c] ? names[c] : "");
                    if (it != by_name.end()) fc = it->second;
                } else if ((size_t)c < root.subtypes.size()) fc = c;
                if (fc < 0) {
                    if (c < s->n_key + 2 || !s->field(c).nullable)
                        return fail(PG_ERR_UNSUPPORTED, std::string("orc: the file has no column '") + (names ? names[c] : "?") +
                                                        "' and the read schema does not allow NULL for it");
                    col_missing[(size_t)files[f].run * nc + c] |= 1;
                    continue;
                }
                const uint32_t tid = root.subtypes[fc];
                if (tid >= t.types.size() || !orc_type_ok(s->field(c).type, t.types[tid]))
                    return fail(PG_ERR_UNSUPPORTED, "orc: column " + std::to_string(c) + " has an ORC type the device decoder does "
                                                    "not map to the table type (timestamps, DECIMAL(p > 18), nested types: Java side)");
                file_col[c] = fc;
            }
            if (!names && (int)root.subtypes.size() != nc)
                return fail(PG_ERR_UNSUPPORTED, "orc: the file's column count differs from the read schema (pass the field names)");
            plans[f] = orc::plan_file(t, files[f].bytes, files[f].size, file_col);
        } catch (const std::exception &e) {
            // (a codec or type this decoder does not cover is a refusal, not a malformed file)
            const bool refusal = strstr(e.what(), "is not decoded") != nullptr || strstr(e.what(), "not supported") != nullptr;
            return fail(refusal ? PG_ERR_UNSUPPORTED : PG_ERR_FORMAT, e.what());
        }
        file_row0[f] = run_rows[files[f].run];
        run_rows[files[f].run] += (int64_t)tails[f].rows;
        files_of_run[files[f].run]++;
        file_bytes += files[f].size;
        uint8_t *d = (uint8_t *)scratch.take((size_t)files[f].size + 64);
        if (!d) return fail(PG_ERR_CUDA, "orc: out of device memory");
        PG_CUDA(cudaMemcpyAsync(d, files[f].bytes, (size_t)files[f].size, cudaMemcpyHostToDevice, sm));
        d_file[f] = d;
    }
    for (int r = 0; r < n_runs; r++)
        if (run_rows[r] > 0x7fffffffLL) return fail(PG_ERR_UNSUPPORTED, "orc: more than 2^31 rows in one run");
    // a var-len column that only some files of a run have would need offsets filled for the other files' rows
    {
        std::vector<int> present_files((size_t)n_runs * nc, 0);
        for (int f = 0; f < nf; f++)
            for (const orc::PlanTask &t : plans[f].tasks) if (t.stripe == 0) present_files[(size_t)files[f].run * nc + t.col]++;
        for (int f = 0; f < nf; f++) {
            if (!tails[f].stripes.empty()) continue;       // a file without stripes has no tasks: it lacks nothing
            for (int c = 0; c < nc; c++) if (wanted[c]) present_files[(size_t)files[f].run * nc + c]++;
        }
        for (int r = 0; r < n_runs; r++)
            for (int c = 0; c < nc; c++)
                if (wanted[c] && is_varlen(s->field(c).type) && col_missing[(size_t)r * nc + c] && present_files[(size_t)r * nc + c] > 0 &&
                    present_files[(size_t)r * nc + c] < files_of_run[r])
                    return fail(PG_ERR_UNSUPPORTED, "orc: a var-len column exists in some files of a sorted run only");
    }

    // ---- output columns (validity bitmaps first and contiguous: one memset)
    std::vector<std::unique_ptr<Run>> runs(n_runs);
    struct RunGuard {
        std::vector<std::unique_ptr<Run>> &runs;
        ~RunGuard() {
            for (auto &r : runs)
                if (r) for (size_t q = 0; q < r->owned.size(); q++) device_buffer_give(r->owned[q], r->owned_bytes[q]);
        }
    } run_guard{runs};
    struct OutCol { void *data = nullptr; int32_t *offsets = nullptr; uint32_t *validity = nullptr; };
    std::vector<OutCol> outs((size_t)n_runs * nc);
    int64_t decoded_bytes = 0;
    for (int r = 0; r < n_runs; r++) {
        const int64_t n = run_rows[r];
        auto run = std::make_unique<Run>();
        run->own_schema = *s;
        run->schema = &run->own_schema;
        run->n_rows = n;
        run->cols.resize(nc);
        run->varlen_bytes.assign(nc, 0);
        run->varlen_base.assign(nc, 0);
        const size_t vb = pad((size_t)((n + 31) / 32) * 4 + 64);
        size_t vbytes = 0, total = 0;
        // ORC columns are nullable by format: every column the read schema calls nullable gets a bitmap
        for (int c = 0; c < nc; c++) if (wanted[c] && s->field(c).nullable) vbytes += vb;
        total = vbytes;
        std::vector<size_t> o_main(nc);
        for (int c = 0; c < nc; c++) {
            const int ow = orc_out_width(s->field(c).type);
            o_main[c] = total;
            if (wanted[c]) total += ow ? pad((size_t)n * ow + 64) : pad(4 * (size_t)(n + 1) + 64);
        }
        size_t got = 0;
        unsigned char *base = (unsigned char *)device_buffer_take(total + 256, &got);
        if (!base) return fail(PG_ERR_CUDA, "orc: out of device memory");
        run->owned.push_back(base);
        run->owned_bytes.push_back(got);
        if (vbytes) PG_CUDA(cudaMemsetAsync(base, 0, vbytes, sm));
        size_t vt = 0;
        for (int c = 0; c < nc; c++) {
            if (!wanted[c]) continue;
            OutCol &o = outs[(size_t)r * nc + c];
            const int ow = orc_out_width(s->field(c).type);
            if (s->field(c).nullable) { o.validity = (uint32_t *)(base + vt); vt += vb; decoded_bytes += (n + 7) / 8; }
            if (ow) { o.data = base + o_main[c]; decoded_bytes += n * ow; }
            else { o.offsets = (int32_t *)(base + o_main[c]); decoded_bytes += 4 * (n + 1); }
            // var-len lengths are scanned in place: rows nobody writes must read 0; missing columns are all NULL
            if (!ow || col_missing[(size_t)r * nc + c])
                PG_CUDA(cudaMemsetAsync(base + o_main[c], 0, ow ? (size_t)n * ow : 4 * (size_t)(n + 1), sm));
        }
        runs[r] = std::move(run);
    }

    // ---- stream and task tables
    std::vector<OrcStream> h_streams;
    std::vector<orcdev::Task> h_tasks;
    std::vector<OrcTaskRef> h_refs;
    std::vector<int32_t> h_task_out;
    uint64_t sc_bytes = 0, dict_entries = 0;
    for (int f = 0; f < nf; f++) { sc_bytes += plans[f].scratch_bytes; dict_entries += plans[f].dict_entries; }
    uint8_t *d_sc = !any_compressed ? nullptr : (uint8_t *)scratch.take((size_t)sc_bytes + 256);
    if (any_compressed && !d_sc) return fail(PG_ERR_CUDA, "orc: out of device memory");
    int32_t *d_dict_off = (int32_t *)scratch.take(4 * (size_t)(dict_entries + 1) + 256);
    if (!d_dict_off) return fail(PG_ERR_CUDA, "orc: out of device memory");
    uint64_t sc_base = 0, dict_base = 0;
    for (int f = 0; f < nf; f++) {
        const int s0 = (int)h_streams.size();
        for (const orc::PlanStream &ps : plans[f].streams) {
            OrcStream st{};
            st.src = d_file[f] + ps.offset;
            st.length = (int64_t)ps.length;
            st.bound = (int64_t)ps.out_bound;
            st.codec = tails[f].compression;
            st.block_size = (int32_t)tails[f].block_size;
            st.dst = d_sc ? d_sc + sc_base + ps.out_off : nullptr;
            h_streams.push_back(st);
            page_bytes += (int64_t)ps.length;
        }
        auto idx = [&](int i) { return i < 0 ? -1 : s0 + i; };
        for (const orc::PlanTask &p : plans[f].tasks) {
            const int r = files[f].run;
            const OutCol &o = outs[(size_t)r * nc + p.col];
            orcdev::Task k;
            memset(&k, 0, sizeof(k));
            k.row0 = file_row0[f] + p.row0;
            k.rows = p.rows;
            k.kind = p.kind; k.enc = p.enc; k.dict_size = (int32_t)p.dict_size; k.scale = p.scale;
            k.out_width = orc_out_width(s->field(p.col).type);
            k.out_data = o.data;
            k.out_offsets = o.offsets;
            k.out_validity = o.validity;
            k.dict_off = d_dict_off + dict_base + p.dict_off_base;
            h_tasks.push_back(k);
            h_refs.push_back(OrcTaskRef{idx(p.s_present), idx(p.s_data), idx(p.s_length), idx(p.s_dict), idx(p.s_secondary)});
            h_task_out.push_back(r * nc + p.col);
        }
        sc_base += plans[f].scratch_bytes;
        dict_base += plans[f].dict_entries;
    }
    const int n_streams = (int)h_streams.size(), n_tasks = (int)h_tasks.size();
    const size_t tb_s = pad(sizeof(OrcStream) * (size_t)std::max(n_streams, 1)), tb_t = pad(sizeof(orcdev::Task) * (size_t)std::max(n_tasks, 1));
    const size_t tb_r = pad(sizeof(OrcTaskRef) * (size_t)std::max(n_tasks, 1)), tb_o = pad(4 * (size_t)std::max(n_tasks, 1));
    const size_t tb_p = pad(sizeof(void *) * outs.size());
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int inflate_ctas = !any_compressed ? std::max(1, std::min(sms, (n_streams + kOrcWarps - 1) / kOrcWarps))
                                                  : std::max(1, std::min(sms * 4, (n_streams + kOrcWarps - 1) / kOrcWarps));
    const size_t tb_lit = any_zstd ? pad((size_t)inflate_ctas * kOrcWarps * (size_t)(zs::kMaxBlock + 64)) : 256;
    unsigned char *tb = (unsigned char *)scratch.take(tb_s + tb_t + tb_r + tb_o + tb_p + tb_lit + 1024);
    if (!tb) return fail(PG_ERR_CUDA, "orc: out of device memory");
    OrcStream *d_streams = (OrcStream *)tb;
    orcdev::Task *d_tasks = (orcdev::Task *)(tb + tb_s);
    OrcTaskRef *d_refs = (OrcTaskRef *)(tb + tb_s + tb_t);
    int32_t *d_task_out = (int32_t *)(tb + tb_s + tb_t + tb_r);
    uint8_t **d_payload = (uint8_t **)(tb + tb_s + tb_t + tb_r + tb_o);
    uint8_t *d_lit = tb + tb_s + tb_t + tb_r + tb_o + tb_p;
    int32_t *d_err = (int32_t *)(d_lit + tb_lit);
    int32_t *d_counter = d_err + 4;
    PG_CUDA(cudaMemsetAsync(d_err, 0, 64, sm));
    int launches = 0;
    if (n_streams) PG_CUDA(cudaMemcpyAsync(d_streams, h_streams.data(), sizeof(OrcStream) * n_streams, cudaMemcpyHostToDevice, sm));
    if (n_tasks) {
        PG_CUDA(cudaMemcpyAsync(d_tasks, h_tasks.data(), sizeof(orcdev::Task) * n_tasks, cudaMemcpyHostToDevice, sm));
        PG_CUDA(cudaMemcpyAsync(d_refs, h_refs.data(), sizeof(OrcTaskRef) * n_tasks, cudaMemcpyHostToDevice, sm));
        PG_CUDA(cudaMemcpyAsync(d_task_out, h_task_out.data(), 4 * (size_t)n_tasks, cudaMemcpyHostToDevice, sm));
    }
    if (n_streams) {
        k_orc_inflate<<<inflate_ctas, kOrcWarps * 32, 0, sm>>>(d_streams, n_streams, d_lit, d_counter, d_err);
        launches++;
    }
    if (n_tasks) {
        k_orc_task<0><<<(n_tasks + 31) / 32, 32, 0, sm>>>(d_tasks, d_refs, d_streams, n_tasks, d_err);
        launches++;
    }
    // ---- var-len columns: lengths -> offsets, exact payload sizes (one read-back), payload
    std::vector<std::pair<int, int>> vl;                 // (run, col)
    for (int r = 0; r < n_runs; r++)
        for (int c = 0; c < nc; c++)
            if (wanted[c] && is_varlen(s->field(c).type)) vl.push_back({r, c});
    std::vector<int32_t> totals(vl.size() + 1, 0);
    int32_t herr = 0;
    if (!vl.empty()) {
        int64_t max_n = 0;
        for (int r = 0; r < n_runs; r++) max_n = std::max(max_n, run_rows[r]);
        int64_t *d_sums = (int64_t *)scratch.take(8 * (size_t)(max_n / 4096 + 4));
        if (!d_sums) return fail(PG_ERR_CUDA, "orc: out of device memory");
        for (size_t i = 0; i < vl.size(); i++) {
            const OutCol &o = outs[(size_t)vl[i].first * nc + vl[i].second];
            const int64_t n = run_rows[vl[i].first];
            launch_offsets_scan(o.offsets, n, d_sums, d_err, sm);
            launches += n > 0 ? 3 : 0;
            PG_CUDA(cudaMemcpyAsync(&totals[i], o.offsets + n, 4, cudaMemcpyDeviceToHost, sm));
        }
        PG_CUDA(cudaMemcpyAsync(&herr, d_err, 4, cudaMemcpyDeviceToHost, sm));
        PG_CUDA(cudaStreamSynchronize(sm));               // the read-back: exact payload sizes
        if (herr != KERR_NONE)
            return fail(herr == KERR_OFFSET_OVERFLOW ? PG_ERR_INTERNAL : PG_ERR_FORMAT,
                        herr == KERR_OFFSET_OVERFLOW ? "orc: a var-len column exceeds 2 GiB of payload"
                                                     : "orc: a stream does not decode (malformed file or unsupported encoding)");
        std::vector<uint8_t *> h_payload(outs.size(), nullptr);
        for (int r = 0; r < n_runs; r++) {
            size_t sum = 256;
            for (size_t i = 0; i < vl.size(); i++) if (vl[i].first == r) sum += pad((size_t)totals[i] + 64);
            size_t got = 0;
            unsigned char *pl = (unsigned char *)device_buffer_take(sum, &got);
            if (!pl) return fail(PG_ERR_CUDA, "orc: out of device memory");
            runs[r]->owned.push_back(pl);
            runs[r]->owned_bytes.push_back(got);
            size_t pt = 0;
            for (size_t i = 0; i < vl.size(); i++) {
                if (vl[i].first != r) continue;
                outs[(size_t)r * nc + vl[i].second].data = pl + pt;
                h_payload[(size_t)r * nc + vl[i].second] = pl + pt;
                runs[r]->varlen_bytes[vl[i].second] = totals[i];
                decoded_bytes += totals[i];
                pt += pad((size_t)totals[i] + 64);
            }
        }
        PG_CUDA(cudaMemcpyAsync(d_payload, h_payload.data(), sizeof(void *) * outs.size(), cudaMemcpyHostToDevice, sm));
        if (n_tasks) {
            k_orc_set_payload<<<(n_tasks + 127) / 128, 128, 0, sm>>>(d_tasks, n_tasks, d_task_out, d_payload);
            k_orc_task<1><<<(n_tasks + 31) / 32, 32, 0, sm>>>(d_tasks, d_refs, d_streams, n_tasks, d_err);
            launches += 2;
        }
        PG_CUDA(cudaStreamSynchronize(sm));               // (h_payload is a local)
    }
    PG_CUDA(cudaEventRecord(e1, sm));
    PG_CUDA(cudaMemcpyAsync(&herr, d_err, 4, cudaMemcpyDeviceToHost, sm));
    PG_CUDA(cudaStreamSynchronize(sm));
    PG_CUDA(cudaGetLastError());
    if (herr != KERR_NONE) return fail(PG_ERR_FORMAT, "orc: a stream does not decode (malformed file or unsupported encoding)");
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int64_t n_rows = 0;
    for (int r = 0; r < n_runs; r++) {
        for (int c = 0; c < nc; c++) {
            const OutCol &o = outs[(size_t)r * nc + c];
            DevColumn dc;
            if (wanted[c]) {
                dc.data = o.data ? o.data : (const void *)runs[r]->owned[0];
                dc.offsets = o.offsets;
                dc.validity = (const uint8_t *)o.validity;
            }
            runs[r]->cols[c] = dc;
        }
        runs[r]->bytes_h2d = r == 0 ? file_bytes : 0;
        n_rows += run_rows[r];
        out_runs[r] = register_run(std::move(runs[r]));
    }
    if (info) {
        memset(info, 0, sizeof(*info));
        info->n_rows = n_rows;
        info->file_bytes = file_bytes;
        info->page_bytes = page_bytes;
        info->decoded_bytes = decoded_bytes;
        info->n_files = nf;
        info->n_runs = n_runs;
        info->n_chunks = n_tasks;
        info->n_data_pages = n_streams;
        info->launches = launches;
        info->ms_decode = ms;
    }
    return PG_OK;
}

}  // namespace pg

using namespace pg;

extern "C" pg_status pg_orc_read_section(uint64_t schema, const pg_file_desc *files, int32_t n_files, int32_t n_runs,
                                         const char *const *column_names, const uint8_t *read_columns, uint64_t *out_runs,
                                         pg_section_info *info) {
    Schema *s = schema_from_handle(schema);
    if (!s || !out_runs || n_files < 0 || n_runs < 0 || (n_files > 0 && !files))
        return fail(PG_ERR_INVALID, "bad schema handle or null argument");
    if (n_runs == 0) return n_files == 0 ? PG_OK : fail(PG_ERR_INVALID, "files without runs");
    pg_status st = require_device();
    if (st) return st;
    const Schema own = *s;
    return orc_decode_section(&own, files, n_files, n_runs, column_names, read_columns, out_runs, info);
}
