// upload.cu — start moving the data files of the NEXT section to the device while the current one merges.
//
// The Java read path fetches a bucket's data files through FileIO and hands them to the format readers
// (KeyValueFileReaderFactory.java:104-140); a compaction / scan task walks many sections and buckets one after the
// other (MergeTreeCompactRewriter.java:77-106 per section).  On the device path the host -> device copy of a section's
// file bytes (the encoded pages: ~0.3 of the decoded bytes) is the longest leg of a step, so it runs on its own copy
// stream: pg_files_upload_begin returns at once, pg_files_upload_wait blocks until the bytes are resident and hands
// back device descriptors for pg_parquet_read_section.  Host buffers should be page-locked (a pageable source makes
// the copy synchronous and staged by the driver).
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "pg_internal.h"

namespace pg {

void *device_buffer_take(size_t bytes, size_t *got);    // api.cu: recycled device buffers
void device_buffer_give(void *p, size_t bytes);
cudaStream_t thread_stream();
pg_status require_device();

namespace {

struct Upload {
    std::vector<void *> bufs;
    std::vector<size_t> got;
    std::vector<pg_file_desc> files;
    cudaEvent_t done = nullptr;
    ~Upload() {
        if (done) cudaEventDestroy(done);
        for (size_t i = 0; i < bufs.size(); i++) if (bufs[i]) device_buffer_give(bufs[i], got[i]);
    }
};

std::mutex g_up_mu;
std::unordered_map<uint64_t, std::unique_ptr<Upload>> g_up;
uint64_t g_up_next = 1;
cudaStream_t g_up_stream = nullptr;

}  // namespace
}  // namespace pg

using namespace pg;

extern "C" pg_status pg_files_upload_begin(const pg_file_desc *files, int32_t n_files, uint64_t *out_upload) {
    if (!out_upload || n_files < 0 || (n_files > 0 && !files)) return fail(PG_ERR_INVALID, "null argument");
    pg_status st = require_device();
    if (st) return st;
    auto up = std::make_unique<Upload>();
    {
        std::lock_guard<std::mutex> lk(g_up_mu);
        if (!g_up_stream) PG_CUDA(cudaStreamCreateWithFlags(&g_up_stream, cudaStreamNonBlocking));
    }
    PG_CUDA(cudaEventCreateWithFlags(&up->done, cudaEventDisableTiming));
    for (int i = 0; i < n_files; i++) {
        if (files[i].size < 0 || (files[i].size > 0 && !files[i].bytes)) return fail(PG_ERR_INVALID, "upload: bad file descriptor");
        pg_file_desc d = files[i];
        if (files[i].mem == PG_MEM_HOST) {
            size_t got = 0;
            void *b = device_buffer_take((size_t)files[i].size + 64, &got);       // (readers may look 8 bytes past a page)
            if (!b) return fail(PG_ERR_CUDA, "upload: out of device memory (" + std::to_string(files[i].size) + " bytes)");
            up->bufs.push_back(b);
            up->got.push_back(got);
            PG_CUDA(cudaMemcpyAsync(b, files[i].bytes, (size_t)files[i].size, cudaMemcpyHostToDevice, g_up_stream));
            d.bytes = (const uint8_t *)b;
            d.mem = PG_MEM_DEVICE;
        }
        up->files.push_back(d);
    }
    PG_CUDA(cudaEventRecord(up->done, g_up_stream));
    std::lock_guard<std::mutex> lk(g_up_mu);
    const uint64_t h = ((uint64_t)7 << 56) | g_up_next++;
    g_up.emplace(h, std::move(up));
    *out_upload = h;
    return PG_OK;
}

extern "C" pg_status pg_files_upload_wait(uint64_t upload, pg_file_desc *out_files, int32_t n_files) {
    Upload *up;
    {
        std::lock_guard<std::mutex> lk(g_up_mu);
        auto it = g_up.find(upload);
        if (it == g_up.end()) return fail(PG_ERR_INVALID, "unknown upload handle");
        up = it->second.get();
    }
    if (n_files != (int32_t)up->files.size() || (n_files > 0 && !out_files)) return fail(PG_ERR_INVALID, "upload: one descriptor per file");
    PG_CUDA(cudaEventSynchronize(up->done));
    for (int i = 0; i < n_files; i++) out_files[i] = up->files[i];
    return PG_OK;
}

extern "C" pg_status pg_files_upload_free(uint64_t upload) {
    std::unique_ptr<Upload> up;
    {
        std::lock_guard<std::mutex> lk(g_up_mu);
        auto it = g_up.find(upload);
        if (it == g_up.end()) return fail(PG_ERR_INVALID, "unknown upload handle");
        up = std::move(it->second);
        g_up.erase(it);
    }
    // the copy itself, and the decode launches of the calling thread that read the bytes, must be done before the
    // buffers go back to the cache
    cudaEventSynchronize(up->done);
    cudaStreamSynchronize(thread_stream());
    return PG_OK;
}
