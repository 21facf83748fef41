// paimon_gpu_jni.cc — thin JNI shim over the C ABI of libpaimon_gpu.so (include/paimon_gpu.h).
//
// One Java native method per C function; handles are jlong, buffers are direct ByteBuffers
// (GetDirectBufferAddress), errors become java.lang.RuntimeException / UnsupportedOperationException /
// IllegalArgumentException with the message of pg_last_error() — the convention of the reference's only
// in-tree native code (paimon-tantivy/paimon-tantivy-jni/rust/src/lib.rs:32-35, 72-155).
//
// The build image has no JDK (no jni.h), so this file is compiled only where JAVA_HOME is set:
//   g++ -std=c++17 -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
//       jni/paimon_gpu_jni.cc -Lpaimon_b200 -lpaimon_gpu -o libpaimon_gpu_jni.so
// The Java side (org.apache.paimon.gpu.NativeMerge) is listed in INTEGRATION.md.
#if __has_include(<jni.h>)
#include <jni.h>

#include <string>
#include <vector>

#include "paimon_gpu.h"

namespace {

void throw_for(JNIEnv *env, pg_status st) {
    const char *cls = "java/lang/RuntimeException";
    if (st == PG_ERR_UNSUPPORTED) cls = "java/lang/UnsupportedOperationException";
    if (st == PG_ERR_INVALID || st == PG_ERR_MERGE_FUNCTION) cls = "java/lang/IllegalArgumentException";
    env->ThrowNew(env->FindClass(cls), pg_last_error());
}

#define PG_CHECK(expr)                        \
    do {                                      \
        pg_status _st = (expr);               \
        if (_st != PG_OK) {                   \
            throw_for(env, _st);              \
            return 0;                         \
        }                                     \
    } while (0)

}  // namespace

extern "C" {

JNIEXPORT jint JNICALL Java_org_apache_paimon_gpu_NativeMerge_init(JNIEnv *env, jclass, jint device) {
    PG_CHECK(pg_init(device));
    return 0;
}

// int[] keyTypes, int[] valTypes, boolean[] valNullable -> schema handle
JNIEXPORT jlong JNICALL Java_org_apache_paimon_gpu_NativeMerge_schemaCreate(JNIEnv *env, jclass, jintArray keyTypes,
                                                                             jintArray valTypes,
                                                                             jbooleanArray valNullable) {
    jsize nk = env->GetArrayLength(keyTypes), nv = env->GetArrayLength(valTypes);
    std::vector<jint> kt(nk), vt(nv);
    std::vector<jboolean> vn(nv);
    env->GetIntArrayRegion(keyTypes, 0, nk, kt.data());
    env->GetIntArrayRegion(valTypes, 0, nv, vt.data());
    env->GetBooleanArrayRegion(valNullable, 0, nv, vn.data());
    std::vector<pg_field> kf(nk), vf(nv);
    for (jsize i = 0; i < nk; i++) kf[i] = pg_field{kt[i], 0};
    for (jsize i = 0; i < nv; i++) vf[i] = pg_field{vt[i], vn[i] ? 1 : 0};
    pg_schema_desc d{nk, nv, kf.data(), vf.data()};
    uint64_t h = 0;
    PG_CHECK(pg_schema_create(&d, &h));
    return (jlong)h;
}

JNIEXPORT jint JNICALL Java_org_apache_paimon_gpu_NativeMerge_schemaFree(JNIEnv *env, jclass, jlong h) {
    PG_CHECK(pg_schema_free((uint64_t)h));
    return 0;
}

// the declarative MergeFunction (what MergeFunctionFactory.create(readType) would have built)
JNIEXPORT jlong JNICALL Java_org_apache_paimon_gpu_NativeMerge_mergeSpecCreate(
    JNIEnv *env, jclass, jlong schema, jint engine, jboolean ignoreDelete, jboolean removeRecordOnDelete,
    jboolean dropDelete, jintArray seqFields, jboolean seqAscending, jintArray agg, jbooleanArray ignoreRetract,
    jintArray groupSeqStart, jintArray groupSeqFields, jintArray fieldGroup, jbooleanArray groupPartialDelete) {
    jsize ns = seqFields ? env->GetArrayLength(seqFields) : 0;
    jsize nv = agg ? env->GetArrayLength(agg) : 0;
    std::vector<jint> sf(ns), ag(nv);
    std::vector<jboolean> ir(nv);
    if (ns) env->GetIntArrayRegion(seqFields, 0, ns, sf.data());
    if (nv) env->GetIntArrayRegion(agg, 0, nv, ag.data());
    if (nv && ignoreRetract) env->GetBooleanArrayRegion(ignoreRetract, 0, nv, ir.data());
    std::vector<uint8_t> ir8(ir.begin(), ir.end());
    pg_merge_spec sp{};
    sp.engine = engine;
    sp.ignore_delete = ignoreDelete;
    sp.remove_record_on_delete = removeRecordOnDelete;
    sp.drop_delete = dropDelete;
    sp.n_seq_fields = ns;
    sp.seq_fields = ns ? sf.data() : nullptr;
    sp.seq_ascending = seqAscending;
    sp.agg = nv ? ag.data() : nullptr;
    sp.ignore_retract = nv ? ir8.data() : nullptr;
    // partial-update sequence groups (PartialUpdateMergeFunction.Factory: fields.<seq>.sequence-group)
    std::vector<jint> gs, gf, fg;
    std::vector<uint8_t> gpd8;
    if (groupSeqStart && env->GetArrayLength(groupSeqStart) > 1) {
        gs.resize(env->GetArrayLength(groupSeqStart));
        gf.resize(env->GetArrayLength(groupSeqFields));
        fg.resize(nv);
        env->GetIntArrayRegion(groupSeqStart, 0, (jsize)gs.size(), gs.data());
        env->GetIntArrayRegion(groupSeqFields, 0, (jsize)gf.size(), gf.data());
        env->GetIntArrayRegion(fieldGroup, 0, nv, fg.data());
        if (groupPartialDelete) {
            std::vector<jboolean> gpd(nv);
            env->GetBooleanArrayRegion(groupPartialDelete, 0, nv, gpd.data());
            gpd8.assign(gpd.begin(), gpd.end());
        }
        sp.n_sequence_groups = (int32_t)gs.size() - 1;
        sp.group_seq_start = gs.data();
        sp.group_seq_fields = gf.data();
        sp.field_group = fg.data();
        sp.group_partial_delete = gpd8.empty() ? nullptr : gpd8.data();
    }
    uint64_t h = 0;
    PG_CHECK(pg_merge_spec_create((uint64_t)schema, &sp, &h));
    return (jlong)h;
}

JNIEXPORT jint JNICALL Java_org_apache_paimon_gpu_NativeMerge_mergeSpecFree(JNIEnv *env, jclass, jlong h) {
    PG_CHECK(pg_merge_spec_free((uint64_t)h));
    return 0;
}

// One sorted run from direct ByteBuffers: per column {data, offsets|null, validity|null}
JNIEXPORT jlong JNICALL Java_org_apache_paimon_gpu_NativeMerge_runOpen(JNIEnv *env, jclass, jlong schema,
                                                                        jlong nRows, jobjectArray data,
                                                                        jobjectArray offsets, jobjectArray validity) {
    jsize nc = env->GetArrayLength(data);
    std::vector<pg_column> cols(nc);
    for (jsize c = 0; c < nc; c++) {
        jobject d = env->GetObjectArrayElement(data, c);
        jobject o = env->GetObjectArrayElement(offsets, c);
        jobject v = env->GetObjectArrayElement(validity, c);
        cols[c].data = d ? env->GetDirectBufferAddress(d) : nullptr;
        cols[c].offsets = o ? (const int32_t *)env->GetDirectBufferAddress(o) : nullptr;
        cols[c].validity = v ? (const uint8_t *)env->GetDirectBufferAddress(v) : nullptr;
    }
    pg_run_desc rd{nRows, cols.data()};
    uint64_t h = 0;
    PG_CHECK(pg_run_open((uint64_t)schema, &rd, PG_MEM_HOST, &h));
    return (jlong)h;
}

JNIEXPORT jint JNICALL Java_org_apache_paimon_gpu_NativeMerge_runFree(JNIEnv *env, jclass, jlong h) {
    PG_CHECK(pg_run_free((uint64_t)h));
    return 0;
}

JNIEXPORT jlong JNICALL Java_org_apache_paimon_gpu_NativeMerge_mergeOpen(JNIEnv *env, jclass, jlong spec,
                                                                          jlongArray runs) {
    jsize k = env->GetArrayLength(runs);
    std::vector<jlong> r(k);
    env->GetLongArrayRegion(runs, 0, k, r.data());
    std::vector<uint64_t> ru(r.begin(), r.end());
    uint64_t h = 0;
    PG_CHECK(pg_merge_open((uint64_t)spec, ru.data(), k, &h));
    return (jlong)h;
}

// Re-use the merge handle for other runs of the same schema (the key-range streaming reader)
JNIEXPORT jint JNICALL Java_org_apache_paimon_gpu_NativeMerge_mergeRebind(JNIEnv *env, jclass, jlong merge,
                                                                          jlongArray runs, jlongArray startRows) {
    jsize k = env->GetArrayLength(runs);
    std::vector<jlong> r(k), sr(k, 0);
    env->GetLongArrayRegion(runs, 0, k, r.data());
    if (startRows) env->GetLongArrayRegion(startRows, 0, k, sr.data());
    std::vector<uint64_t> ru(r.begin(), r.end());
    std::vector<int64_t> st(sr.begin(), sr.end());
    PG_CHECK(pg_merge_rebind((uint64_t)merge, ru.data(), k, st.data()));
    return 0;
}

// Format seam: one Parquet data file (bytes read by the Java FileIO into a direct buffer) -> device-resident run
JNIEXPORT jlong JNICALL Java_org_apache_paimon_gpu_NativeMerge_parquetOpen(JNIEnv *env, jclass, jlong schema,
                                                                           jobject fileBytes, jlong size) {
    uint64_t h = 0;
    PG_CHECK(pg_parquet_open((uint64_t)schema, (const uint8_t *)env->GetDirectBufferAddress(fileBytes), (int64_t)size, &h));
    return (jlong)h;
}
JNIEXPORT jlong JNICALL Java_org_apache_paimon_gpu_NativeMerge_parquetReadRun(JNIEnv *env, jclass, jlong file) {
    uint64_t run = 0;
    PG_CHECK(pg_parquet_read_run((uint64_t)file, &run));
    return (jlong)run;
}
// ApplyDeletionVectorReader: the run minus the rows whose file position is set in the (expanded) deletion vector
JNIEXPORT jlong JNICALL Java_org_apache_paimon_gpu_NativeMerge_applyDeletionVector(JNIEnv *env, jclass, jlong run,
                                                                                   jobject bitmap, jlong nBits) {
    uint64_t out = 0;
    PG_CHECK(pg_run_apply_deletion_vector((uint64_t)run, (const uint8_t *)env->GetDirectBufferAddress(bitmap), nBits, &out));
    return (jlong)out;
}
JNIEXPORT jint JNICALL Java_org_apache_paimon_gpu_NativeMerge_parquetFree(JNIEnv *env, jclass, jlong file) {
    PG_CHECK(pg_parquet_free((uint64_t)file));
    return 0;
}

// Compaction output encode: rows [row0, row0 + nRows) of a merge batch / run -> one Parquet file on the device
JNIEXPORT jlong JNICALL Java_org_apache_paimon_gpu_NativeMerge_parquetEncode(JNIEnv *env, jclass, jlong source,
                                                                             jobjectArray names, jlong row0,
                                                                             jlong nRows, jlong rowGroupRows,
                                                                             jlong pageRows) {
    jsize nc = env->GetArrayLength(names);
    std::vector<std::string> keep(nc);
    std::vector<const char *> ptrs(nc);
    for (jsize c = 0; c < nc; c++) {
        jstring js = (jstring)env->GetObjectArrayElement(names, c);
        const char *u = env->GetStringUTFChars(js, nullptr);
        keep[c] = u;
        env->ReleaseStringUTFChars(js, u);
        ptrs[c] = keep[c].c_str();
    }
    pg_parquet_write_options opt{rowGroupRows, pageRows};
    uint64_t h = 0;
    PG_CHECK(pg_parquet_encode((uint64_t)source, ptrs.data(), row0, nRows, &opt, &h));
    return (jlong)h;
}
JNIEXPORT jlongArray JNICALL Java_org_apache_paimon_gpu_NativeMerge_fileMeta(JNIEnv *env, jclass, jlong file) {
    pg_file_meta m{};
    pg_status fst = pg_parquet_file_meta((uint64_t)file, &m);
    if (fst != PG_OK) { throw_for(env, fst); return nullptr; }
    jlong v[7] = {m.n_rows, m.file_bytes, m.min_sequence_number, m.max_sequence_number, m.delete_row_count,
                  m.n_row_groups, m.n_pages};
    jlongArray out = env->NewLongArray(7);
    env->SetLongArrayRegion(out, 0, 7, v);
    return out;
}
JNIEXPORT jint JNICALL Java_org_apache_paimon_gpu_NativeMerge_fileFetch(JNIEnv *env, jclass, jlong file, jobject dst) {
    PG_CHECK(pg_parquet_file_fetch((uint64_t)file, env->GetDirectBufferAddress(dst), env->GetDirectBufferCapacity(dst)));
    return 0;
}
JNIEXPORT jint JNICALL Java_org_apache_paimon_gpu_NativeMerge_fileFree(JNIEnv *env, jclass, jlong file) {
    PG_CHECK(pg_parquet_file_free((uint64_t)file));
    return 0;
}

// readBatch(): runs the merge; returns the row count.  Column sizes follow via batchColumnBytes().
JNIEXPORT jlong JNICALL Java_org_apache_paimon_gpu_NativeMerge_mergeExecute(JNIEnv *env, jclass, jlong merge) {
    PG_CHECK(pg_merge_execute((uint64_t)merge));
    pg_batch b{};
    PG_CHECK(pg_merge_device_batch((uint64_t)merge, &b));
    return (jlong)b.n_rows;
}

// long[3*nCols]: {dataBytes, hasOffsets, hasValidity} per column, so Java can size its direct buffers
JNIEXPORT jlongArray JNICALL Java_org_apache_paimon_gpu_NativeMerge_batchLayout(JNIEnv *env, jclass, jlong merge) {
    pg_batch b{};
    pg_status st = pg_merge_device_batch((uint64_t)merge, &b);
    if (st != PG_OK) { throw_for(env, st); return nullptr; }
    std::vector<jlong> out(3 * (size_t)b.n_cols);
    for (int c = 0; c < b.n_cols; c++) {
        out[3 * c] = b.cols[c].data_bytes;
        out[3 * c + 1] = b.cols[c].offsets != nullptr;
        out[3 * c + 2] = b.cols[c].validity != nullptr;
    }
    jlongArray arr = env->NewLongArray((jsize)out.size());
    env->SetLongArrayRegion(arr, 0, (jsize)out.size(), out.data());
    return arr;
}

// copy the merged batch into caller-owned direct buffers (what ArrowBatchReader then wraps)
JNIEXPORT jint JNICALL Java_org_apache_paimon_gpu_NativeMerge_mergeFetch(JNIEnv *env, jclass, jlong merge,
                                                                          jobjectArray data, jobjectArray offsets,
                                                                          jobjectArray validity) {
    jsize nc = env->GetArrayLength(data);
    std::vector<pg_out_column> cols(nc);
    for (jsize c = 0; c < nc; c++) {
        jobject d = env->GetObjectArrayElement(data, c);
        jobject o = env->GetObjectArrayElement(offsets, c);
        jobject v = env->GetObjectArrayElement(validity, c);
        cols[c].data = d ? env->GetDirectBufferAddress(d) : nullptr;
        cols[c].offsets = o ? (int32_t *)env->GetDirectBufferAddress(o) : nullptr;
        cols[c].validity = v ? (uint8_t *)env->GetDirectBufferAddress(v) : nullptr;
        cols[c].data_bytes = d ? env->GetDirectBufferCapacity(d) : 0;
    }
    PG_CHECK(pg_merge_fetch((uint64_t)merge, cols.data(), nc));
    return 0;
}

JNIEXPORT jint JNICALL Java_org_apache_paimon_gpu_NativeMerge_mergeRelease(JNIEnv *env, jclass, jlong merge) {
    PG_CHECK(pg_merge_release((uint64_t)merge));
    return 0;
}

JNIEXPORT jint JNICALL Java_org_apache_paimon_gpu_NativeMerge_mergeFree(JNIEnv *env, jclass, jlong merge) {
    PG_CHECK(pg_merge_free((uint64_t)merge));
    return 0;
}

}  // extern "C"
#endif  // __has_include(<jni.h>)
