// orc_meta.h — host-side ORC metadata: Protocol Buffers wire reader, PostScript, Footer, stripe footers.
// Replaces what the reference takes from orc-core 1.9.2 (org.apache.orc.impl.ReaderImpl / RecordReaderImpl behind
// paimon-format/src/main/java/org/apache/paimon/format/orc/OrcReaderFactory.java:98-163, createRecordReader :280-330);
// the dependency is not under /root/reference.  The layout restated here is the public ORC specification (file tail:
// footer, postscript, 1-byte postscript length; orc_proto.proto field numbers) and the protobuf wire format.
#pragma once

#include <stdint.h>

#include <string>
#include <vector>

namespace orc {

enum Compression { C_NONE = 0, C_ZLIB = 1, C_SNAPPY = 2, C_LZO = 3, C_LZ4 = 4, C_ZSTD = 5 };
enum TypeKind { K_BOOLEAN = 0, K_BYTE = 1, K_SHORT = 2, K_INT = 3, K_LONG = 4, K_FLOAT = 5, K_DOUBLE = 6, K_STRING = 7,
                K_BINARY = 8, K_TIMESTAMP = 9, K_LIST = 10, K_MAP = 11, K_STRUCT = 12, K_UNION = 13, K_DECIMAL = 14,
                K_DATE = 15, K_VARCHAR = 16, K_CHAR = 17, K_TIMESTAMP_INSTANT = 18 };
enum StreamKind { S_PRESENT = 0, S_DATA = 1, S_LENGTH = 2, S_DICTIONARY_DATA = 3, S_DICTIONARY_COUNT = 4, S_SECONDARY = 5,
                  S_ROW_INDEX = 6, S_BLOOM_FILTER = 7, S_BLOOM_FILTER_UTF8 = 8 };
enum EncodingKind { E_DIRECT = 0, E_DICTIONARY = 1, E_DIRECT_V2 = 2, E_DICTIONARY_V2 = 3 };

struct Type {
    int kind = -1;
    std::vector<uint32_t> subtypes;
    std::vector<std::string> field_names;
    uint32_t precision = 0, scale = 0;
};
struct StripeInfo {
    uint64_t offset = 0, index_length = 0, data_length = 0, footer_length = 0, rows = 0;
};
struct StreamInfo {
    int kind = 0;
    uint32_t column = 0;
    uint64_t length = 0;
    uint64_t offset = 0;          // absolute file offset (derived)
};
struct ColumnEncoding {
    int kind = 0;
    uint32_t dictionary_size = 0;
};
struct StripeFooter {
    std::vector<StreamInfo> streams;
    std::vector<ColumnEncoding> columns;
};
struct FileTail {
    int compression = C_NONE;
    uint64_t block_size = 262144;
    std::vector<uint32_t> version;
    uint64_t rows = 0;
    std::vector<Type> types;
    std::vector<StripeInfo> stripes;
    std::vector<StripeFooter> stripe_footers;
};

// Throws std::runtime_error on malformed / unsupported input (only NONE, ZLIB and ZSTD metadata is inflated).
FileTail parse_file(const uint8_t *file, int64_t size);

// ---- decode plan of one file: which streams to inflate, and one task per (stripe, wanted column)
struct PlanStream {
    uint64_t offset = 0, length = 0;     // in the file
    uint64_t out_bound = 0;              // upper bound of the inflated bytes
    uint64_t out_off = 0;                // position in the file's stream scratch (64-byte aligned)
};
struct PlanTask {
    int stripe = 0;
    int col = 0;                         // caller's column index
    int type_id = 0;                     // ORC type id (flat schema: file column + 1)
    int kind = 0, enc = 0, scale = 0;
    uint32_t dict_size = 0;
    uint64_t dict_off_base = 0;          // position in the dictionary-offset scratch (entries)
    int64_t row0 = 0, rows = 0;          // file-relative
    int s_present = -1, s_data = -1, s_length = -1, s_dict = -1, s_secondary = -1;   // indexes into Plan::streams
};
struct Plan {
    std::vector<PlanStream> streams;
    std::vector<PlanTask> tasks;
    uint64_t scratch_bytes = 0;          // inflated streams
    uint64_t dict_entries = 0;           // dictionary-offset scratch entries
};
// file_col_of[c] = the file's column (0-based child of the root struct) for caller column c, or < 0 = skip
Plan plan_file(const FileTail &t, const uint8_t *file, int64_t size, const std::vector<int> &file_col_of);

}  // namespace orc
