// parquet_meta.h — host-side Parquet metadata: Thrift compact-protocol reader, footer (FileMetaData) and
// page headers.  Replaces what the reference takes from parquet-mr 1.16.0
// (org.apache.parquet.format.* via PQ3P/hadoop/ParquetFileReader.java:277-334 footer read,
// :1345 Chunk.readAllPages page-header loop).  The algorithm restated here is the public Parquet format
// specification (parquet-format: "Thrift Compact Protocol" + parquet.thrift field ids), which is what
// parquet-mr implements; the dependency itself is not vendored in /root/reference.
#pragma once

#include <stdint.h>

#include <string>
#include <vector>

namespace pq {

enum PhysType { T_BOOLEAN = 0, T_INT32 = 1, T_INT64 = 2, T_INT96 = 3, T_FLOAT = 4, T_DOUBLE = 5, T_BYTE_ARRAY = 6,
                T_FIXED_LEN_BYTE_ARRAY = 7 };
enum Encoding { E_PLAIN = 0, E_PLAIN_DICTIONARY = 2, E_RLE = 3, E_BIT_PACKED = 4, E_DELTA_BINARY_PACKED = 5,
                E_DELTA_LENGTH_BYTE_ARRAY = 6, E_DELTA_BYTE_ARRAY = 7, E_RLE_DICTIONARY = 8, E_BYTE_STREAM_SPLIT = 9 };
enum Codec { C_UNCOMPRESSED = 0, C_SNAPPY = 1, C_GZIP = 2, C_LZO = 3, C_BROTLI = 4, C_LZ4 = 5, C_ZSTD = 6, C_LZ4_RAW = 7 };
enum PageType { P_DATA = 0, P_INDEX = 1, P_DICTIONARY = 2, P_DATA_V2 = 3 };
enum Repetition { R_REQUIRED = 0, R_OPTIONAL = 1, R_REPEATED = 2 };

struct SchemaElement {
    int32_t type = -1;            // PhysType; -1 for groups
    int32_t type_length = 0;
    int32_t repetition = 0;
    std::string name;
    int32_t num_children = 0;
    int32_t converted_type = -1;
};

struct ColumnChunk {
    int32_t type = -1;
    int32_t codec = 0;
    int64_t num_values = 0;
    int64_t total_uncompressed_size = 0;
    int64_t total_compressed_size = 0;
    int64_t data_page_offset = 0;
    int64_t dictionary_page_offset = 0;   // 0 = none
    std::vector<int32_t> encodings;
    std::vector<std::string> path;
    int64_t start() const {
        return (dictionary_page_offset > 0 && dictionary_page_offset < data_page_offset) ? dictionary_page_offset
                                                                                        : data_page_offset;
    }
};

struct RowGroup {
    int64_t num_rows = 0;
    int64_t total_byte_size = 0;
    std::vector<ColumnChunk> columns;
};

struct FileMetaData {
    int32_t version = 0;
    int64_t num_rows = 0;
    std::vector<SchemaElement> schema;   // flattened, root first
    std::vector<RowGroup> row_groups;
    std::string created_by;
};

struct PageHeader {
    int32_t type = -1;
    int32_t uncompressed_size = 0;
    int32_t compressed_size = 0;
    int32_t num_values = 0;
    int32_t encoding = 0;
    int32_t def_level_encoding = E_RLE;
    // v2
    int32_t num_nulls = 0, num_rows = 0;
    int32_t def_levels_byte_length = 0, rep_levels_byte_length = 0;
    bool is_compressed = true;
    int32_t header_size = 0;      // bytes the Thrift header occupied
};

// Throws std::runtime_error on malformed input.
FileMetaData parse_footer(const uint8_t *file, int64_t size);
// the pieces of parse_footer, for files whose bytes live on the device: the last 8 bytes of the file
// ([footer length:4 LE]["PAR1"]) -> footer length; the Thrift FileMetaData bytes in front of them -> metadata
int64_t footer_length(const uint8_t *tail8);
FileMetaData parse_footer_thrift(const uint8_t *footer, int64_t flen);
PageHeader parse_page_header(const uint8_t *p, int64_t avail);

}  // namespace pq
