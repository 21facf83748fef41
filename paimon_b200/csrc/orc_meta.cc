// orc_meta.cc — protobuf wire reader + ORC file tail / stripe footers (see orc_meta.h).
#include "orc_meta.h"

#include <algorithm>
#include <stdexcept>

#include "inflate_device.cuh"
#include "zstd_device.cuh"

namespace orc {

namespace {

struct Pb {
    const uint8_t *p, *end;
    Pb(const uint8_t *b, size_t n) : p(b), end(b + n) {}
    bool done() const { return p >= end; }
    uint64_t varint() {
        uint64_t v = 0;
        for (int sh = 0; sh < 70; sh += 7) {
            if (p >= end) throw std::runtime_error("orc: truncated protobuf varint");
            const uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << sh;
            if (!(b & 0x80)) return v;
        }
        throw std::runtime_error("orc: bad protobuf varint");
    }
    // next field: returns false at the end; wire types 0 varint, 1 fixed64, 2 length-delimited, 5 fixed32
    bool next(uint32_t &field, int &wire) {
        if (p >= end) return false;
        const uint64_t key = varint();
        field = (uint32_t)(key >> 3);
        wire = (int)(key & 7);
        return true;
    }
    Pb bytes() {
        const uint64_t n = varint();
        if ((uint64_t)(end - p) < n) throw std::runtime_error("orc: truncated protobuf field");
        Pb sub(p, (size_t)n);
        p += n;
        return sub;
    }
    void skip(int wire) {
        switch (wire) {
            case 0: varint(); return;
            case 1: if (end - p < 8) throw std::runtime_error("orc: truncated protobuf"); p += 8; return;
            case 2: bytes(); return;
            case 5: if (end - p < 4) throw std::runtime_error("orc: truncated protobuf"); p += 4; return;
            default: throw std::runtime_error("orc: unsupported protobuf wire type");
        }
    }
    // a repeated uint32 that may be packed (wire 2) or not (wire 0)
    void repeated_u32(int wire, std::vector<uint32_t> &out) {
        if (wire == 0) { out.push_back((uint32_t)varint()); return; }
        Pb sub = bytes();
        while (!sub.done()) out.push_back((uint32_t)sub.varint());
    }
};

// a metadata section stored as compression chunks: [3-byte header: length << 1 | isOriginal][bytes]...
std::vector<uint8_t> inflate_section(const uint8_t *p, uint64_t n, int codec, uint64_t block_size) {
    if (codec == C_NONE) return std::vector<uint8_t>(p, p + n);
    if (codec != C_ZLIB && codec != C_ZSTD)
        throw std::runtime_error("orc: compression kind " + std::to_string(codec) + " is not decoded (NONE, ZLIB and ZSTD are)");
    std::vector<uint8_t> out;
    std::vector<uint8_t> buf(block_size + 64), lit(zs::kMaxBlock + 64);
    inflate::Tables it;
    static thread_local zs::Tables *zt = nullptr;
    if (!zt) zt = new zs::Tables();
    uint64_t pos = 0;
    while (pos < n) {
        if (n - pos < 3) throw std::runtime_error("orc: truncated compression chunk header");
        const uint32_t h = p[pos] | (p[pos + 1] << 8) | (p[pos + 2] << 16);
        const uint32_t len = h >> 1;
        pos += 3;
        if (n - pos < len) throw std::runtime_error("orc: truncated compression chunk");
        if (h & 1) out.insert(out.end(), p + pos, p + pos + len);
        else {
            const int64_t got = codec == C_ZLIB ? inflate::inflate_raw(p + pos, len, buf.data(), (int64_t)block_size, it, nullptr)
                                                : zs::decode(p + pos, len, buf.data(), (int64_t)block_size, lit.data(), *zt);
            if (got < 0) throw std::runtime_error("orc: a metadata compression chunk does not inflate");
            out.insert(out.end(), buf.data(), buf.data() + got);
        }
        pos += len;
    }
    return out;
}

Type read_type(Pb pb) {
    Type t;
    uint32_t f;
    int w;
    while (pb.next(f, w)) {
        if (f == 1 && w == 0) t.kind = (int)pb.varint();
        else if (f == 2) pb.repeated_u32(w, t.subtypes);
        else if (f == 3 && w == 2) { Pb s = pb.bytes(); t.field_names.emplace_back((const char *)s.p, (size_t)(s.end - s.p)); }
        else if (f == 5 && w == 0) t.precision = (uint32_t)pb.varint();
        else if (f == 6 && w == 0) t.scale = (uint32_t)pb.varint();
        else pb.skip(w);
    }
    return t;
}

StripeInfo read_stripe_info(Pb pb) {
    StripeInfo s;
    uint32_t f;
    int w;
    while (pb.next(f, w)) {
        if (w != 0) { pb.skip(w); continue; }
        const uint64_t v = pb.varint();
        if (f == 1) s.offset = v;
        else if (f == 2) s.index_length = v;
        else if (f == 3) s.data_length = v;
        else if (f == 4) s.footer_length = v;
        else if (f == 5) s.rows = v;
    }
    return s;
}

StripeFooter read_stripe_footer(Pb pb) {
    StripeFooter sf;
    uint32_t f;
    int w;
    while (pb.next(f, w)) {
        if (f == 1 && w == 2) {
            Pb s = pb.bytes();
            StreamInfo si;
            uint32_t f2;
            int w2;
            while (s.next(f2, w2)) {
                if (w2 != 0) { s.skip(w2); continue; }
                const uint64_t v = s.varint();
                if (f2 == 1) si.kind = (int)v;
                else if (f2 == 2) si.column = (uint32_t)v;
                else if (f2 == 3) si.length = v;
            }
            sf.streams.push_back(si);
        } else if (f == 2 && w == 2) {
            Pb s = pb.bytes();
            ColumnEncoding ce;
            uint32_t f2;
            int w2;
            while (s.next(f2, w2)) {
                if (w2 != 0) { s.skip(w2); continue; }
                const uint64_t v = s.varint();
                if (f2 == 1) ce.kind = (int)v;
                else if (f2 == 2) ce.dictionary_size = (uint32_t)v;
            }
            sf.columns.push_back(ce);
        } else pb.skip(w);
    }
    return sf;
}

}  // namespace

FileTail parse_file(const uint8_t *file, int64_t size) {
    if (size < 4 || file[0] != 'O' || file[1] != 'R' || file[2] != 'C') throw std::runtime_error("orc: missing ORC magic");
    const uint64_t ps_len = file[size - 1];
    if ((int64_t)ps_len + 1 > size) throw std::runtime_error("orc: bad postscript length");
    FileTail t;
    uint64_t footer_len = 0;
    {
        Pb pb(file + size - 1 - ps_len, (size_t)ps_len);
        uint32_t f;
        int w;
        while (pb.next(f, w)) {
            if (f == 1 && w == 0) footer_len = pb.varint();
            else if (f == 2 && w == 0) t.compression = (int)pb.varint();
            else if (f == 3 && w == 0) t.block_size = pb.varint();
            else if (f == 4) pb.repeated_u32(w, t.version);
            else pb.skip(w);
        }
    }
    if (footer_len + ps_len + 1 > (uint64_t)size) throw std::runtime_error("orc: bad footer length");
    const std::vector<uint8_t> footer = inflate_section(file + size - 1 - ps_len - footer_len, footer_len, t.compression, t.block_size);
    {
        Pb pb(footer.data(), footer.size());
        uint32_t f;
        int w;
        while (pb.next(f, w)) {
            if (f == 3 && w == 2) t.stripes.push_back(read_stripe_info(pb.bytes()));
            else if (f == 4 && w == 2) t.types.push_back(read_type(pb.bytes()));
            else if (f == 6 && w == 0) t.rows = pb.varint();
            else pb.skip(w);
        }
    }
    for (const StripeInfo &si : t.stripes) {
        const uint64_t fo = si.offset + si.index_length + si.data_length;
        if (fo + si.footer_length > (uint64_t)size) throw std::runtime_error("orc: stripe footer outside the file");
        const std::vector<uint8_t> raw = inflate_section(file + fo, si.footer_length, t.compression, t.block_size);
        StripeFooter sf = read_stripe_footer(Pb(raw.data(), raw.size()));
        uint64_t off = si.offset;
        for (StreamInfo &s : sf.streams) {
            s.offset = off;
            off += s.length;
        }
        if (off > fo) throw std::runtime_error("orc: stream lengths exceed the stripe");
        t.stripe_footers.push_back(std::move(sf));
    }
    return t;
}

Plan plan_file(const FileTail &t, const uint8_t *file, int64_t size, const std::vector<int> &file_col_of) {
    Plan pl;
    if (t.types.empty() || t.types[0].kind != K_STRUCT) throw std::runtime_error("orc: the root type is not a struct");
    int64_t row0 = 0;
    for (size_t si = 0; si < t.stripes.size(); si++) {
        const StripeFooter &sf = t.stripe_footers[si];
        for (size_t c = 0; c < file_col_of.size(); c++) {
            if (file_col_of[c] < 0) continue;
            if ((size_t)file_col_of[c] >= t.types[0].subtypes.size()) throw std::runtime_error("orc: column outside the file schema");
            const uint32_t tid = t.types[0].subtypes[file_col_of[c]];
            if (tid >= t.types.size() || tid >= sf.columns.size()) throw std::runtime_error("orc: type id outside the footer");
            PlanTask task;
            task.stripe = (int)si;
            task.col = (int)c;
            task.type_id = (int)tid;
            task.kind = t.types[tid].kind;
            task.scale = (int)t.types[tid].scale;
            task.enc = sf.columns[tid].kind;
            task.dict_size = sf.columns[tid].dictionary_size;
            task.row0 = row0;
            task.rows = (int64_t)t.stripes[si].rows;
            for (const StreamInfo &s : sf.streams) {
                if (s.column != tid || s.length == 0) continue;
                int *slot = nullptr;
                if (s.kind == S_PRESENT) slot = &task.s_present;
                else if (s.kind == S_DATA) slot = &task.s_data;
                else if (s.kind == S_LENGTH) slot = &task.s_length;
                else if (s.kind == S_DICTIONARY_DATA) slot = &task.s_dict;
                else if (s.kind == S_SECONDARY) slot = &task.s_secondary;
                if (!slot) continue;                      // row indexes, bloom filters
                if (s.offset + s.length > (uint64_t)size) throw std::runtime_error("orc: stream outside the file");
                PlanStream ps;
                ps.offset = s.offset;
                ps.length = s.length;
                if (t.compression == C_NONE) ps.out_bound = s.length;
                else {
                    // walk the chunk headers: an original chunk keeps its size, a compressed one inflates to at most
                    // the block size (DEFLATE cannot expand more than 1032 : 1)
                    uint64_t pos = 0, bound = 0;
                    while (pos < s.length) {
                        if (s.length - pos < 3) throw std::runtime_error("orc: truncated compression chunk header");
                        const uint8_t *h = file + s.offset + pos;
                        const uint32_t hv = h[0] | (h[1] << 8) | (h[2] << 16);
                        const uint64_t len = hv >> 1;
                        if (s.length - pos - 3 < len) throw std::runtime_error("orc: truncated compression chunk");
                        if (hv & 1) bound += len;
                        else if (t.compression == C_ZLIB) bound += std::min<uint64_t>(t.block_size, len * 1032 + 64);
                        else bound += t.block_size;
                        pos += 3 + len;
                    }
                    ps.out_bound = bound;
                }
                ps.out_off = pl.scratch_bytes;
                pl.scratch_bytes += (ps.out_bound + 64 + 63) & ~(uint64_t)63;
                *slot = (int)pl.streams.size();
                pl.streams.push_back(ps);
            }
            if (task.enc == E_DICTIONARY || task.enc == E_DICTIONARY_V2) {
                task.dict_off_base = pl.dict_entries;
                pl.dict_entries += (uint64_t)task.dict_size + 1;
            }
            pl.tasks.push_back(task);
        }
        row0 += (int64_t)t.stripes[si].rows;
    }
    return pl;
}

}  // namespace orc
