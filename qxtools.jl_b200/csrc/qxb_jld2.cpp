// qxb200 -- JLD2 / HDF5-subset reader and writer (see qxb_jld2.h for scope and citations).
//
// Written from the HDF5 file-format specification (superblock, object headers, messages 0x01 dataspace,
// 0x02 link info, 0x03 datatype, 0x06 link, 0x08 layout, 0x10 continuation, 0x11 symbol table; v1 B-tree
// group nodes, local heaps) -- no HDF5 library exists in this image.  The reader is pinned against a file
// written by the HDF5 C library (tests/test_jld2.py); the JLD2-shaped layout (v2 structures, committed
// compound datatype) is pinned only against this writer until a file from JLD2.jl is available.
#include "qxb_jld2.h"

#include <cstdio>
#include <cstring>
#include <set>

#include "../../include/qxb200.h"
#include "qxb_ir.h"

namespace qxb {
namespace jld2 {

// ------------------------------------------------------------------ lookup3
static inline uint32_t rot(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

uint32_t lookup3(const uint8_t* k, size_t length, uint32_t initval) {
    uint32_t a, b, c;
    a = b = c = 0xdeadbeefu + (uint32_t)length + initval;
    auto le = [](const uint8_t* p, int n) { uint32_t v = 0; for (int i = 0; i < n; ++i) v |= (uint32_t)p[i] << (8 * i); return v; };
    while (length > 12) {
        a += le(k, 4); b += le(k + 4, 4); c += le(k + 8, 4);
        a -= c; a ^= rot(c, 4);  c += b;
        b -= a; b ^= rot(a, 6);  a += c;
        c -= b; c ^= rot(b, 8);  b += a;
        a -= c; a ^= rot(c, 16); c += b;
        b -= a; b ^= rot(a, 19); a += c;
        c -= b; c ^= rot(b, 4);  b += a;
        length -= 12; k += 12;
    }
    if (length == 0) return c;
    a += le(k, (int)(length < 4 ? length : 4));
    if (length > 4) b += le(k + 4, (int)(length < 8 ? length - 4 : 4));
    if (length > 8) c += le(k + 8, (int)(length - 8));
    c ^= b; c -= rot(b, 14);
    a ^= c; a -= rot(c, 11);
    b ^= a; b -= rot(a, 25);
    c ^= b; c -= rot(b, 16);
    a ^= c; a -= rot(c, 4);
    b ^= a; b -= rot(a, 14);
    c ^= b; c -= rot(b, 24);
    return c;
}

const Dataset* File::find(const std::string& name) const {
    for (const Dataset& d : datasets) if (d.name == name) return &d;
    return nullptr;
}

namespace {

const uint8_t kSig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
const uint64_t kUndef = ~0ull;

[[noreturn]] void bad(const std::string& path, uint64_t off, const std::string& what) {
    throw Error(QXB_ERR_ARG, path + ": offset " + std::to_string(off) + ": " + what);
}
[[noreturn]] void unsupp(const std::string& path, const std::string& what) {
    throw Error(QXB_ERR_UNSUPP, path + ": " + what);
}

struct Msg { int type; int flags; uint64_t off; uint64_t size; };

struct DType {
    int kind = EK_OTHER;
    int size = 0;
    bool is_signed = true;
    uint64_t len = 0;     // encoded length; 0 when the class is not one we can step over
};

struct Reader {
    File& f;
    std::vector<uint8_t> buf;
    int so = 8, sl = 8;
    std::set<uint64_t> visited;

    explicit Reader(File& file) : f(file) {}

    void need(uint64_t off, uint64_t n) const {
        if (off > buf.size() || n > buf.size() - off) bad(f.path, off, "read of " + std::to_string(n) + " bytes past the end of the file");
    }
    uint64_t rd(uint64_t off, int n) const {
        need(off, n);
        uint64_t v = 0;
        for (int i = 0; i < n; ++i) v |= (uint64_t)buf[off + i] << (8 * i);
        return v;
    }
    // a file address ("size of offsets" bytes); all ones = undefined
    uint64_t ra(uint64_t off) const {
        uint64_t v = rd(off, so);
        return (so < 8 && v == (1ull << (8 * so)) - 1) ? kUndef : v;
    }
    uint64_t abs_addr(uint64_t rel) const { return rel == kUndef ? kUndef : rel + f.base_address; }
    bool is(uint64_t off, const char* sig4) const { need(off, 4); return memcmp(&buf[off], sig4, 4) == 0; }
    void check_sum(uint64_t begin, uint64_t end) {
        need(end, 4);
        if (lookup3(&buf[begin], end - begin) != (uint32_t)rd(end, 4)) ++f.checksum_failures;
    }

    // offset / length sizes steer every later read (rd shifts by 8 * size): only 2, 4 and 8 bytes exist
    void check_sizes(uint64_t sb) {
        if ((so != 2 && so != 4 && so != 8) || (sl != 2 && sl != 4 && sl != 8))
            bad(f.path, sb, "superblock: size of offsets " + std::to_string(so) + " / size of lengths " + std::to_string(sl) +
                            " (must be 2, 4 or 8)");
    }

    // ---------------------------------------------------------- superblock
    uint64_t superblock() {
        uint64_t sb = kUndef;
        for (uint64_t off = 0; off + 8 <= buf.size(); off = off ? off * 2 : 512)
            if (memcmp(&buf[off], kSig, 8) == 0) { sb = off; break; }
        if (sb == kUndef) bad(f.path, 0, "no HDF5 superblock signature at 0, 512, 1024, ... (not a JLD2/HDF5 file)");
        f.header.assign((const char*)buf.data(), strnlen((const char*)buf.data(), sb));
        int ver = (int)rd(sb + 8, 1);
        f.superblock_version = ver;
        if (ver == 0 || ver == 1) {
            so = (int)rd(sb + 13, 1); sl = (int)rd(sb + 14, 1);
            check_sizes(sb);
            uint64_t p = sb + 24 + (ver == 1 ? 4 : 0);
            f.base_address = ra(p);
            p += 4 * so;                                  // base, free-space, eof, driver-info
            return abs_addr(ra(p + so));              // root symbol-table entry: name offset, header address
        }
        if (ver == 2 || ver == 3) {
            so = (int)rd(sb + 9, 1); sl = (int)rd(sb + 10, 1);
            check_sizes(sb);
            f.base_address = ra(sb + 12);
            uint64_t root = ra(sb + 12 + 3 * so);
            check_sum(sb, sb + 12 + 4 * so);
            return abs_addr(root);
        }
        bad(f.path, sb + 8, "superblock version " + std::to_string(ver) + " not known");
    }

    // ------------------------------------------------------ object headers
    void scan_v2(uint64_t q, uint64_t end, int oflags, std::vector<Msg>& out, std::vector<std::pair<uint64_t, uint64_t>>& cont) {
        const uint64_t hdr = 4 + ((oflags & 0x04) ? 2 : 0);
        while (q + hdr <= end) {
            Msg m;
            m.type = (int)rd(q, 1); m.size = rd(q + 1, 2); m.flags = (int)rd(q + 3, 1); m.off = q + hdr;
            if (m.off + m.size > end) bad(f.path, q, "object-header message runs past its chunk");
            if (m.type == 0x10) cont.push_back({abs_addr(ra(m.off)), rd(m.off + so, sl)});
            else if (m.type != 0) out.push_back(m);
            q = m.off + m.size;
        }
    }
    std::vector<Msg> messages(uint64_t at) {
        std::vector<Msg> out;
        std::vector<std::pair<uint64_t, uint64_t>> cont;
        if (is(at, "OHDR")) {
            if (rd(at + 4, 1) != 2) bad(f.path, at, "OHDR version is not 2");
            int oflags = (int)rd(at + 5, 1);
            uint64_t p = at + 6;
            if (oflags & 0x20) p += 16;
            if (oflags & 0x10) p += 4;
            int nb = 1 << (oflags & 3);
            uint64_t csz = rd(p, nb);
            p += nb;
            need(p, csz + 4);
            scan_v2(p, p + csz, oflags, out, cont);
            check_sum(at, p + csz);
            for (size_t i = 0; i < cont.size(); ++i) {
                uint64_t c = cont[i].first, len = cont[i].second;
                if (len < 8 || !is(c, "OCHK")) bad(f.path, c, "object-header continuation without OCHK signature");
                need(c, len);
                scan_v2(c + 4, c + len - 4, oflags, out, cont);
                check_sum(c, c + len - 4);
                if (cont.size() > 4096) bad(f.path, c, "object-header continuation loop");
            }
            return out;
        }
        if (rd(at, 1) != 1) bad(f.path, at, "neither a version-1 nor a version-2 object header");
        uint64_t nmsg = rd(at + 2, 2), hsize = rd(at + 8, 4);
        cont.push_back({at + 16, hsize});
        uint64_t seen = 0;
        for (size_t i = 0; i < cont.size() && seen < nmsg; ++i) {
            uint64_t q = cont[i].first, end = q + cont[i].second;
            need(q, cont[i].second);
            while (q + 8 <= end && seen < nmsg) {
                Msg m;
                m.type = (int)rd(q, 2); m.size = rd(q + 2, 2); m.flags = (int)rd(q + 4, 1); m.off = q + 8;
                if (m.off + m.size > end) bad(f.path, q, "object-header message runs past its block");
                ++seen;
                if (m.type == 0x10) cont.push_back({abs_addr(ra(m.off)), rd(m.off + so, sl)});
                else if (m.type != 0) out.push_back(m);
                q = m.off + m.size;
            }
            if (cont.size() > 4096) bad(f.path, at, "object-header continuation loop");
        }
        return out;
    }

    // ------------------------------------------------------------ datatype
    DType datatype(uint64_t q, uint64_t end) {
        DType t;
        if (q + 8 > end) bad(f.path, q, "truncated datatype message");
        int cv = (int)rd(q, 1), cls = cv & 15, ver = cv >> 4;
        uint32_t bits = (uint32_t)rd(q + 1, 3);
        t.size = (int)rd(q + 4, 4);
        if ((cls == 0 || cls == 1) && (bits & 1)) unsupp(f.path, "big-endian numeric data");
        if (cls == 0) {
            t.kind = (t.size == 1 || t.size == 2 || t.size == 4 || t.size == 8) ? EK_INT : EK_OTHER;
            t.is_signed = (bits & 8) != 0; t.len = 12;
        }
        else if (cls == 1) {
            t.len = 20;
            t.kind = t.size == 8 ? EK_F64 : t.size == 4 ? EK_F32 : EK_OTHER;
        } else if (cls == 3) { t.kind = t.size > 0 ? EK_STRING : EK_OTHER; t.len = 8; }
        else if (cls == 7) { t.len = 8; }
        else if (cls == 6) {
            int nmemb = bits & 0xffff;
            uint64_t p = q + 8;
            int obytes = 1; while (obytes < 4 && (uint64_t)t.size >= (1ull << (8 * obytes))) ++obytes;
            std::vector<std::pair<uint64_t, DType>> memb;
            for (int i = 0; i < nmemb; ++i) {
                uint64_t n = 0;
                while (p + n < end && buf[p + n]) ++n;
                if (p + n >= end) bad(f.path, p, "unterminated compound member name");
                p += (ver >= 3) ? n + 1 : ((n + 8) / 8) * 8;
                uint64_t off;
                if (ver >= 3) { off = rd(p, obytes); p += obytes; }
                else { off = rd(p, 4); p += 4; if (ver == 1) p += 28; }
                DType m = datatype(p, end);
                if (!m.len) return t;                       // member we cannot step over: EK_OTHER
                p += m.len;
                memb.push_back({off, m});
            }
            t.len = p - q;
            if (nmemb == 2 && memb[0].first == 0 && memb[0].second.kind == memb[1].second.kind &&
                memb[1].first == (uint64_t)memb[0].second.size && t.size == 2 * memb[0].second.size) {
                if (memb[0].second.kind == EK_F64) t.kind = EK_C64;
                if (memb[0].second.kind == EK_F32) t.kind = EK_C32;
            }
        }
        return t;
    }

    DType resolve_datatype(const Msg& m, bool& committed, int depth = 0) {
        committed = false;
        if (!(m.flags & 0x02)) return datatype(m.off, m.off + m.size);
        // a committed datatype whose object header holds another shared datatype message may point back at itself
        if (depth > 8) bad(f.path, m.off, "committed datatype references nest deeper than 8 levels (cycle?)");
        int ver = (int)rd(m.off, 1), typ = (int)rd(m.off + 1, 1);
        uint64_t a;
        if (ver == 1) a = ra(m.off + 8);
        else if (ver == 2 || (ver == 3 && typ == 2)) a = ra(m.off + 2);
        else unsupp(f.path, "datatype stored in the shared-message heap");
        committed = true;
        for (const Msg& c : messages(abs_addr(a)))
            if (c.type == 0x03) { bool dummy; return resolve_datatype(c, dummy, depth + 1); }
        bad(f.path, abs_addr(a), "committed datatype object holds no datatype message");
    }

    // ------------------------------------------------------------ datasets
    void dataset(const std::string& name, const std::vector<Msg>& ms) {
        Dataset d;
        d.name = name;
        const Msg *sp = nullptr, *dt = nullptr, *lay = nullptr;
        for (const Msg& m : ms) {
            if (m.type == 0x01) sp = &m;
            if (m.type == 0x03) dt = &m;
            if (m.type == 0x08) lay = &m;
            if (m.type == 0x0B) unsupp(f.path, "dataset '" + name + "' uses a filter pipeline (compression); write the file with compress=false");
        }
        // dataspace
        int sver = (int)rd(sp->off, 1), rank = (int)rd(sp->off + 1, 1);
        uint64_t dp = sp->off + (sver == 1 ? 8 : 4);
        bool null_space = sver >= 2 && rd(sp->off + 3, 1) == 2;
        if (rank > 32) bad(f.path, sp->off, "dataset '" + name + "': rank above 32");
        std::vector<int64_t> h5dims(rank);
        for (int i = 0; i < rank; ++i) h5dims[i] = (int64_t)rd(dp + (uint64_t)i * sl, sl);
        d.dims.assign(h5dims.rbegin(), h5dims.rend());
        if (null_space) d.dims = {0};
        // datatype
        DType t = resolve_datatype(*dt, d.committed_type);
        d.kind = t.kind; d.elem_size = t.size; d.is_signed = t.is_signed;
        if (d.kind == EK_OTHER) { f.datasets.push_back(d); return; }
        // extents come from the file: bound them by what the file can hold before any size arithmetic
        uint64_t count = 1;
        for (int64_t e : d.dims) {
            if (e < 0 || (uint64_t)e > buf.size() || (e > 0 && count > buf.size() / (uint64_t)e + 1)) bad(f.path, sp->off, "dataset '" + name + "': extents exceed the file size");
            count *= (uint64_t)e;
        }
        if (d.elem_size <= 0 || count > buf.size() / (uint64_t)d.elem_size + 1)
            bad(f.path, sp->off, "dataset '" + name + "': extents exceed the file size");
        uint64_t nbytes = count * (uint64_t)d.elem_size;
        // layout
        int lver = (int)rd(lay->off, 1);
        int cls;
        uint64_t addr = kUndef, csize = 0, cdata = 0;
        if (lver == 1 || lver == 2) {
            int ndim = (int)rd(lay->off + 1, 1);
            cls = (int)rd(lay->off + 2, 1);
            uint64_t p = lay->off + 8;
            if (cls != 0) { addr = ra(p); p += so; }
            p += 4ull * ndim;
            if (cls == 0) { csize = rd(p, 4); cdata = p + 4; }
        } else if (lver == 3 || lver == 4) {
            cls = (int)rd(lay->off + 1, 1);
            if (cls == 0) { csize = rd(lay->off + 2, 2); cdata = lay->off + 4; }
            else if (cls == 1) addr = ra(lay->off + 2);
        } else bad(f.path, lay->off, "data-layout message version " + std::to_string(lver));
        if (cls == 2) unsupp(f.path, "dataset '" + name + "' is chunked; only contiguous and compact layouts are read");
        if (cls > 2) unsupp(f.path, "dataset '" + name + "' uses virtual storage");
        d.raw.assign(nbytes, 0);
        if (cls == 0) {
            if (csize < nbytes) bad(f.path, lay->off, "compact data smaller than the dataspace");
            need(cdata, nbytes);
            memcpy(d.raw.data(), &buf[cdata], nbytes);
        } else if (addr != kUndef && nbytes) {               // undefined address = never written = fill value (zeros)
            d.data_offset = abs_addr(addr);
            need(d.data_offset, nbytes);
            memcpy(d.raw.data(), &buf[d.data_offset], nbytes);
        }
        f.datasets.push_back(std::move(d));
    }

    // -------------------------------------------------------------- groups
    void child(const std::string& prefix, const std::string& name, uint64_t at, int depth) {
        if (at == kUndef) return;
        if (prefix.empty() && !name.empty() && name[0] == '_') return;     // JLD2's `_types` bookkeeping group
        object(prefix + name, at, depth + 1);
    }
    void object(const std::string& name, uint64_t at, int depth) {
        if (depth > 16 || !visited.insert(at).second) return;
        std::vector<Msg> ms = messages(at);
        bool has_sp = false, has_dt = false, has_lay = false, groupish = false;
        for (const Msg& m : ms) {
            has_sp |= m.type == 0x01; has_dt |= m.type == 0x03; has_lay |= m.type == 0x08;
            groupish |= m.type == 0x11 || m.type == 0x02 || m.type == 0x06 || m.type == 0x0A;
        }
        if (has_sp && has_dt && has_lay) { dataset(name, ms); return; }
        if (!groupish) return;                               // a committed datatype or something we do not need
        std::string prefix = name.empty() ? "" : name + "/";
        for (const Msg& m : ms) {
            if (m.type == 0x06) {                            // link message
                int lf = (int)rd(m.off + 1, 1);
                uint64_t p = m.off + 2;
                int ltype = 0;
                if (lf & 0x08) { ltype = (int)rd(p, 1); p += 1; }
                if (lf & 0x04) p += 8;
                if (lf & 0x10) p += 1;
                int nb = 1 << (lf & 3);
                uint64_t nlen = rd(p, nb);
                p += nb;
                need(p, nlen);
                std::string ln((const char*)&buf[p], nlen);
                p += nlen;
                if (ltype == 0) child(prefix, ln, abs_addr(ra(p)), depth);
            } else if (m.type == 0x02) {                     // link info: dense storage when a fractal heap exists
                int lf = (int)rd(m.off + 1, 1);
                uint64_t p = m.off + 2 + ((lf & 1) ? 8 : 0);
                if (ra(p) != kUndef) unsupp(f.path, "group '" + name + "' stores its links densely (fractal heap)");
            } else if (m.type == 0x11) {                     // symbol table: v1 B-tree + local heap
                uint64_t bt = abs_addr(ra(m.off)), hp = abs_addr(ra(m.off + so));
                if (!is(hp, "HEAP")) bad(f.path, hp, "local heap signature missing");
                uint64_t hdata = abs_addr(ra(hp + 8 + 2 * sl));
                btree(prefix, bt, hdata, depth, 0);
            }
        }
    }
    void btree(const std::string& prefix, uint64_t at, uint64_t hdata, int depth, int rec) {
        if (at == kUndef) return;
        if (rec > 32) bad(f.path, at, "B-tree deeper than 32 levels");
        if (is(at, "SNOD")) {
            uint64_t n = rd(at + 6, 2), p = at + 8;
            for (uint64_t i = 0; i < n; ++i, p += 2 * so + 24) {
                uint64_t noff = rd(p, so), oh = ra(p + so);
                need(hdata + noff, 1);
                std::string nm((const char*)&buf[hdata + noff], strnlen((const char*)&buf[hdata + noff], buf.size() - hdata - noff));
                child(prefix, nm, abs_addr(oh), depth);
            }
            return;
        }
        if (!is(at, "TREE")) bad(f.path, at, "expected a TREE or SNOD node");
        if (rd(at + 4, 1) != 0) bad(f.path, at, "B-tree node is not a group node");
        uint64_t n = rd(at + 6, 2), p = at + 8 + 2 * so;
        for (uint64_t i = 0; i < n; ++i) {
            p += sl;                                         // key i
            btree(prefix, abs_addr(ra(p)), hdata, depth, rec + 1);
            p += so;
        }
    }
};

// ------------------------------------------------------------------ writer
struct Out {
    std::vector<uint8_t> b;
    void u(uint64_t v, int n) { for (int i = 0; i < n; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
    void bytes(const void* p, size_t n) { const uint8_t* q = (const uint8_t*)p; b.insert(b.end(), q, q + n); }
    void str(const std::string& s) { bytes(s.data(), s.size()); }
    void align8() { while (b.size() % 8) b.push_back(0); }
};

struct WMsg { int type; int flags; std::vector<uint8_t> data; };

const uint64_t kBase = 512;

// appends a version-2 object header, returns its address relative to the base
uint64_t put_header(Out& o, const std::vector<WMsg>& ms) {
    o.align8();
    uint64_t at = o.b.size();
    uint64_t body = 0;
    for (const WMsg& m : ms) body += 4 + m.data.size();
    o.str("OHDR"); o.u(2, 1); o.u(0x02, 1);                  // flags: 4-byte chunk size, no times, no creation order
    o.u(body, 4);
    for (const WMsg& m : ms) { o.u(m.type, 1); o.u(m.data.size(), 2); o.u(m.flags, 1); o.bytes(m.data.data(), m.data.size()); }
    o.u(lookup3(&o.b[at], o.b.size() - at), 4);
    return at - kBase;
}

std::vector<uint8_t> dt_float(int size) {
    Out o;
    o.u(0x11, 1); o.u(0x20, 1); o.u(size * 8 - 1, 1); o.u(0, 1); o.u(size, 4);
    o.u(0, 2); o.u(size * 8, 2);
    if (size == 8) { o.u(52, 1); o.u(11, 1); o.u(0, 1); o.u(52, 1); o.u(1023, 4); }
    else           { o.u(23, 1); o.u(8, 1);  o.u(0, 1); o.u(23, 1); o.u(127, 4); }
    return o.b;
}
std::vector<uint8_t> dt_complex(int fsize) {
    Out o;
    o.u(0x36, 1); o.u(2, 2); o.u(0, 1); o.u(2 * fsize, 4);   // compound, version 3, two members
    std::vector<uint8_t> m = dt_float(fsize);
    o.str("re"); o.u(0, 1); o.u(0, 1); o.bytes(m.data(), m.size());
    o.str("im"); o.u(0, 1); o.u(fsize, 1); o.bytes(m.data(), m.size());
    return o.b;
}
std::vector<uint8_t> dt_of(const WriteArray& a) {
    Out o;
    switch (a.kind) {
        case EK_C64: return dt_complex(8);
        case EK_C32: return dt_complex(4);
        case EK_F64: return dt_float(8);
        case EK_F32: return dt_float(4);
        case EK_INT: o.u(0x10, 1); o.u(0x08, 1); o.u(0, 2); o.u(a.elem_size, 4); o.u(0, 2); o.u(a.elem_size * 8, 2); return o.b;
        case EK_STRING: o.u(0x13, 1); o.u(0x01, 1); o.u(0, 2); o.u(a.elem_size, 4); return o.b;   // null-padded ASCII
    }
    throw Error(QXB_ERR_ARG, "jld2 writer: element kind " + std::to_string(a.kind) + " cannot be written");
}
WMsg link_msg(const std::string& name, uint64_t rel) {
    Out o;
    int nb = name.size() < 256 ? 0 : 1;
    o.u(1, 1); o.u(0x10 | nb, 1); o.u(1, 1);                 // version, flags (charset present), UTF-8
    o.u(name.size(), 1 << nb); o.str(name); o.u(rel, 8);
    return {0x06, 0, o.b};
}
uint64_t put_group(Out& o, const std::vector<std::pair<std::string, uint64_t>>& links) {
    std::vector<WMsg> ms;
    Out li; li.u(0, 2); li.u(kUndef, 8); li.u(kUndef, 8);
    ms.push_back({0x02, 0, li.b});
    Out gi; gi.u(0, 2);
    ms.push_back({0x0A, 0, gi.b});
    for (auto& l : links) ms.push_back(link_msg(l.first, l.second));
    return put_header(o, ms);
}

}  // namespace

File read_file(const std::string& path) {
    File f;
    f.path = path;
    Reader r(f);
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) throw Error(QXB_ERR_ARG, "cannot open '" + path + "'");
    fseek(fp, 0, SEEK_END);
    long n = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    r.buf.resize(n > 0 ? (size_t)n : 0);
    size_t got = r.buf.empty() ? 0 : fread(r.buf.data(), 1, r.buf.size(), fp);
    fclose(fp);
    if (got != r.buf.size()) throw Error(QXB_ERR_ARG, "short read of '" + path + "'");
    uint64_t root = r.superblock();
    r.object("", root, 0);
    return f;
}

std::vector<std::complex<double>> as_c64(const Dataset& d) {
    const int64_t n = d.count();
    std::vector<std::complex<double>> v((size_t)n);
    const uint8_t* p = d.raw.data();
    auto ld = [&](const uint8_t* q, int size) -> double {
        if (size == 8) { double x; memcpy(&x, q, 8); return x; }
        float x; memcpy(&x, q, 4); return x;
    };
    switch (d.kind) {
        case EK_C64: for (int64_t i = 0; i < n; ++i) v[i] = {ld(p + 16 * i, 8), ld(p + 16 * i + 8, 8)}; break;
        case EK_C32: for (int64_t i = 0; i < n; ++i) v[i] = {ld(p + 8 * i, 4), ld(p + 8 * i + 4, 4)}; break;
        case EK_F64: for (int64_t i = 0; i < n; ++i) v[i] = ld(p + 8 * i, 8); break;
        case EK_F32: for (int64_t i = 0; i < n; ++i) v[i] = ld(p + 4 * i, 4); break;
        case EK_INT:
            for (int64_t i = 0; i < n; ++i) {
                uint64_t u = 0;
                if (d.elem_size > 8) throw Error(QXB_ERR_UNSUPP, "dataset '" + d.name + "': integers wider than 64 bits");
                memcpy(&u, p + (size_t)d.elem_size * i, d.elem_size);
                if (d.is_signed && d.elem_size < 8 && (u >> (8 * d.elem_size - 1))) u |= ~0ull << (8 * d.elem_size);
                v[i] = d.is_signed ? (double)(int64_t)u : (double)u;
            }
            break;
        default:
            throw Error(QXB_ERR_UNSUPP, "dataset '" + d.name + "' is not numeric");
    }
    return v;
}

void write_file(const std::string& path, const std::vector<WriteArray>& arrays, bool commit_types) {
    Out o;
    std::string head = "HDF5-based Julia Data Format, version 0.1.1 (written by qxb200, not by JLD2.jl)";
    o.str(head);
    o.b.resize(kBase, 0);
    o.b.resize(kBase + 48, 0);                               // superblock, filled in last
    std::vector<std::pair<std::string, uint64_t>> links, type_links;
    uint64_t committed[2] = {kUndef, kUndef};                // EK_C64, EK_C32
    if (commit_types) {
        for (int k = 0; k < 2; ++k) {
            bool used = false;
            for (const WriteArray& a : arrays) used |= a.kind == k;
            if (!used) continue;
            committed[k] = put_header(o, {{0x03, 0, dt_complex(k == EK_C64 ? 8 : 4)}});
            char nm[16]; snprintf(nm, sizeof nm, "%08d", (int)type_links.size() + 1);
            type_links.push_back({nm, committed[k]});
        }
    }
    for (const WriteArray& a : arrays) {
        if (a.name.empty() || a.name.find('/') != std::string::npos)
            throw Error(QXB_ERR_ARG, "jld2 writer: dataset names must be non-empty and flat, got '" + a.name + "'");
        int64_t count = 1;
        for (int64_t d : a.dims) { if (d < 0) throw Error(QXB_ERR_ARG, "jld2 writer: negative extent"); count *= d; }
        uint64_t nbytes = (uint64_t)count * (uint64_t)a.elem_size;
        o.align8();
        uint64_t data_rel = nbytes ? o.b.size() - kBase : kUndef;
        if (nbytes) o.bytes(a.data, nbytes);
        std::vector<WMsg> ms;
        Out sp;
        sp.u(2, 1); sp.u(a.dims.size(), 1); sp.u(0, 1); sp.u(a.dims.empty() ? 0 : 1, 1);
        for (size_t i = a.dims.size(); i-- > 0;) sp.u((uint64_t)a.dims[i], 8);
        ms.push_back({0x01, 0, sp.b});
        if (commit_types && (a.kind == EK_C64 || a.kind == EK_C32)) {
            Out sh; sh.u(3, 1); sh.u(2, 1); sh.u(committed[a.kind], 8);
            ms.push_back({0x03, 0x02, sh.b});
        } else ms.push_back({0x03, 0x01, dt_of(a)});         // flag 1: constant message
        Out lay; lay.u(3, 1); lay.u(1, 1); lay.u(data_rel, 8); lay.u(nbytes, 8);
        ms.push_back({0x08, 0, lay.b});
        links.push_back({a.name, put_header(o, ms)});
    }
    if (!type_links.empty()) links.push_back({"_types", put_group(o, type_links)});
    uint64_t root = put_group(o, links);
    o.align8();
    Out sb;
    sb.bytes(kSig, 8); sb.u(2, 1); sb.u(8, 1); sb.u(8, 1); sb.u(0, 1);
    sb.u(kBase, 8); sb.u(kUndef, 8); sb.u(o.b.size() - kBase, 8); sb.u(root, 8);
    sb.u(lookup3(sb.b.data(), sb.b.size()), 4);
    memcpy(&o.b[kBase], sb.b.data(), sb.b.size());
    FILE* fp = fopen(path.c_str(), "wb");
    if (!fp) throw Error(QXB_ERR_ARG, "cannot create '" + path + "'");
    size_t put = fwrite(o.b.data(), 1, o.b.size(), fp);
    if (fclose(fp) != 0 || put != o.b.size()) throw Error(QXB_ERR_ARG, "short write of '" + path + "'");
}

}  // namespace jld2
}  // namespace qxb
