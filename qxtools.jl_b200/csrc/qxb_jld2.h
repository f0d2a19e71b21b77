// qxb200 -- native reader/writer for the data files of a simulation triple.
//
// The reference stores its leaf tensors with JLD2.jl, one dataset per label holding the N-d
// ComplexF64 array (/root/reference/src/compute_graph/tensor_cache.jl:90-106), and the runner writes
// its results to a `.jld2` file as well (/root/reference/docs/src/distributed.md:30-33).  JLD2 files are
// HDF5 containers: a 512-byte text header, a version-2 superblock with base address 512, version-2
// object headers, link messages in the group header, contiguous (or compact) dataset layout, and
// Complex{Float64} as a compound datatype {re: f64 @0, im: f64 @8}, usually *committed* under
// `_types/` and referenced from the dataset through a shared-message.  Julia arrays are column-major
// and JLD2 writes the dataspace extents reversed, so the raw bytes are the Julia memory order.
//
// This file implements the HDF5 subset needed for that (plus version-0/1 superblocks, version-1 object
// headers and symbol-table groups so files written by the HDF5 C library can be read too).  Not
// supported (reported, never guessed): chunked / filtered (compressed) datasets, dense link storage,
// virtual / external storage, big-endian files.
#pragma once
#include <complex>
#include <cstdint>
#include <string>
#include <vector>

namespace qxb {
namespace jld2 {

enum ElemKind {
    EK_C64 = 0,      // compound of two little-endian float64 (re, im)
    EK_C32 = 1,      // compound of two float32
    EK_F64 = 2,
    EK_F32 = 3,
    EK_INT = 4,      // fixed-point, elem_size bytes, is_signed
    EK_STRING = 5,   // fixed-length string of elem_size bytes
    EK_OTHER = 6     // present in the file, not convertible (vlen strings, references, Julia structs ...)
};

struct Dataset {
    std::string name;                 // path below the root group, '/'-separated
    std::vector<int64_t> dims;        // Julia (column-major) order = HDF5 extents reversed
    int kind = EK_OTHER;
    int elem_size = 0;
    bool is_signed = true;
    bool committed_type = false;      // the datatype was a shared message pointing at a committed datatype
    uint64_t data_offset = 0;         // absolute file offset of the raw data (contiguous) / 0 for compact
    std::vector<uint8_t> raw;         // the elements, file byte order (little endian)
    int64_t count() const { int64_t n = 1; for (int64_t d : dims) n *= d; return n; }
};

struct File {
    std::string path;
    std::string header;               // the text in front of the superblock (JLD2: "HDF5-based Julia Data Format, ...")
    int superblock_version = -1;
    uint64_t base_address = 0;
    int checksum_failures = 0;        // version-2 structures whose lookup3 checksum did not match
    std::vector<Dataset> datasets;    // every dataset reachable from the root group (groups starting with '_' skipped)
    const Dataset* find(const std::string& name) const;
};

// Throws qxb::Error(QXB_ERR_ARG) with the offending offset on malformed input, QXB_ERR_UNSUPP for the
// unsupported features listed above.
File read_file(const std::string& path);
// Elements of a numeric dataset as ComplexF64 (real types get a zero imaginary part).
std::vector<std::complex<double>> as_c64(const Dataset& d);

struct WriteArray {
    std::string name;
    std::vector<int64_t> dims;        // Julia order
    int kind = EK_C64;                // EK_C64 / EK_C32 / EK_F64 / EK_F32 / EK_INT (int64) / EK_STRING
    int elem_size = 16;
    const void* data = nullptr;       // count() * elem_size bytes
};
// Writes a JLD2-layout container (512-byte header, superblock v2, base address 512, OHDR v2, link
// messages, contiguous layout).  commit_types = true stores the complex datatypes once under `_types/`
// and references them with shared messages, as JLD2.jl does; false writes them inline (readable by
// every HDF5 tool, and by JLD2.jl as a NamedTuple{(:re, :im)}).
void write_file(const std::string& path, const std::vector<WriteArray>& arrays, bool commit_types);

// Bob Jenkins' lookup3 hashlittle(), the checksum of every version-2 HDF5 metadata structure.
uint32_t lookup3(const uint8_t* key, size_t length, uint32_t initval = 0);

}  // namespace jld2
}  // namespace qxb
