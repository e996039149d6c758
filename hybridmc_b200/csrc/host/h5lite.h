// h5lite — a dependency-free writer (and a reader for what it writes) of HDF5 files, for
// the outputs the reference produces through libhdf5's C++ API (src/writer.h,
// src/config.cc:18-60, src/snapshot.cc:14-131, src/main.cc:383-409).  The build image has no
// libhdf5, so the file format is produced directly from the HDF5 File Format
// Specification, version 1.1 subset that every libhdf5 (1.6 ... 1.14) and h5py read:
//
//   superblock v0 (8-byte offsets and lengths, group leaf K 4, internal K 16)
//   groups        = object header v1 + symbol-table message (0x11) -> v1 B-tree ("TREE",
//                   node type 0) -> symbol-table nodes ("SNOD") + local heap ("HEAP")
//   datasets      = object header v1 with dataspace v1 (0x01), datatype v1 (0x03),
//                   fill value v1 (0x05), layout v3 CONTIGUOUS (0x08)
//   attributes    = attribute message v1 (0x0c) in the root group's object header
//   strings       = variable-length string (datatype class 9) through a global heap
//                   collection ("GCOL"), as snapshot.cc:41-47 writes the RNG state
//
// Differences from what the reference writes, all invisible to H5Cpp/h5py readers:
// datasets are contiguous with fixed extents instead of chunked/extensible (+deflate on
// pos/vel) because the file is written once at the end of the run.
//
// Byte-level conventions were checked against a libhdf5-written file
// (tests/test_h5lite.py: scipy's testhdf5_7.4_GLNX86.mat): local-heap free-list
// terminator 1, B-tree key layout, 16-byte v1 object-header prefix, message padding to 8.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unistd.h>
#include <vector>
#ifdef H5LITE_WITH_ZLIB
#include <zlib.h>
#endif

namespace h5lite {

enum Type : uint8_t { F64 = 0, F32 = 1, U64 = 2, U32 = 3, VSTR = 4 };

inline size_t type_size(uint8_t t) {
  switch (t) {
  case F64:
  case U64: return 8;
  case F32:
  case U32: return 4;
  default: return 1;
  }
}

struct Item {
  uint8_t dtype = F64;
  std::vector<uint64_t> dims; // empty = scalar
  std::string bytes;          // raw little-endian elements; VSTR: the characters
};

namespace detail {

constexpr uint64_t UNDEF = ~uint64_t(0);
constexpr unsigned LEAF_K = 4, INTERNAL_K = 16; // superblock defaults of libhdf5
constexpr size_t SNOD_SIZE = 8 + 2 * LEAF_K * 40;
constexpr size_t TREE_SIZE = 24 + 2 * INTERNAL_K * 8 + (2 * INTERNAL_K + 1) * 8;

inline size_t pad8(size_t n) { return (n + 7) & ~size_t(7); }

struct Buf {
  std::string s;
  size_t size() const { return s.size(); }
  void u8(uint8_t v) { s.push_back(char(v)); }
  void u16(uint16_t v) { for (int i = 0; i < 2; i++) s.push_back(char(v >> (8 * i))); }
  void u32(uint32_t v) { for (int i = 0; i < 4; i++) s.push_back(char(v >> (8 * i))); }
  void u64(uint64_t v) { for (int i = 0; i < 8; i++) s.push_back(char(v >> (8 * i))); }
  void raw(const void *p, size_t n) { s.append(static_cast<const char *>(p), n); }
  void zeros(size_t n) { s.append(n, '\0'); }
  void align8() { zeros(pad8(s.size()) - s.size()); }
  void patch64(size_t at, uint64_t v) { for (int i = 0; i < 8; i++) s[at + i] = char(v >> (8 * i)); }
};

// datatype message bodies (format spec IV.A.2.d)
inline std::string datatype_msg(uint8_t t) {
  Buf b;
  switch (t) {
  case F64: // class 1 v1, LE, implied-msb mantissa, sign bit 63; exponent 52/11, bias 1023
    b.u8(0x11); b.u8(0x20); b.u8(63); b.u8(0); b.u32(8);
    b.u16(0); b.u16(64); b.u8(52); b.u8(11); b.u8(0); b.u8(52); b.u32(1023);
    break;
  case F32:
    b.u8(0x11); b.u8(0x20); b.u8(31); b.u8(0); b.u32(4);
    b.u16(0); b.u16(32); b.u8(23); b.u8(8); b.u8(0); b.u8(23); b.u32(127);
    break;
  case U64: // class 0 v1, LE, unsigned
    b.u8(0x10); b.u8(0); b.u8(0); b.u8(0); b.u32(8); b.u16(0); b.u16(64);
    break;
  case U32:
    b.u8(0x10); b.u8(0); b.u8(0); b.u8(0); b.u32(4); b.u16(0); b.u16(32);
    break;
  case VSTR: // class 9 v1: type 1 = string, null-terminated, ASCII; 16-byte element; base u8
    b.u8(0x19); b.u8(0x01); b.u8(0); b.u8(0); b.u32(16);
    b.u8(0x10); b.u8(0); b.u8(0); b.u8(0); b.u32(1); b.u16(0); b.u16(8);
    break;
  default: throw std::runtime_error("h5lite: unknown type");
  }
  return b.s;
}

inline std::string dataspace_msg(const std::vector<uint64_t> &dims) {
  Buf b; // v1: version, rank, flags (no max dims: max = current), 5 reserved bytes, dims
  b.u8(1); b.u8(uint8_t(dims.size())); b.u8(0); b.u8(0); b.u32(0);
  for (uint64_t d : dims) b.u64(d);
  return b.s;
}

inline void message(Buf &b, uint16_t type, const std::string &body, uint8_t flags = 0) {
  const size_t n = pad8(body.size());
  if (n > 0xfff8) throw std::runtime_error("h5lite: object header message too large");
  b.u16(type); b.u16(uint16_t(n)); b.u8(flags); b.u8(0); b.u16(0);
  b.raw(body.data(), body.size());
  b.zeros(n - body.size());
}

// object header v1: 12-byte prefix padded to 16, then the messages
inline uint64_t object_header(Buf &f, unsigned nmsgs, const Buf &msgs) {
  f.align8();
  const uint64_t at = f.size();
  f.u8(1); f.u8(0); f.u16(uint16_t(nmsgs)); f.u32(1); f.u32(uint32_t(msgs.size())); f.u32(0);
  f.raw(msgs.s.data(), msgs.size());
  return at;
}

} // namespace detail

class Writer {
public:
  // path: "name" or "group/name" (groups are created on demand); replaces an existing item
  void dataset(const std::string &path, uint8_t dtype, const std::vector<uint64_t> &dims,
               const void *data, size_t nbytes) {
    Group *g = &root_;
    size_t start = 0, slash;
    while ((slash = path.find('/', start)) != std::string::npos) {
      auto &sub = g->groups[path.substr(start, slash - start)];
      if (!sub) sub.reset(new Group);
      g = sub.get();
      start = slash + 1;
    }
    Item &it = g->datasets[path.substr(start)];
    it.dtype = dtype;
    it.dims = dims;
    it.bytes.assign(static_cast<const char *>(data), nbytes);
    check(it, path);
  }
  // attribute of the root group (H5File::createAttribute)
  void attribute(const std::string &name, uint8_t dtype, const std::vector<uint64_t> &dims,
                 const void *data, size_t nbytes) {
    Item &it = attrs_[name];
    it.dtype = dtype;
    it.dims = dims;
    it.bytes.assign(static_cast<const char *>(data), nbytes);
    check(it, name);
  }

  std::string serialize() const {
    using namespace detail;
    Buf f;
    // superblock v0 (spec II.A); EOF address and root entry patched below
    f.raw("\x89HDF\r\n\x1a\n", 8);
    f.u8(0); f.u8(0); f.u8(0); f.u8(0); f.u8(0); f.u8(8); f.u8(8); f.u8(0);
    f.u16(LEAF_K); f.u16(INTERNAL_K); f.u32(0);
    f.u64(0);     // base address
    f.u64(UNDEF); // free-space info
    const size_t eof_at = f.size();
    f.u64(0);     // end of file
    f.u64(UNDEF); // driver info
    const size_t root_entry_at = f.size();
    f.zeros(40);
    Placed root = write_group(f, root_, &attrs_);
    f.align8();
    f.patch64(eof_at, f.size());
    f.patch64(root_entry_at + 0, 0);
    f.patch64(root_entry_at + 8, root.header);
    f.s[root_entry_at + 16] = 1; // cache type 1: scratch pad = B-tree and heap addresses
    f.patch64(root_entry_at + 24, root.btree);
    f.patch64(root_entry_at + 32, root.heap);
    return f.s;
  }

  bool write(const std::string &path, bool do_fsync) const {
    const std::string bytes = serialize();
    FILE *fp = std::fopen(path.c_str(), "wb");
    if (!fp) return false;
    bool ok = std::fwrite(bytes.data(), 1, bytes.size(), fp) == bytes.size();
    ok = std::fflush(fp) == 0 && ok;
    if (do_fsync) ok = fsync(fileno(fp)) == 0 && ok;
    return std::fclose(fp) == 0 && ok;
  }

private:
  struct Group {
    std::map<std::string, std::unique_ptr<Group>> groups; // std::map: strcmp order, as SNODs need
    std::map<std::string, Item> datasets;
  };
  struct Placed {
    uint64_t header = 0, btree = 0, heap = 0;
    bool is_group = false;
  };
  Group root_;
  std::map<std::string, Item> attrs_;

  static void check(const Item &it, const std::string &name) {
    if (it.dtype == VSTR) {
      if (!it.dims.empty()) throw std::runtime_error("h5lite: " + name + ": strings are scalar");
      return;
    }
    uint64_t n = 1;
    for (uint64_t d : it.dims) n *= d;
    if (n * type_size(it.dtype) != it.bytes.size())
      throw std::runtime_error("h5lite: " + name + ": data size does not match the extent");
  }

  static uint64_t write_dataset(detail::Buf &f, const Item &it) {
    using namespace detail;
    uint64_t addr = UNDEF, size = 0;
    if (it.dtype == VSTR) {
      // global heap collection (spec III.E): one object (index 1) + the free-space object 0
      f.align8();
      const uint64_t gcol = f.size();
      const size_t len = it.bytes.size();
      size_t total = 16 + 16 + pad8(len) + 16;
      if (total < 4096) total = 4096;
      f.raw("GCOL", 4); f.u8(1); f.u8(0); f.u16(0); f.u64(total);
      f.u16(1); f.u16(1); f.u32(0); f.u64(len); // index 1, reference count 1
      f.raw(it.bytes.data(), len); f.zeros(pad8(len) - len);
      const size_t free_bytes = gcol + total - f.size();
      f.u16(0); f.u16(0); f.u32(0); f.u64(free_bytes); // object 0: free space incl. this header
      f.zeros(free_bytes - 16);
      addr = f.size();
      f.u32(uint32_t(len)); f.u64(gcol); f.u32(1); // vlen element: length, collection, index
      size = 16;
    } else if (!it.bytes.empty()) {
      f.align8();
      addr = f.size();
      size = it.bytes.size();
      f.raw(it.bytes.data(), it.bytes.size());
    }
    Buf m;
    message(m, 0x0001, dataspace_msg(it.dims));
    message(m, 0x0003, datatype_msg(it.dtype), 1);
    { // fill value: allocate late, write if set, default (zero-length) value defined — the
      // very bytes libhdf5 wrote for the contiguous dataset of the fixture (version 1)
      Buf b; b.u8(1); b.u8(2); b.u8(2); b.u8(1); b.u32(0);
      message(m, 0x0005, b.s, 1);
    }
    { // layout v3, class 1 = contiguous
      Buf b; b.u8(3); b.u8(1); b.u64(addr); b.u64(size);
      message(m, 0x0008, b.s);
    }
    return object_header(f, 4, m);
  }

  static Placed write_group(detail::Buf &f, const Group &g,
                            const std::map<std::string, Item> *attrs) {
    using namespace detail;
    // children first, so that their object-header addresses are known
    std::map<std::string, Placed> kids;
    for (const auto &kv : g.datasets) kids[kv.first].header = write_dataset(f, kv.second);
    for (const auto &kv : g.groups) {
      if (kids.count(kv.first)) throw std::runtime_error("h5lite: name used twice: " + kv.first);
      kids[kv.first] = write_group(f, *kv.second, nullptr);
      kids[kv.first].is_group = true;
    }
    if (kids.size() > 2 * LEAF_K * 2 * INTERNAL_K)
      throw std::runtime_error("h5lite: too many links in one group");

    // local heap (spec III.D): "" at offset 0, then the names, then one free block
    Buf heap_data;
    heap_data.zeros(8);
    std::vector<uint64_t> name_off;
    for (const auto &kv : kids) {
      name_off.push_back(heap_data.size());
      heap_data.raw(kv.first.c_str(), kv.first.size() + 1);
      heap_data.align8();
    }
    const uint64_t free_off = heap_data.size();
    const uint64_t free_len = 32;
    heap_data.u64(1); // next free block: 1 terminates the list (H5HL_FREE_NULL)
    heap_data.u64(free_len);
    heap_data.zeros(free_len - 16);
    f.align8();
    Placed out;
    out.is_group = true;
    out.heap = f.size();
    f.raw("HEAP", 4); f.u8(0); f.u8(0); f.u16(0);
    f.u64(heap_data.size()); f.u64(free_off); f.u64(out.heap + 32);
    f.raw(heap_data.s.data(), heap_data.size());

    // symbol-table nodes: up to 2*LEAF_K entries each, names ascending
    std::vector<uint64_t> snod_addr, last_name;
    size_t idx = 0;
    auto it = kids.begin();
    while (it != kids.end()) {
      f.align8();
      const uint64_t at = f.size();
      const size_t n = std::min<size_t>(2 * LEAF_K, kids.size() - idx);
      f.raw("SNOD", 4); f.u8(1); f.u8(0); f.u16(uint16_t(n));
      for (size_t k = 0; k < n; k++, ++it, ++idx) {
        f.u64(name_off[idx]);
        f.u64(it->second.header);
        f.u32(it->second.is_group ? 1 : 0);
        f.u32(0);
        if (it->second.is_group) { f.u64(it->second.btree); f.u64(it->second.heap); }
        else f.zeros(16);
      }
      f.zeros(at + SNOD_SIZE - f.size());
      snod_addr.push_back(at);
      last_name.push_back(name_off[idx - 1]);
    }

    // one level-0 B-tree node (spec III.A.1): key[0] = "" ; key[i+1] = largest name of child i
    f.align8();
    out.btree = f.size();
    f.raw("TREE", 4); f.u8(0); f.u8(0); f.u16(uint16_t(snod_addr.size()));
    f.u64(UNDEF); f.u64(UNDEF);
    f.u64(0);
    for (size_t k = 0; k < snod_addr.size(); k++) { f.u64(snod_addr[k]); f.u64(last_name[k]); }
    f.zeros(out.btree + TREE_SIZE - f.size());

    // object header: symbol-table message + attributes
    Buf m;
    unsigned nmsgs = 1;
    { Buf b; b.u64(out.btree); b.u64(out.heap); message(m, 0x0011, b.s); }
    if (attrs)
      for (const auto &kv : *attrs) {
        const std::string dt = datatype_msg(kv.second.dtype), ds = dataspace_msg(kv.second.dims);
        if (kv.second.dtype == VSTR) throw std::runtime_error("h5lite: string attributes unsupported");
        Buf b; // attribute message v1 (spec IV.A.2.m): each part padded to 8
        b.u8(1); b.u8(0); b.u16(uint16_t(kv.first.size() + 1)); b.u16(uint16_t(dt.size()));
        b.u16(uint16_t(ds.size()));
        b.raw(kv.first.c_str(), kv.first.size() + 1); b.zeros(pad8(b.size()) - b.size());
        b.raw(dt.data(), dt.size()); b.zeros(pad8(b.size()) - b.size());
        b.raw(ds.data(), ds.size()); b.zeros(pad8(b.size()) - b.size());
        b.raw(kv.second.bytes.data(), kv.second.bytes.size());
        message(m, 0x000c, b.s);
        nmsgs++;
      }
    out.header = object_header(f, nmsgs, m);
    return out;
  }
};

// Reader for the subset above (also reads libhdf5 files restricted to it: v0/v1 superblock,
// v1 object headers incl. continuation blocks, symbol-table groups, contiguous or compact
// datasets, v1 attributes, variable-length strings).
class Reader {
public:
  std::map<std::string, Item> datasets; // full paths "group/name"
  std::map<std::string, Item> attrs;    // root attributes
  std::vector<std::string> skipped;     // items that could not be decoded, with the reason

  bool open(const std::string &path) {
    FILE *fp = std::fopen(path.c_str(), "rb");
    if (!fp) return false;
    std::string s;
    char tmp[1 << 16];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof tmp, fp)) > 0) s.append(tmp, n);
    std::fclose(fp);
    return parse(s);
  }

  static bool is_hdf5(const std::string &path) {
    FILE *fp = std::fopen(path.c_str(), "rb");
    if (!fp) return false;
    char sig[8] = {0};
    const size_t n = std::fread(sig, 1, 8, fp);
    std::fclose(fp);
    return n == 8 && std::memcmp(sig, "\x89HDF\r\n\x1a\n", 8) == 0;
  }

  bool parse(const std::string &s) {
    d_ = &s;
    datasets.clear();
    attrs.clear();
    skipped.clear();
    try {
      // the superblock sits at 0, 512, 1024, ... (user block)
      size_t sb = 0;
      for (;; sb = sb ? sb * 2 : 512) {
        if (sb + 8 > s.size()) return false;
        if (std::memcmp(s.data() + sb, "\x89HDF\r\n\x1a\n", 8) == 0) break;
      }
      const uint8_t ver = u8(sb + 8);
      if (ver > 1 || u8(sb + 13) != 8 || u8(sb + 14) != 8) return false;
      size_t p = sb + 24 + (ver == 1 ? 4 : 0);
      base_ = u64(p);
      const size_t root = p + 32;
      read_group(base_ + u64(root + 8), "", true);
    } catch (const std::exception &) {
      return false;
    }
    return true;
  }

private:
  struct Msg {
    uint16_t type;
    size_t at, size;
  };
  const std::string *d_ = nullptr;
  uint64_t base_ = 0;

  void need(size_t at, size_t n) const {
    if (at + n > d_->size() || at + n < at) throw std::runtime_error("h5lite: truncated file");
  }
  uint8_t u8(size_t at) const { need(at, 1); return uint8_t((*d_)[at]); }
  uint16_t u16(size_t at) const { return uint16_t(u8(at) | (u8(at + 1) << 8)); }
  uint32_t u32(size_t at) const { return uint32_t(u16(at)) | (uint32_t(u16(at + 2)) << 16); }
  uint64_t u64(size_t at) const { return uint64_t(u32(at)) | (uint64_t(u32(at + 4)) << 32); }

  std::vector<Msg> messages(uint64_t header) const {
    if (u8(header) != 1) throw std::runtime_error("h5lite: only version-1 object headers");
    std::vector<Msg> out;
    std::vector<std::pair<size_t, size_t>> blocks{{header + 16, u32(header + 8)}};
    for (size_t b = 0; b < blocks.size(); b++) {
      size_t p = blocks[b].first;
      const size_t end = p + blocks[b].second;
      while (p + 8 <= end) {
        const Msg m{u16(p), p + 8, u16(p + 2)};
        if (m.type == 0x0010) blocks.push_back({base_ + u64(m.at), u64(m.at + 8)});
        else out.push_back(m);
        p = m.at + m.size;
      }
    }
    return out;
  }

  uint8_t parse_type(size_t at) const {
    const uint8_t cls = u8(at) & 15;
    const uint32_t size = u32(at + 4);
    if (cls == 1 && size == 8) return F64;
    if (cls == 1 && size == 4) return F32;
    if (cls == 0 && size == 8) return U64;
    if (cls == 0 && size == 4) return U32;
    if (cls == 9 && (u8(at + 1) & 15) == 1) return VSTR;
    throw std::runtime_error("h5lite: unsupported datatype");
  }
  std::vector<uint64_t> parse_space(size_t at) const {
    const uint8_t ver = u8(at), rank = u8(at + 1);
    std::vector<uint64_t> dims;
    const size_t p = at + (ver == 1 ? 8 : 4);
    for (unsigned k = 0; k < rank; k++) dims.push_back(u64(p + 8 * k));
    return dims;
  }

  void fill(Item &it, const char *src, size_t nbytes) const {
    if (it.dtype == VSTR) {
      const size_t at = size_t(src - d_->data());
      const uint32_t len = u32(at);
      const uint64_t gcol = base_ + u64(at + 4);
      const uint32_t index = u32(at + 12);
      need(gcol, 16);
      if (std::memcmp(d_->data() + gcol, "GCOL", 4) != 0) throw std::runtime_error("h5lite: bad global heap");
      const size_t end = gcol + u64(gcol + 8);
      for (size_t p = gcol + 16; p + 16 <= end;) {
        const uint16_t idx = u16(p);
        const uint64_t sz = u64(p + 8);
        if (idx == index) {
          need(p + 16, len);
          it.bytes.assign(d_->data() + p + 16, len);
          return;
        }
        if (idx == 0) break;
        p += 16 + detail::pad8(sz);
      }
      throw std::runtime_error("h5lite: global heap object not found");
    }
    it.bytes.assign(src, nbytes);
  }

  void read_dataset(uint64_t header, const std::vector<Msg> &msgs, const std::string &path) {
    Item it;
    bool have_type = false, have_space = false;
    const Msg *layout = nullptr;
    for (const Msg &m : msgs) {
      if (m.type == 0x0003) { it.dtype = parse_type(m.at); have_type = true; }
      if (m.type == 0x0001) { it.dims = parse_space(m.at); have_space = true; }
      if (m.type == 0x0008) layout = &m;
    }
    (void)header;
    if (!have_type || !have_space || !layout) throw std::runtime_error("h5lite: incomplete dataset " + path);
    uint64_t n = 1;
    for (uint64_t d : it.dims) n *= d;
    const size_t nbytes = size_t(n) * (it.dtype == VSTR ? 16 : type_size(it.dtype));
    // layout message: v3 = version, class, ...; v1/v2 = version, rank+1, class, 5 reserved,
    // [address], dims (4 bytes each), [element size]
    const uint8_t lver = u8(layout->at);
    if (lver < 1 || lver > 3) throw std::runtime_error("h5lite: unknown layout version");
    const uint8_t cls = lver == 3 ? u8(layout->at + 1) : u8(layout->at + 2);
    if (cls == 1) {
      const uint64_t addr = u64(layout->at + (lver == 3 ? 2 : 8));
      if (nbytes && addr != detail::UNDEF) {
        need(base_ + addr, nbytes);
        fill(it, d_->data() + base_ + addr, nbytes);
      } else if (nbytes) {
        it.bytes.assign(nbytes, '\0'); // never written: fill value
      }
    } else if (cls == 0) {
      const size_t at = lver == 3 ? layout->at + 4 : layout->at + 8 + 4 * u8(layout->at + 1) + 4;
      need(at, nbytes);
      fill(it, d_->data() + at, nbytes);
    } else if (cls == 2 && lver == 3) {
      read_chunked(it, msgs, *layout, nbytes, path);
    } else {
      throw std::runtime_error("h5lite: layout of dataset " + path + " not supported");
    }
    datasets[path] = it;
  }

  // layout v3 class 2 (what the stock reference executable writes, src/writer.h): rank+1,
  // chunk B-tree address, chunk dims (last = element size); deflate needs zlib
  // (-DH5LITE_WITH_ZLIB).  Per the specification; not pinned against a libhdf5 file.
  struct Chunk {
    std::vector<uint64_t> offs;
    uint32_t size, mask;
    uint64_t addr;
  };
  void walk_chunks(uint64_t node, unsigned rank1, std::vector<Chunk> &out) const {
    need(node, 24);
    if (std::memcmp(d_->data() + node, "TREE", 4) != 0 || u8(node + 4) != 1)
      throw std::runtime_error("h5lite: bad chunk B-tree node");
    const uint8_t level = u8(node + 5);
    const unsigned used = u16(node + 6);
    const size_t ksize = 8 + 8 * rank1;
    size_t p = node + 24;
    for (unsigned k = 0; k < used; k++, p += ksize + 8) {
      const uint64_t child = base_ + u64(p + ksize);
      if (level) {
        walk_chunks(child, rank1, out);
        continue;
      }
      Chunk c;
      c.size = u32(p);
      c.mask = u32(p + 4);
      for (unsigned a = 0; a + 1 < rank1; a++) c.offs.push_back(u64(p + 8 + 8 * a));
      c.addr = child;
      out.push_back(c);
    }
  }
  void read_chunked(Item &it, const std::vector<Msg> &msgs, const Msg &layout, size_t nbytes,
                    const std::string &path) const {
    if (it.dtype == VSTR) throw std::runtime_error("h5lite: chunked strings not supported");
    const unsigned rank1 = u8(layout.at + 2), rank = rank1 - 1;
    const uint64_t btree = u64(layout.at + 3);
    std::vector<uint64_t> cdim(rank);
    for (unsigned a = 0; a < rank; a++) cdim[a] = u32(layout.at + 11 + 4 * a);
    const size_t esize = type_size(it.dtype);
    if (rank != it.dims.size() || rank == 0 || u32(layout.at + 11 + 4 * rank) != esize)
      throw std::runtime_error("h5lite: inconsistent chunked layout of " + path);
    // filter pipeline (message 0x000b, v1/v2): only deflate (1) is undone here
    std::vector<uint16_t> filters;
    for (const Msg &m : msgs) {
      if (m.type != 0x000b) continue;
      const uint8_t ver = u8(m.at), n = u8(m.at + 1);
      size_t p = m.at + (ver == 1 ? 8 : 2);
      for (unsigned f = 0; f < n; f++) {
        const uint16_t id = u16(p);
        size_t nlen = 0;
        if (ver == 1 || id >= 256) { nlen = u16(p + 2); p += 4; } else p += 2;
        const uint16_t ncd = u16(p + 2);
        p += 4 + (ver == 1 ? detail::pad8(nlen) : nlen) + 4 * size_t(ncd);
        if (ver == 1 && (ncd & 1)) p += 4;
        filters.push_back(id);
      }
    }
    it.bytes.assign(nbytes, '\0');
    if (!nbytes || btree == detail::UNDEF) return;
    std::vector<Chunk> chunks;
    walk_chunks(base_ + btree, rank1, chunks);
    size_t cbytes = esize;
    for (uint64_t c : cdim) cbytes *= size_t(c);
    std::string raw;
    for (const Chunk &c : chunks) {
      need(c.addr, c.size);
      const char *src = d_->data() + c.addr;
      for (size_t f = filters.size(); f-- > 0;) {
        if (c.mask & (1u << f)) continue;
        if (filters[f] != 1) throw std::runtime_error("h5lite: only the deflate filter is supported");
#ifdef H5LITE_WITH_ZLIB
        raw.assign(cbytes, '\0');
        uLongf n = uLongf(cbytes);
        if (uncompress(reinterpret_cast<Bytef *>(&raw[0]), &n, reinterpret_cast<const Bytef *>(src),
                       c.size) != Z_OK || n != cbytes)
          throw std::runtime_error("h5lite: cannot inflate a chunk of " + path);
        src = raw.data();
#else
        throw std::runtime_error("h5lite: built without zlib, cannot read deflated dataset " + path);
#endif
      }
      // copy the part of the chunk that lies inside the dataset, row by row of the last axis
      std::vector<uint64_t> idx(rank, 0), lim(rank);
      bool empty = false;
      for (unsigned a = 0; a < rank; a++) {
        if (c.offs[a] >= it.dims[a]) { empty = true; break; }
        lim[a] = std::min<uint64_t>(cdim[a], it.dims[a] - c.offs[a]);
      }
      if (empty) continue;
      for (;;) {
        size_t in = 0, out = 0;
        for (unsigned a = 0; a < rank; a++) {
          in = in * size_t(cdim[a]) + size_t(idx[a]);
          out = out * size_t(it.dims[a]) + size_t(c.offs[a] + idx[a]);
        }
        std::memcpy(&it.bytes[out * esize], src + in * esize, size_t(lim[rank - 1]) * esize);
        int a = int(rank) - 2;
        for (; a >= 0; a--) {
          if (++idx[a] < lim[a]) break;
          idx[a] = 0;
        }
        if (a < 0) break;
      }
    }
  }

  void read_attr(const Msg &m) {
    if (u8(m.at) != 1) throw std::runtime_error("h5lite: only version-1 attributes");
    const size_t nl = u16(m.at + 2), tl = u16(m.at + 4), sl = u16(m.at + 6);
    size_t p = m.at + 8;
    need(p, nl);
    const std::string name(d_->data() + p, nl ? nl - 1 : 0);
    p += detail::pad8(nl);
    Item it;
    it.dtype = parse_type(p);
    p += detail::pad8(tl);
    it.dims = parse_space(p);
    p += detail::pad8(sl);
    uint64_t n = 1;
    for (uint64_t d : it.dims) n *= d;
    const size_t nbytes = size_t(n) * type_size(it.dtype);
    need(p, nbytes);
    it.bytes.assign(d_->data() + p, nbytes);
    attrs[name] = it;
  }

  void walk_btree(uint64_t node, uint64_t heap_data, const std::string &prefix) {
    need(node, 24);
    if (std::memcmp(d_->data() + node, "TREE", 4) != 0) throw std::runtime_error("h5lite: bad B-tree node");
    const uint8_t level = u8(node + 5);
    const unsigned used = u16(node + 6);
    for (unsigned k = 0; k < used; k++) {
      const uint64_t child = base_ + u64(node + 24 + 8 + 16 * k);
      if (level > 0) {
        walk_btree(child, heap_data, prefix);
        continue;
      }
      need(child, 8);
      if (std::memcmp(d_->data() + child, "SNOD", 4) != 0) throw std::runtime_error("h5lite: bad symbol node");
      const unsigned n = u16(child + 6);
      for (unsigned e = 0; e < n; e++) {
        const size_t ent = child + 8 + 40 * e;
        const char *nm = d_->data() + heap_data + u64(ent);
        need(heap_data + u64(ent), 1);
        read_group(base_ + u64(ent + 8), prefix + nm, false);
      }
    }
  }

  // an object header that carries a symbol-table message is a group, otherwise a dataset
  void read_group(uint64_t header, const std::string &path, bool is_root) {
    const std::vector<Msg> msgs = messages(header);
    const Msg *stab = nullptr;
    for (const Msg &m : msgs)
      if (m.type == 0x0011) stab = &m;
    // an item this reader cannot decode (another datatype, a newer message version, ...)
    // is skipped and listed, it does not make the rest of the file unreadable
    if (!stab) {
      try {
        read_dataset(header, msgs, path);
      } catch (const std::runtime_error &e) {
        skipped.push_back(path + ": " + e.what());
      }
      return;
    }
    if (is_root)
      for (const Msg &m : msgs)
        if (m.type == 0x000c) {
          try {
            read_attr(m);
          } catch (const std::runtime_error &e) {
            skipped.push_back(std::string("attribute: ") + e.what());
          }
        }
    const uint64_t btree = base_ + u64(stab->at), heap = base_ + u64(stab->at + 8);
    need(heap, 32);
    if (std::memcmp(d_->data() + heap, "HEAP", 4) != 0) throw std::runtime_error("h5lite: bad local heap");
    const uint64_t heap_data = base_ + u64(heap + 24);
    walk_btree(btree, heap_data, path.empty() ? "" : path + "/");
  }
};

} // namespace h5lite
